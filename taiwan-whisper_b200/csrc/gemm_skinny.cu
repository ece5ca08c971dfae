// Skinny GEMM for the decode steps (M = batch <= 64 rows):  C[M,N] = epilogue(X[M,K] . W[N,K]^T + bias), bf16.
//
// At M <= 64 the GEMM is a pure weight stream (HBM-bound, 0.5-2 us of traffic per matrix), so the kernel is
// built for latency and bytes-in-flight rather than tensor throughput:
//   * grid = (N / BN) x S CTAs (BN = 16/32/64 weight rows, S-way split of K for the residual-add epilogue) so
//     that >= ~148 CTAs each pull a 20-40 KB slab of W;
//   * W goes global -> registers directly in mma.sync A-fragment order with 16-byte loads (each weight is used
//     exactly once; staging it in shared memory would only add latency).  A consistent permutation of K inside
//     each 32-wide chunk lets one 16-byte load feed two m16n8k16 MMAs;
//   * the activation slice X[:, kslice] is staged once per CTA with cp.async (padded pitch, conflict-free
//     16-byte reads) while the first W loads are already in flight;
//   * the 8 warps split (weight-row group) x (K sub-range); partial sums are reduced through shared memory and
//     the epilogue (bias, exact GELU, fp32 residual add — atomic when K is split across CTAs — or the
//     KV-cache scatter of the fused QKV projection) writes the [M, BN] tile.
// Legacy mma.sync is the right tool here: the work is ~0.2 GFLOP per launch and bandwidth-bound.
#include "kernels.cuh"

namespace tw {

constexpr int SK_THREADS = 256;
constexpr int SK_NT = 8;                 // 8 n-tiles of 8 batch rows = 64
constexpr int SK_MAXM = SK_NT * 8;
constexpr int SK_UNR = 5;                // K chunks (of 32) loaded per batch

__device__ __forceinline__ void mma_bf16_16816(float* c, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint4 ldg_stream16(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

// staged-row pitch with pitch % 128 == 64: the 16-byte reads of a quarter-warp (two batch rows) never collide
__host__ __device__ inline int sk_pitch(int chunks_per_cta) {
    const int p = chunks_per_cta * 64;
    return (p % 128 == 0) ? p + 64 : p;
}

struct SkinnyArgs {
    const __nv_bfloat16* X;
    int64_t ldx;
    const __nv_bfloat16* W;
    int64_t ldw;
    int M, N, K;
    int RG;              // weight-row groups (of 16) per CTA: 1, 2 or 4
    int chunks_per_cta;  // K chunks of 32 handled by one CTA (grid.y = S splits)
    GemmEpi epi;
};

__global__ void __launch_bounds__(SK_THREADS, 1)
gemm_skinny_kernel(SkinnyArgs a) {
    extern __shared__ __align__(16) unsigned char sk_smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int RG = a.RG, KG = 8 / RG;
    const int rg = warp % RG, kg = warp / RG;
    const int BN = 16 * RG;
    const int n0 = blockIdx.x * BN;
    const int total_chunks = a.K >> 5;
    const int c_begin = blockIdx.y * a.chunks_per_cta;
    const int c_end = min(total_chunks, c_begin + a.chunks_per_cta);
    const int n_chunks = c_end - c_begin;                         // chunks of this CTA
    const int pitch = sk_pitch(a.chunks_per_cta);                 // bytes per staged X row (padded)

    // this warp's K sub-range (in chunks, relative to c_begin) and weight rows
    const int w_begin = (n_chunks * kg) / KG, w_end = (n_chunks * (kg + 1)) / KG;
    const int row_a = n0 + rg * 16 + g, row_b = row_a + 8;
    const bool ok_a = row_a < a.N, ok_b = row_b < a.N;
    const __nv_bfloat16* wa = a.W + (int64_t)(ok_a ? row_a : 0) * a.ldw + (int64_t)c_begin * 32 + 8 * t;
    const __nv_bfloat16* wb = a.W + (int64_t)(ok_b ? row_b : 0) * a.ldw + (int64_t)c_begin * 32 + 8 * t;

    // first batch of W loads is issued before anything else
    uint4 ra[SK_UNR], rb[SK_UNR];
#pragma unroll
    for (int u = 0; u < SK_UNR; ++u) {
        const int c = w_begin + u;
        ra[u] = make_uint4(0, 0, 0, 0);
        rb[u] = make_uint4(0, 0, 0, 0);
        if (c < w_end) {
            if (ok_a) ra[u] = ldg_stream16(wa + (int64_t)c * 32);
            if (ok_b) rb[u] = ldg_stream16(wb + (int64_t)c * 32);
        }
    }

    // stage X[:, kslice] -> smem (rows >= M are zero)
    {
        const int vec_per_row = n_chunks * 4;                      // 16-byte vectors per row
        const int total = SK_MAXM * vec_per_row;
        for (int i = tid; i < total; i += SK_THREADS) {
            const int r = i / vec_per_row, v = i - r * vec_per_row;
            unsigned char* dst = sk_smem + (size_t)r * pitch + v * 16;
            if (r < a.M) {
                const __nv_bfloat16* src = a.X + (int64_t)r * a.ldx + (int64_t)c_begin * 32 + v * 8;
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src)
                             : "memory");
            } else {
                *reinterpret_cast<uint4*>(dst) = make_uint4(0, 0, 0, 0);
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();

    float acc[SK_NT][4];
#pragma unroll
    for (int j = 0; j < SK_NT; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.0f;

    for (int cb = w_begin; cb < w_end; cb += SK_UNR) {
        // prefetch the next batch of W while this one is consumed
        uint4 na[SK_UNR], nb[SK_UNR];
#pragma unroll
        for (int u = 0; u < SK_UNR; ++u) {
            const int c = cb + SK_UNR + u;
            na[u] = make_uint4(0, 0, 0, 0);
            nb[u] = make_uint4(0, 0, 0, 0);
            if (c < w_end) {
                if (ok_a) na[u] = ldg_stream16(wa + (int64_t)c * 32);
                if (ok_b) nb[u] = ldg_stream16(wb + (int64_t)c * 32);
            }
        }
#pragma unroll
        for (int u = 0; u < SK_UNR; ++u) {
            const int c = cb + u;
            if (c < w_end) {                                      // warp-uniform
                const unsigned char* xrow = sk_smem + (size_t)g * pitch + c * 64 + t * 16;
#pragma unroll
                for (int j = 0; j < SK_NT; ++j) {
                    const uint4 xb = *reinterpret_cast<const uint4*>(xrow + (size_t)j * 8 * pitch);
                    mma_bf16_16816(acc[j], ra[u].x, rb[u].x, ra[u].y, rb[u].y, xb.x, xb.y);
                    mma_bf16_16816(acc[j], ra[u].z, rb[u].z, ra[u].w, rb[u].w, xb.z, xb.w);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < SK_UNR; ++u) { ra[u] = na[u]; rb[u] = nb[u]; }
    }

    // ---- reduce the KG partial sums through shared memory: red[kg][n (BN)][m (64 + 1)]
    __syncthreads();                                               // everyone is done reading X
    float* red = reinterpret_cast<float*>(sk_smem);
    constexpr int MP = SK_MAXM + 1;
#pragma unroll
    for (int j = 0; j < SK_NT; ++j) {
        float* base = red + ((size_t)kg * BN + rg * 16) * MP;
        const int m = j * 8 + 2 * t;
        base[(size_t)g * MP + m] = acc[j][0];
        base[(size_t)g * MP + m + 1] = acc[j][1];
        base[(size_t)(g + 8) * MP + m] = acc[j][2];
        base[(size_t)(g + 8) * MP + m + 1] = acc[j][3];
    }
    __syncthreads();
    const GemmEpi& e = a.epi;
    const bool split = gridDim.y > 1;
    for (int i = tid; i < BN * a.M; i += SK_THREADS) {
        const int m = i / BN, nn = i - m * BN;
        const int n = n0 + nn;
        if (n >= a.N) continue;
        float v = 0.0f;
        for (int k = 0; k < KG; ++k) v += red[((size_t)k * BN + nn) * MP + m];
        if (e.bias && blockIdx.y == 0) v += e.bias[n];
        if (e.mode == EPI_STORE || e.mode == EPI_GELU) {
            if (e.mode == EPI_GELU) v = gelu_erf(v);
            if (n < e.n_split) reinterpret_cast<__nv_bfloat16*>(e.C)[(int64_t)m * e.ldc + n] = __float2bfloat16_rn(v);
            else
                reinterpret_cast<__nv_bfloat16*>(e.C2)[(n - e.n_split) +
                    (e.page_table ? kv_page_row(e.page_table, e.pt_stride, m, *e.d_row2) * e.row2_stride
                                  : (int64_t)m * e.ldc2 + (e.d_row2 ? (int64_t)(*e.d_row2) * e.row2_stride : 0))] = __float2bfloat16_rn(v);
        } else if (e.mode == EPI_RESID) {
            float* dst = reinterpret_cast<float*>(e.C) + (int64_t)m * e.ldc + n;
            if (split) atomicAdd(dst, v);
            else *dst += v;
        } else {
            reinterpret_cast<float*>(e.C)[(int64_t)m * e.ldc + n] = v;
        }
    }
}

bool gemm_skinny_supported(int M, int N, int K, const GemmEpi& epi) {
    if (M < 1 || M > SK_MAXM || (K & 31) || N < 16) return false;
    if (epi.mode == EPI_GELU_POS) return false;
    if (epi.mode != EPI_RESID && K > 1280) return false;          // the whole K slice of X must fit in shared memory
    return true;
}

int gemm_skinny(tw_ctx* ctx, const __nv_bfloat16* X, int64_t ldx, const __nv_bfloat16* W, int64_t ldw, int M, int N, int K,
                const GemmEpi& epi, cudaStream_t st) {
    if (!gemm_skinny_supported(M, N, K, epi) || (ldx % 8) || (ldw % 8) || (reinterpret_cast<uintptr_t>(X) & 15) ||
        (reinterpret_cast<uintptr_t>(W) & 15)) {
        ctx->set_error(TW_E_UNSUPPORTED, "gemm_skinny: unsupported shape / alignment");
        return TW_E_UNSUPPORTED;
    }
    static bool attr = false;
    if (!attr) {
        TW_CUDA_OK(ctx, cudaFuncSetAttribute(gemm_skinny_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr = true;
    }
    const int chunks = K / 32;
    SkinnyArgs a;
    a.X = X; a.ldx = ldx; a.W = W; a.ldw = ldw; a.M = M; a.N = N; a.K = K; a.epi = epi;
    int S = 1;
    if (epi.mode == EPI_RESID) {
        a.RG = 2;
        const int nblk = ceil_div(N, 32);
        S = ceil_div(ctx->sm_count, nblk);
        const int max_s = chunks / 4 > 0 ? chunks / 4 : 1;
        if (S > max_s) S = max_s;
        if (S < 1) S = 1;
        while (ceil_div(chunks, S) > 40) ++S;                      // K slice <= 1280
    } else {
        a.RG = (N / 64 >= ctx->sm_count) ? 4 : ((N / 32 >= 100) ? 2 : 1);
    }
    a.chunks_per_cta = ceil_div(chunks, S);
    S = ceil_div(chunks, a.chunks_per_cta);
    const int BN = 16 * a.RG, KG = 8 / a.RG;
    const size_t x_bytes = (size_t)SK_MAXM * sk_pitch(a.chunks_per_cta);
    const size_t red_bytes = (size_t)KG * BN * (SK_MAXM + 1) * sizeof(float);
    const size_t smem = x_bytes > red_bytes ? x_bytes : red_bytes;
    dim3 grid(ceil_div(N, BN), S);
    gemm_skinny_kernel<<<grid, SK_THREADS, smem, st>>>(a);
    TW_CUDA_OK(ctx, cudaGetLastError());
    return TW_OK;
}

}  // namespace tw
