// Attention kernels.
//  * encoder_attention_simt: tiled online-softmax attention in fp32 on CUDA cores — the fp32
//    check-mode path and the bring-up path for bf16 (the tcgen05 flash kernel is attention_tc.cu).
//  * decode_attention: one query per clip against a K|V row store (cross-attention K/V computed once
//    per window, or the self-attention cache).  HBM-bound streaming kernel: each CTA owns a
//    contiguous chunk of rows for ALL heads (rows are [K(d) | V(d)] contiguous), online softmax per
//    head in registers, partials merged by a second tiny kernel (flash-decoding split over time).
// Arithmetic: HF WhisperAttention.forward (modeling_whisper.py:284-358): q already scaled by
// head_dim^-0.5 (folded into the weights), softmax in fp32, no mask in the encoder.
#include "kernels.cuh"

namespace tw {

constexpr int HD = 64;   // Whisper head_dim is 64 for every checkpoint size

// ------------------------------------------------------------------------------------------------
constexpr int EA_BQ = 64, EA_BK = 64, EA_THREADS = 256;

struct EaSmem {
    float q[EA_BQ][HD + 1];
    float k[EA_BK][HD + 1];
    float v[EA_BK][HD + 4];
    float s[EA_BQ][EA_BK + 1];
    float m[EA_BQ], l[EA_BQ], scale[EA_BQ];
};

template <typename T>
__global__ void __launch_bounds__(EA_THREADS)
encoder_attention_simt_kernel(const T* __restrict__ qkv, T* __restrict__ out, int S, int H) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    EaSmem& sm = *reinterpret_cast<EaSmem*>(smem_raw);
    const int d = H * HD;
    const int q0 = blockIdx.x * EA_BQ, h = blockIdx.y, b = blockIdx.z;
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;        // 16 x 16 threads, 4 x 4 outputs each
    const T* base = qkv + (int64_t)b * S * 3 * d;

    for (int i = tid; i < EA_BQ * HD; i += EA_THREADS) {
        const int r = i / HD, c = i % HD;
        sm.q[r][c] = (q0 + r < S) ? to_f32(base[(int64_t)(q0 + r) * 3 * d + h * HD + c]) : 0.0f;
    }
    if (tid < EA_BQ) {
        sm.m[tid] = -INFINITY;
        sm.l[tid] = 0.0f;
    }
    float o[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) o[i][j] = 0.0f;

    for (int k0 = 0; k0 < S; k0 += EA_BK) {
        __syncthreads();
        for (int i = tid; i < EA_BK * HD; i += EA_THREADS) {
            const int r = i / HD, c = i % HD;
            const bool ok = k0 + r < S;
            const int64_t row = (int64_t)(k0 + r) * 3 * d;
            sm.k[r][c] = ok ? to_f32(base[row + d + h * HD + c]) : 0.0f;
            sm.v[r][c] = ok ? to_f32(base[row + 2 * d + h * HD + c]) : 0.0f;
        }
        __syncthreads();
        // S tile = Q K^T, 4x4 per thread (rows ty*4.., cols tx*4..)
        float acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;
#pragma unroll 8
        for (int c = 0; c < HD; ++c) {
            float a[4], bb[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = sm.q[ty * 4 + i][c];
#pragma unroll
            for (int j = 0; j < 4; ++j) bb[j] = sm.k[tx * 4 + j][c];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) sm.s[ty * 4 + i][tx * 4 + j] = (k0 + tx * 4 + j < S) ? acc[i][j] : -INFINITY;
        __syncthreads();
        // online softmax: warp w handles rows w*8 .. w*8+7, lanes cover 64 columns (2 each)
        {
            const int w = tid >> 5, lane = tid & 31;
            for (int r = w * 8; r < w * 8 + 8; ++r) {
                const float s0 = sm.s[r][lane], s1 = sm.s[r][lane + 32];
                const float mx = warp_max(fmaxf(s0, s1));
                const float m_old = sm.m[r];
                const float m_new = fmaxf(m_old, mx);
                const float p0 = __expf(s0 - m_new), p1 = __expf(s1 - m_new);
                const float ps = warp_sum(p0 + p1);
                sm.s[r][lane] = p0;
                sm.s[r][lane + 32] = p1;
                if (lane == 0) {
                    const float sc = (m_old == -INFINITY) ? 0.0f : __expf(m_old - m_new);
                    sm.scale[r] = sc;
                    sm.l[r] = sm.l[r] * sc + ps;
                    sm.m[r] = m_new;
                }
            }
        }
        __syncthreads();
        // O = O*scale + P V ; 4x4 per thread (rows ty*4.., dims tx*4..)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float sc = sm.scale[ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) o[i][j] *= sc;
        }
#pragma unroll 8
        for (int kk = 0; kk < EA_BK; ++kk) {
            const float4 vv = *reinterpret_cast<const float4*>(&sm.v[kk][tx * 4]);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float p = sm.s[ty * 4 + i][kk];
                o[i][0] = fmaf(p, vv.x, o[i][0]);
                o[i][1] = fmaf(p, vv.y, o[i][1]);
                o[i][2] = fmaf(p, vv.z, o[i][2]);
                o[i][3] = fmaf(p, vv.w, o[i][3]);
            }
        }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = q0 + ty * 4 + i;
        if (r >= S) continue;
        const float inv = 1.0f / sm.l[ty * 4 + i];
        T* orow = out + ((int64_t)b * S + r) * d + h * HD + tx * 4;
#pragma unroll
        for (int j = 0; j < 4; ++j) orow[j] = from_f32<T>(o[i][j] * inv);
    }
}

template <typename T>
void encoder_attention_simt(const T* qkv, T* out, int B, int S, int H, cudaStream_t st) {
    static bool attr_set[2] = {false, false};
    const int which = sizeof(T) == 4 ? 0 : 1;
    if (!attr_set[which]) {
        cudaFuncSetAttribute(encoder_attention_simt_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(EaSmem));
        attr_set[which] = true;
    }
    dim3 grid(ceil_div(S, EA_BQ), H, B);
    encoder_attention_simt_kernel<T><<<grid, EA_THREADS, sizeof(EaSmem), st>>>(qkv, out, S, H);
}
template void encoder_attention_simt<float>(const float*, float*, int, int, int, cudaStream_t);
template void encoder_attention_simt<__nv_bfloat16>(const __nv_bfloat16*, __nv_bfloat16*, int, int, int, cudaStream_t);

// ------------------------------------------------------------------------------------------------
// decode attention
constexpr int DA_ROWS = 64;        // rows per CTA
constexpr int DA_WARPS = 8;
constexpr int DA_MAXSLOT = 5;      // ceil(H*8/32) with H <= 20
constexpr int DA_PSTRIDE = HD + 2; // partial record: m, l, o[64]

int decode_attention_chunks(int Tk) { return ceil_div(Tk, DA_ROWS); }

template <typename T> struct Vec8;
template <> struct Vec8<float> {
    static __device__ __forceinline__ void load(const float* p, float* f) {
        const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
        f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
    }
};
template <> struct Vec8<__nv_bfloat16> {
    static __device__ __forceinline__ void load(const __nv_bfloat16* p, float* f) {
        uint4 raw;
        asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                     : "=r"(raw.x), "=r"(raw.y), "=r"(raw.z), "=r"(raw.w) : "l"(p));
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float2 t = __bfloat1622float2(h[i]);
            f[2 * i] = t.x;
            f[2 * i + 1] = t.y;
        }
    }
};

template <typename T>
__global__ void __launch_bounds__(DA_WARPS * 32)
decode_attention_partial(const T* __restrict__ q, int64_t q_stride, const T* __restrict__ kv, int64_t kv_clip_stride, int Tk, int H,
                         float* __restrict__ partial) {
    extern __shared__ __align__(16) float da_smem[];          // [DA_WARPS][H][DA_PSTRIDE]
    const int chunk = blockIdx.x, b = blockIdx.y, nchunks = gridDim.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int d = H * HD;
    const int nslots = H * 8;
    const T* kvb = kv + (int64_t)b * kv_clip_stride;
    const T* qb = q + (int64_t)b * q_stride;

    float qf[DA_MAXSLOT][8], of[DA_MAXSLOT][8], mrun[DA_MAXSLOT], lrun[DA_MAXSLOT];
#pragma unroll
    for (int s = 0; s < DA_MAXSLOT; ++s) {
        const int slot = lane + 32 * s;
        mrun[s] = -INFINITY;
        lrun[s] = 0.0f;
#pragma unroll
        for (int e = 0; e < 8; ++e) { of[s][e] = 0.0f; qf[s][e] = 0.0f; }
        if (slot < nslots) {
#pragma unroll
            for (int e = 0; e < 8; ++e) qf[s][e] = to_f32(qb[slot * 8 + e]);
        }
    }
    const int r_begin = chunk * DA_ROWS;
    const int r_end = min(Tk, r_begin + DA_ROWS);
    for (int r = r_begin + warp; r < r_end; r += DA_WARPS) {
        const T* row = kvb + (int64_t)r * 2 * d;
        float kf[DA_MAXSLOT][8], vf[DA_MAXSLOT][8];
#pragma unroll
        for (int s = 0; s < DA_MAXSLOT; ++s) {
            const int slot = lane + 32 * s;
            if (slot < nslots) {
                Vec8<T>::load(row + slot * 8, kf[s]);
                Vec8<T>::load(row + d + slot * 8, vf[s]);
            } else {
#pragma unroll
                for (int e = 0; e < 8; ++e) { kf[s][e] = 0.0f; vf[s][e] = 0.0f; }
            }
        }
        // math is unconditional (idle slots carry zeros) so the shuffles stay warp-converged
#pragma unroll
        for (int s = 0; s < DA_MAXSLOT; ++s) {
            float dot = 0.0f;
#pragma unroll
            for (int e = 0; e < 8; ++e) dot = fmaf(qf[s][e], kf[s][e], dot);
            dot += __shfl_xor_sync(0xffffffffu, dot, 1);
            dot += __shfl_xor_sync(0xffffffffu, dot, 2);
            dot += __shfl_xor_sync(0xffffffffu, dot, 4);
            const float m_new = fmaxf(mrun[s], dot);
            const float sc = __expf(mrun[s] - m_new);     // exp(-inf) = 0 on the first row
            const float p = __expf(dot - m_new);
            lrun[s] = lrun[s] * sc + p;
            mrun[s] = m_new;
#pragma unroll
            for (int e = 0; e < 8; ++e) of[s][e] = fmaf(p, vf[s][e], of[s][e] * sc);
        }
    }
    // per-warp records -> smem
#pragma unroll
    for (int s = 0; s < DA_MAXSLOT; ++s) {
        const int slot = lane + 32 * s;
        if (slot < nslots) {
            const int h = slot >> 3, e0 = (slot & 7) * 8;
            float* rec = da_smem + ((int64_t)warp * H + h) * DA_PSTRIDE;
            if ((slot & 7) == 0) { rec[0] = mrun[s]; rec[1] = lrun[s]; }
#pragma unroll
            for (int e = 0; e < 8; ++e) rec[2 + e0 + e] = of[s][e];
        }
    }
    __syncthreads();
    // merge the warps: thread -> (head, dim)
    for (int i = threadIdx.x; i < H * HD; i += blockDim.x) {
        const int h = i / HD, e = i % HD;
        float m = -INFINITY;
#pragma unroll
        for (int w = 0; w < DA_WARPS; ++w) m = fmaxf(m, da_smem[((int64_t)w * H + h) * DA_PSTRIDE]);
        float l = 0.0f, o = 0.0f;
#pragma unroll
        for (int w = 0; w < DA_WARPS; ++w) {
            const float* rec = da_smem + ((int64_t)w * H + h) * DA_PSTRIDE;
            const float sc = (rec[0] == -INFINITY) ? 0.0f : __expf(rec[0] - m);
            l += rec[1] * sc;
            o += rec[2 + e] * sc;
        }
        float* prec = partial + (((int64_t)b * nchunks + chunk) * H + h) * DA_PSTRIDE;
        if (e == 0) { prec[0] = m; prec[1] = l; }
        prec[2 + e] = o;
    }
}

template <typename T>
__global__ void __launch_bounds__(HD)
decode_attention_combine(const float* __restrict__ partial, int nchunks, int H, T* __restrict__ out) {
    const int h = blockIdx.x, b = blockIdx.y, e = threadIdx.x;
    float m = -INFINITY;
    for (int c = 0; c < nchunks; ++c) m = fmaxf(m, partial[(((int64_t)b * nchunks + c) * H + h) * DA_PSTRIDE]);
    float l = 0.0f, o = 0.0f;
    for (int c = 0; c < nchunks; ++c) {
        const float* rec = partial + (((int64_t)b * nchunks + c) * H + h) * DA_PSTRIDE;
        const float sc = (rec[0] == -INFINITY) ? 0.0f : __expf(rec[0] - m);
        l += rec[1] * sc;
        o += rec[2 + e] * sc;
    }
    out[(int64_t)b * H * HD + h * HD + e] = from_f32<T>(o / l);
}

template <typename T>
void decode_attention(const T* q, int64_t q_stride, const T* kv, int64_t kv_clip_stride, int Tk, int B, int H, float* partial, T* out,
                      cudaStream_t st, cudaEvent_t ev0, cudaEvent_t ev1) {
    const int nchunks = decode_attention_chunks(Tk);
    const size_t smem = (size_t)DA_WARPS * H * DA_PSTRIDE * sizeof(float);   // <= 42 KB
    dim3 grid(nchunks, B);
    if (ev0) cudaEventRecord(ev0, st);
    decode_attention_partial<T><<<grid, DA_WARPS * 32, smem, st>>>(q, q_stride, kv, kv_clip_stride, Tk, H, partial);
    if (ev1) cudaEventRecord(ev1, st);
    decode_attention_combine<T><<<dim3(H, B), HD, 0, st>>>(partial, nchunks, H, out);
}
template void decode_attention<float>(const float*, int64_t, const float*, int64_t, int, int, int, float*, float*, cudaStream_t,
                                      cudaEvent_t, cudaEvent_t);
template void decode_attention<__nv_bfloat16>(const __nv_bfloat16*, int64_t, const __nv_bfloat16*, int64_t, int, int, int, float*,
                                              __nv_bfloat16*, cudaStream_t, cudaEvent_t, cudaEvent_t);

}  // namespace tw
