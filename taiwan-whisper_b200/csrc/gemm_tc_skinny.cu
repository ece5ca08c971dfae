// tcgen05 GEMM for the decode steps (M <= 256 rows, in 64-row blocks):  C[M,N] = epilogue(A[M,K] . W[N,K]^T + bias), bf16 -> fp32.
//
// At M <= 64 the GEMM is a weight stream and its K loop is paced by how many bytes TMA has in flight: the general
// kernel (gemm_tc.cu) moves one 64-wide K block per TMA operation and measured ~205 ns per K block whether 8 or 16
// stages were in flight (tools/probes/gemm_trace.py) — the number of outstanding TMA operations, not bytes, is
// what saturates.  This variant makes each operation 4x larger: the operands are described by 3-D tensor maps
// (64 K-elements x rows x K-blocks) and one box fetches KSUB = 4 consecutive K blocks, landing as KSUB
// 128-byte-swizzled K-major tiles.  A is fetched as a 64-row box (8 KB per K block) and the MMA runs as M = 64:
// a tcgen05.mma of this size is paced by its A operand at ~one row per cycle whatever N is (measured: ~100-130
// cycles per M128 N32 K16 instruction here and in absorb.cu, ~63 per M64 N24 K16; profiles/r02_mma_small_n.md), so
// the M = 128 form of round 1 — which read 64 rows past the tile as accumulator rows nobody stored — paid twice
// for the K loop, and N up to 128 rides free.  Accumulator row r of an M = 64 instruction sits in TMEM lane
// (r % 16) + 32 (r / 16): every epilogue warp owns 16 rows in its lanes 0-15.
// Same roles as gemm_tc.cu: warp 0 TMA producer (weights prefetched before griddepcontrol.wait), warp 1 MMA
// issuer + TMEM owner, warps 2-5 epilogue (bias / GELU / fp32 residual add with split-K atomics / fp32 store /
// KV-cache scatter of the fused QKV projection).
// Batches above 64 rows (merged decode of several 64-clip batches) add a row-block index to the tile space instead of
// widening the instruction: an M = 128 instruction would double the K loop of every CTA (A-row pacing) and bring 128 rows
// of A per K block; two M = 64 tiles over a column tile of twice the width keep the per-CTA K loop and the tile count of
// the 64-row case (the second reader of a weight tile is served by L2).
#include <stdlib.h>

#include <map>
#include <tuple>

#include "tc_ptx.cuh"

namespace tw {

constexpr int SK2_THREADS = 192;
constexpr int SK2_KSUB = 4;                   // K blocks (of 64) per TMA box / pipeline stage

// AROWS = rows of the A box
template <int BN, int AROWS> struct Sk2Cfg {
    static constexpr int A_SUB = AROWS * 64 * 2;                           // A tile of one K block
    static constexpr int W_SUB = BN * 64 * 2;
    static constexpr int A_REGION = SK2_KSUB * A_SUB;
    static constexpr int STAGE_BYTES = SK2_KSUB * (A_SUB + W_SUB);         // 48 KB (BN 32) / 64 KB (BN 64)
    static constexpr int STAGES = (BN == 32) ? 4 : (BN == 64 ? 3 : 2);
    static constexpr int MIN_CTAS = 1;
    static constexpr int TMEM_COLS = 2 * BN;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
    static_assert(AROWS == 64, "the MMA runs as M = 64 over the whole A box");
};

__device__ __forceinline__ void tma_load_3d(const CUtensorMap* map, uint32_t bar, uint32_t dst, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

template <int BN, int AROWS>
__global__ void __launch_bounds__(SK2_THREADS, Sk2Cfg<BN, AROWS>::MIN_CTAS)
gemm_tc_skinny_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w, int M, int N, int K,
                      int ksplit, GemmEpi epi) {
    using Cfg = Sk2Cfg<BN, AROWS>;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + Cfg::STAGES;
    uint64_t* tmem_full = bars + 2 * Cfg::STAGES;
    uint64_t* tmem_empty = tmem_full + 2;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;     // warp-uniform role index
    const int tiles_n = (N + BN - 1) / BN;
    const int mblocks = (M + 63) / 64;                          // 64-row blocks of A (row blocks of one column tile are adjacent work items)
    const int num_tiles = tiles_n * mblocks * ksplit;
    const int k_blocks_total = K / 64;                          // K % 64 == 0 (checked by the launcher)
    const int kb_per_split = (k_blocks_total + ksplit - 1) / ksplit;

    pdl_trigger();
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_a);
        tma_prefetch_desc(&map_w);
        for (int s = 0; s < Cfg::STAGES; ++s) {
            mbar_init(smem_u32(&full_bar[s]), 1);
            mbar_init(smem_u32(&empty_bar[s]), 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(smem_u32(&tmem_full[s]), 1);
            mbar_init(smem_u32(&tmem_empty[s]), 4);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(Cfg::TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        // ===================== TMA producer (whole warp runs the loop, one elected lane issues) =====================
        // weights of the first work item do not depend on the previous kernel: stream them before the dependency wait
        int prefetched = 0;
        if ((int)blockIdx.x < num_tiles) {
            const int ks = blockIdx.x % ksplit, n_blk = blockIdx.x / (ksplit * mblocks);
            const int kb0 = ks * kb_per_split, kb1 = min(k_blocks_total, kb0 + kb_per_split);
            const int n_stage = (kb1 - kb0 + SK2_KSUB - 1) / SK2_KSUB;
            prefetched = min(Cfg::STAGES, n_stage);
            if (elect_one_sync()) {
                for (int i = 0; i < prefetched; ++i) {
                    const uint32_t fb = smem_u32(&full_bar[i]);
                    mbar_expect_tx(fb, Cfg::STAGE_BYTES);
                    tma_load_3d(&map_w, fb, smem_u32(smem + i * Cfg::STAGE_BYTES) + Cfg::A_REGION, 0, n_blk * BN, kb0 + i * SK2_KSUB);
                }
            }
            __syncwarp();
        }
        pdl_wait();
        int stage = 0;
        uint32_t phase = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const int ks = tile % ksplit, m_blk = (tile / ksplit) % mblocks, n_blk = tile / (ksplit * mblocks);
            const int kb0 = ks * kb_per_split, kb1 = min(k_blocks_total, kb0 + kb_per_split);
            // grouped GEMM: this tile's output columns belong to group (n_blk*BN)/group_n, whose A slice starts K columns later per group
            const int a_kb0 = epi.group_n > 0 ? ((n_blk * BN) / epi.group_n) * k_blocks_total : 0;
            for (int kb = kb0; kb < kb1; kb += SK2_KSUB) {
                const uint32_t fb = smem_u32(&full_bar[stage]);
                const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES);
                if (prefetched > 0) {
                    --prefetched;
                    if (elect_one_sync()) tma_load_3d(&map_a, fb, sa, 0, m_blk * 64, a_kb0 + kb);
                } else {
                    mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1);
                    if (elect_one_sync()) {
                        mbar_expect_tx(fb, Cfg::STAGE_BYTES);
                        tma_load_3d(&map_a, fb, sa, 0, m_blk * 64, a_kb0 + kb);
                        tma_load_3d(&map_w, fb, sa + Cfg::A_REGION, 0, n_blk * BN, kb);
                    }
                }
                __syncwarp();
                if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (whole warp runs the loop, one elected lane issues) =====================
        constexpr uint32_t idesc = make_idesc(64, BN);
        int stage = 0;
        uint32_t phase = 0;
        int it = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
            const int acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1;
            mbar_wait(smem_u32(&tmem_empty[acc]), acc_phase ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * BN;
            const int ks = tile % ksplit;
            const int kb0 = ks * kb_per_split, kb1 = min(k_blocks_total, kb0 + kb_per_split);
            for (int kb = kb0; kb < kb1; kb += SK2_KSUB) {
                mbar_wait(smem_u32(&full_bar[stage]), phase);
                tc_fence_after();
                const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES);
                const int nsub = min(SK2_KSUB, kb1 - kb);      // K blocks of this stage that belong to the split
                if (elect_one_sync()) {
                    for (int j = 0; j < nsub; ++j) {
                        const uint64_t a_desc = make_sw128_desc(sa + j * Cfg::A_SUB);
                        const uint64_t b_desc = make_sw128_desc(sa + Cfg::A_REGION + j * Cfg::W_SUB);
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            tc_mma_f16(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb > kb0 || j > 0 || k > 0) ? 1u : 0u);
                    }
                    tc_commit(smem_u32(&empty_bar[stage]));
                    if (kb + SK2_KSUB >= kb1) tc_commit(smem_u32(&tmem_full[acc]));
                }
                __syncwarp();
                if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else {
        // ===================== epilogue (warps 2..5); only lanes (rows) < M are stored =====================
        pdl_wait();
        const int quarter = warp & 3;
        int it = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
            const int ks = tile % ksplit, m_blk = (tile / ksplit) % mblocks, n_blk = tile / (ksplit * mblocks);
            const int acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1;
            mbar_wait(smem_u32(&tmem_full[acc]), acc_phase);
            tc_fence_after();
            const int row = m_blk * 64 + quarter * 16 + (lane & 15);     // M = 64 accumulator: 16 rows per TMEM lane quarter
            const bool row_ok = lane < 16 && row < M;
            const uint32_t t_row = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * BN;
            if (m_blk * 64 + quarter * 16 < M) {        // warp-uniform: quarters beyond M have nothing to store
#pragma unroll 1
                for (int c0 = 0; c0 < BN; c0 += 32) {
                    const int n0 = n_blk * BN + c0;
                    if (n0 >= N) break;
                    uint32_t r[32];
                    tmem_ld32(t_row + c0, r);
                    tmem_ld_wait();
                    if (!row_ok) continue;
                    float v[32];
                    const bool full = (n0 + 32 <= N);
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
                    if (epi.bias && ks == 0) {
                        if (full) {
#pragma unroll
                            for (int j = 0; j < 32; j += 4) {
                                const float4 b4 = __ldg(reinterpret_cast<const float4*>(epi.bias + n0 + j));
                                v[j] += b4.x; v[j + 1] += b4.y; v[j + 2] += b4.z; v[j + 3] += b4.w;
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; ++j) if (n0 + j < N) v[j] += __ldg(epi.bias + n0 + j);
                        }
                    }
                    if (epi.mode == EPI_STORE || epi.mode == EPI_GELU) {
                        if (epi.mode == EPI_GELU) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) v[j] = gelu_erf_fast(v[j]);
                        }
                        __nv_bfloat16* cp = reinterpret_cast<__nv_bfloat16*>(epi.C) + (int64_t)row * epi.ldc + n0;
                        if (n0 >= epi.n_split)       // fused QKV: K|V go straight into the cache row of this position
                            cp = reinterpret_cast<__nv_bfloat16*>(epi.C2) + (n0 - epi.n_split) +
                                 (epi.page_table ? kv_page_row(epi.page_table, epi.pt_stride, row, *epi.d_row2) * epi.row2_stride
                                                 : (int64_t)row * epi.ldc2 + (epi.d_row2 ? (int64_t)(*epi.d_row2) * epi.row2_stride : 0));
                        if (full && ((reinterpret_cast<uintptr_t>(cp) & 15) == 0)) {
#pragma unroll
                            for (int j = 0; j < 32; j += 8) {
                                uint4 pk;
                                __nv_bfloat162 h0 = __floats2bfloat162_rn(v[j], v[j + 1]), h1 = __floats2bfloat162_rn(v[j + 2], v[j + 3]);
                                __nv_bfloat162 h2 = __floats2bfloat162_rn(v[j + 4], v[j + 5]), h3 = __floats2bfloat162_rn(v[j + 6], v[j + 7]);
                                pk.x = *reinterpret_cast<uint32_t*>(&h0); pk.y = *reinterpret_cast<uint32_t*>(&h1);
                                pk.z = *reinterpret_cast<uint32_t*>(&h2); pk.w = *reinterpret_cast<uint32_t*>(&h3);
                                *reinterpret_cast<uint4*>(cp + j) = pk;
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; ++j) if (n0 + j < N) cp[j] = __float2bfloat16_rn(v[j]);
                        }
                    } else {
                        float* cp = reinterpret_cast<float*>(epi.C) + (int64_t)row * epi.ldc + n0;
                        const bool vec = full && ((reinterpret_cast<uintptr_t>(cp) & 15) == 0);
                        if (epi.mode == EPI_RESID && ksplit > 1) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) if (n0 + j < N) atomicAdd(cp + j, v[j]);
                        } else if (epi.mode == EPI_RESID) {
                            if (vec) {
#pragma unroll
                                for (int j = 0; j < 32; j += 4) {
                                    float4 c4 = *reinterpret_cast<float4*>(cp + j);
                                    c4.x += v[j]; c4.y += v[j + 1]; c4.z += v[j + 2]; c4.w += v[j + 3];
                                    *reinterpret_cast<float4*>(cp + j) = c4;
                                }
                            } else {
#pragma unroll
                                for (int j = 0; j < 32; ++j) if (n0 + j < N) cp[j] += v[j];
                            }
                        } else {
                            if (vec) {
#pragma unroll
                                for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(cp + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                            } else {
#pragma unroll
                                for (int j = 0; j < 32; ++j) if (n0 + j < N) cp[j] = v[j];
                            }
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&tmem_empty[acc]));
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(Cfg::TMEM_COLS) : "memory");
    }
}

// ---- host ----------------------------------------------------------------------------------------
typedef CUresult (*Sk2EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static Sk2EncodeFn g_sk2_encode = nullptr;
static bool g_sk2_broken = false;            // the driver rejected the 3-D map: the caller keeps using gemm_tc
using Sk2Key = std::tuple<const void*, int64_t, int64_t, int64_t, int>;
static std::map<Sk2Key, CUtensorMap> g_sk2_maps;

// 3-D view of a row-major [rows, K] bf16 matrix: (64 K-elements, rows, K/64 blocks); box = 64 x box_rows x KSUB
static int sk2_map(tw_ctx* ctx, const void* ptr, int64_t rows, int64_t K, int64_t ld, int box_rows, CUtensorMap* out) {
    const Sk2Key key(ptr, rows, K, ld, box_rows);
    auto it = g_sk2_maps.find(key);
    if (it != g_sk2_maps.end()) {
        *out = it->second;
        return TW_OK;
    }
    CUtensorMap m;
    const cuuint64_t dims[3] = {64, (cuuint64_t)rows, (cuuint64_t)(K / 64)};
    const cuuint64_t strides[2] = {(cuuint64_t)ld * 2, 128};
    const cuuint32_t box[3] = {64, (cuuint32_t)box_rows, (cuuint32_t)SK2_KSUB};
    const cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = g_sk2_encode(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        ctx->set_error(TW_E_CUDA, "gemm_tc_skinny: cuTensorMapEncodeTiled (3-D) failed (" + std::to_string((int)r) + ")");
        return TW_E_CUDA;
    }
    if (g_sk2_maps.size() > 4096) g_sk2_maps.clear();
    g_sk2_maps[key] = m;
    *out = m;
    return TW_OK;
}

bool gemm_tc_skinny_supported(int M, int N, int K, const GemmEpi& epi) {
    return !g_sk2_broken && M >= 1 && M <= 256 && (M <= 64 || epi.group_n == 0) && (K % 64) == 0 && K >= 64 && N >= 8 &&
           epi.mode != EPI_GELU_POS;
}

int gemm_tc_skinny_init(tw_ctx* ctx) {
    if (!g_sk2_encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
        if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) {
            ctx->set_error(TW_E_CUDA, "cuTensorMapEncodeTiled not available from the driver");
            return TW_E_CUDA;
        }
        g_sk2_encode = reinterpret_cast<Sk2EncodeFn>(fn);
    }
    TW_CUDA_OK(ctx, cudaFuncSetAttribute(gemm_tc_skinny_kernel<32, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, Sk2Cfg<32, 64>::SMEM_BYTES));
    TW_CUDA_OK(ctx, cudaFuncSetAttribute(gemm_tc_skinny_kernel<64, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, Sk2Cfg<64, 64>::SMEM_BYTES));
    TW_CUDA_OK(ctx, cudaFuncSetAttribute(gemm_tc_skinny_kernel<128, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, Sk2Cfg<128, 64>::SMEM_BYTES));
    // probe once whether the driver accepts the 3-D view (K-block stride 128 B < row stride)
    static __nv_bfloat16* probe = nullptr;
    if (!probe) {
        TW_CUDA_OK(ctx, cudaMalloc(&probe, 64 * 128 * 2));
        CUtensorMap m;
        if (sk2_map(ctx, probe, 64, 128, 128, 64, &m) != TW_OK) {
            g_sk2_broken = true;
            ctx->set_error(0, "");
        }
        g_sk2_maps.clear();
    }
    return TW_OK;
}

int gemm_tc_skinny(tw_ctx* ctx, const __nv_bfloat16* A, int64_t lda, const __nv_bfloat16* W, int64_t ldw, int M, int N, int K,
                   const GemmEpi& epi, cudaStream_t st) {
    if (!gemm_tc_skinny_supported(M, N, K, epi) || (lda % 8) || (ldw % 8) || (reinterpret_cast<uintptr_t>(A) & 15) ||
        (reinterpret_cast<uintptr_t>(W) & 15)) {
        ctx->set_error(TW_E_UNSUPPORTED, "gemm_tc_skinny: unsupported shape / alignment");
        return TW_E_UNSUPPORTED;
    }
    // Column-tile width: the K loop of a tile costs the same for 32, 64 or 128 columns (A-operand pacing, see the header), so the
    // narrowest width whose tile count (column tiles x 64-row blocks) fits one pass over the SMs streams the matrix with the most
    // CTAs (fc1 as 160 32-wide tiles on 148 CTAs instead of 80 64-wide ones is 12 ms per decode slower); very wide outputs
    // (vocabulary head) take 128-wide tiles.  The residual epilogue splits K (fp32 atomics) only as far as idle SMs remain: a
    // cost model that traded wider tiles for deeper splits (more CTAs, shorter K loops) measured 3-5 % SLOWER per decoder layer
    // at 64 and 128 rows — the extra atomic traffic on the residual stream costs more than the shorter loops save
    // (profiles/r02_merged_decode.md); so did 64-wide tiles with a two-way split for the unsplit residual GEMMs of merged batches
    // (fc2 alone 10.5 -> 7.2 us at 192 rows, the whole layer 308 -> 310 us).  TWB200_SK_BN forces a width.
    static const int force_bn = getenv("TWB200_SK_BN") ? atoi(getenv("TWB200_SK_BN")) : 0;
    const int mblocks = ceil_div(M, 64);
    const int kstages = ceil_div(K / 64, SK2_KSUB);
    auto split_for = [&](int bn) {
        if (epi.mode != EPI_RESID) return 1;
        int ks = ctx->sm_count / (ceil_div(N, bn) * mblocks);
        if (ks > kstages) ks = kstages;
        if (ks < 1) ks = 1;
        return ceil_div(kstages, ceil_div(kstages, ks));      // whole stages of KSUB blocks per split
    };
    int BN = 32;
    if (ceil_div(N, 32) * mblocks > ctx->sm_count) BN = 64;
    if (BN == 64 && ceil_div(N, 64) * mblocks > (mblocks > 1 ? 1 : 2) * ctx->sm_count) BN = 128;
    if (force_bn == 32 || force_bn == 64 || force_bn == 128) BN = force_bn;
    int n_groups = 1;
    if (epi.group_n > 0) {
        if (epi.group_n % 32 || N % epi.group_n || epi.mode == EPI_RESID) {
            ctx->set_error(TW_E_UNSUPPORTED, "gemm_tc_skinny: grouped GEMM needs group_n % 32 == 0, N % group_n == 0 and a store epilogue");
            return TW_E_UNSUPPORTED;
        }
        if (epi.group_n % BN) BN = 32;
        n_groups = N / epi.group_n;
    }
    CUtensorMap ma, mw;
    TW_CHECK(sk2_map(ctx, A, M, (int64_t)K * n_groups, lda, 64, &ma));
    TW_CHECK(sk2_map(ctx, W, N, K, ldw, BN, &mw));
    int tiles = ceil_div(N, BN) * mblocks;
    const int ksplit = split_for(BN);
    tiles *= ksplit;
    const int grid = tiles < ctx->sm_count ? tiles : ctx->sm_count;
    // the kernel splits K in units of K blocks: make kb_per_split a multiple of KSUB by construction
    // (k_blocks_total / ksplit rounded up to whole stages)
    if (BN == 128)
        TW_CUDA_OK(ctx, launch_k(gemm_tc_skinny_kernel<128, 64>, dim3(grid), dim3(SK2_THREADS), Sk2Cfg<128, 64>::SMEM_BYTES, st, ma, mw, M, N, K, ksplit, epi));
    else if (BN == 32)
        TW_CUDA_OK(ctx, launch_k(gemm_tc_skinny_kernel<32, 64>, dim3(grid), dim3(SK2_THREADS), Sk2Cfg<32, 64>::SMEM_BYTES, st, ma, mw, M, N, K, ksplit, epi));
    else
        TW_CUDA_OK(ctx, launch_k(gemm_tc_skinny_kernel<64, 64>, dim3(grid), dim3(SK2_THREADS), Sk2Cfg<64, 64>::SMEM_BYTES, st, ma, mw, M, N, K, ksplit, epi));
    TW_CUDA_OK(ctx, cudaGetLastError());
    return TW_OK;
}

}  // namespace tw
