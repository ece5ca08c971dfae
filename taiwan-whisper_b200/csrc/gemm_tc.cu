// tcgen05 / TMEM / TMA GEMM for sm_100a:  C[M,N] = epilogue(A[M,K] . W[N,K]^T + bias), bf16 x bf16 -> fp32.
//
// Persistent, warp-specialised, one CTA per SM:
//   warp 0      TMA producer   : cp.async.bulk.tensor 2D loads of A (128 x 64) and W (BN x 64) tiles, 128-byte
//                                swizzle, into a 4-deep shared-memory ring guarded by full/empty mbarriers
//   warp 1      MMA issuer     : one thread issues tcgen05.mma.cta_group::1.kind::f16 (M=128, N=BN, K=16) from
//                                shared-memory descriptors into a double-buffered TMEM accumulator and
//                                tcgen05.commit's the ring slot / the accumulator to mbarriers
//   warps 2-5   epilogue       : tcgen05.ld the accumulator (thread = row, 32 columns per load), fused
//                                bias / exact GELU / fp32 residual add / sinusoid add, 32-byte-sector stores;
//                                overlaps the next tile's MMAs through the second TMEM buffer
// Edge tiles rely on TMA out-of-bounds zero fill (M, N, K tails) and guarded stores.
#include <cuda.h>

#include <map>
#include <tuple>

#include "kernels.cuh"
#include "tc_ptx.cuh"

namespace tw {

constexpr int TC_BM = 128;
constexpr int TC_BK = 64;            // 64 bf16 = 128 bytes = one swizzle row
constexpr int TC_THREADS = 192;      // 6 warps
constexpr int TC_EPI_WARP0 = 2;

template <int BN> struct TcCfg {
    static constexpr int STAGES = (BN == 256) ? 4 : (BN == 128 ? 6 : 8);     // BN 64 / 32: 8 stages of 24 / 20 KB
    static constexpr int A_BYTES = TC_BM * TC_BK * 2;
    static constexpr int B_BYTES = BN * TC_BK * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    // (probed, tools/probes/gemm_trace.py: at M = 64 the K loop runs at ~205 ns per K block whether 8 or 16 stages are in
    // flight and whether 1 or 4 independent accumulators are used — it is paced by TMA delivery of the 96 x 128-byte rows
    // per block, ~60 GB/s per SM, two thirds of which is the A tile every CTA re-reads from L2)
    static constexpr int NACC = 1;
    static constexpr int TMEM_COLS = 2 * BN * NACC;          // double-buffered fp32 accumulator (power of two >= 32)
    // BN >= 128 (the M >> 128 GEMMs): the epilogue stages 128 x 128-byte boxes in two swizzled buffers and hands them to
    // TMA (tensor store, or tensor reduce-add for the fp32 residual stream)
    static constexpr bool TMA_EPI = (BN >= 128);
    static constexpr int EPI_BYTES = TMA_EPI ? 2 * TC_BM * 128 : 0;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_BYTES + 1024 /*align*/ + 256 /*barriers*/;
};

#ifdef TW_GEMM_TRACE
// probe build only (tools/probes/gemm_trace.py): per-CTA phase timestamps
__device__ unsigned long long g_trace[256 * 8];
__device__ __forceinline__ void trace(int slot) {
    if (blockIdx.x < 256) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        g_trace[blockIdx.x * 8 + slot] = t;
    }
}
#define TW_TRACE(slot) trace(slot)
#else
#define TW_TRACE(slot)
#endif

template <int BN>
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w,
               const __grid_constant__ CUtensorMap map_c, int M, int N, int K, int ksplit, GemmEpi epi) {
    using Cfg = TcCfg<BN>;
    extern __shared__ unsigned char smem_raw[];
    // 1024-byte alignment for the 128-byte swizzle atoms
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char* epi_smem = smem + Cfg::STAGES * Cfg::STAGE_BYTES;      // [2][128 rows][128 B] (TMA_EPI only)
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES + Cfg::EPI_BYTES);
    uint64_t* full_bar = bars;                       // [STAGES]
    uint64_t* empty_bar = bars + Cfg::STAGES;        // [STAGES]
    uint64_t* tmem_full = bars + 2 * Cfg::STAGES;    // [2]
    uint64_t* tmem_empty = tmem_full + 2;            // [2]
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;     // warp-uniform role index
    const int tiles_m = (M + TC_BM - 1) / TC_BM, tiles_n = (N + BN - 1) / BN;
    // a work item = (output tile, K split): split-K (ksplit > 1) is used by the residual-add epilogue of the skinny
    // decode GEMMs, whose partial sums are combined with fp32 atomics
    const int num_tiles = tiles_m * tiles_n * ksplit;
    const int k_blocks_total = (K + TC_BK - 1) / TC_BK;
    const int kb_per_split = (k_blocks_total + ksplit - 1) / ksplit;

    pdl_trigger();
    if (threadIdx.x == 0) TW_TRACE(0);
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_a);
        tma_prefetch_desc(&map_w);
        if (Cfg::TMA_EPI) tma_prefetch_desc(&map_c);
        for (int s = 0; s < Cfg::STAGES; ++s) {
            mbar_init(smem_u32(&full_bar[s]), 1);
            mbar_init(smem_u32(&empty_bar[s]), 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(smem_u32(&tmem_full[s]), 1);
            mbar_init(smem_u32(&tmem_empty[s]), 4);      // one arrive per epilogue warp
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
    if (threadIdx.x == 0) TW_TRACE(1);

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            // PDL prologue: the weight tiles of the first work item do not depend on the previous kernel — start
            // streaming them (up to one ring of stages) before waiting for the producer of A
            int prefetched = 0;
            if ((int)blockIdx.x < num_tiles) {
                const int ks = blockIdx.x % ksplit, mn = blockIdx.x / ksplit;
                const int n_blk = mn % tiles_n;
                const int kb0 = ks * kb_per_split, kb1 = min(k_blocks_total, kb0 + kb_per_split);
                prefetched = min(Cfg::STAGES, kb1 - kb0);
                for (int i = 0; i < prefetched; ++i) {
                    const uint32_t fb = smem_u32(&full_bar[i]);
                    mbar_expect_tx(fb, Cfg::STAGE_BYTES);
                    tma_load_2d(&map_w, fb, smem_u32(smem + i * Cfg::STAGE_BYTES) + Cfg::A_BYTES, (kb0 + i) * TC_BK, n_blk * BN);
                }
            }
            pdl_wait();
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int ks = tile % ksplit, mn = tile / ksplit;
                const int m_blk = mn / tiles_n, n_blk = mn % tiles_n;
                const int kb0 = ks * kb_per_split, kb1 = min(k_blocks_total, kb0 + kb_per_split);
                for (int kb = kb0; kb < kb1; ++kb) {
                    const uint32_t fb = smem_u32(&full_bar[stage]);
                    const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES);
                    if (prefetched > 0) {            // W already in flight for this stage: only A is missing
                        --prefetched;
                        tma_load_2d(&map_a, fb, sa, kb * TC_BK, m_blk * TC_BM);
                    } else {
                        mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1);
                        mbar_expect_tx(fb, Cfg::STAGE_BYTES);
                        tma_load_2d(&map_a, fb, sa, kb * TC_BK, m_blk * TC_BM);
                        tma_load_2d(&map_w, fb, sa + Cfg::A_BYTES, kb * TC_BK, n_blk * BN);
                    }
                    if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (whole warp runs the loop, one elected lane issues: tc_ptx.cuh) =====================
        constexpr uint32_t idesc = make_idesc(TC_BM, BN);
        int stage = 0;
        uint32_t phase = 0;
        int it = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
            const int acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1;
            mbar_wait(smem_u32(&tmem_empty[acc]), acc_phase ^ 1);     // epilogue has drained this buffer
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * (BN * Cfg::NACC);
            const int ks = tile % ksplit;
            const int kb0 = ks * kb_per_split, kb1 = min(k_blocks_total, kb0 + kb_per_split);
            int kstep = 0;                          // K steps issued for this tile
            for (int kb = kb0; kb < kb1; ++kb) {
                mbar_wait(smem_u32(&full_bar[stage]), phase);
                if (it == 0 && kb == kb0 && lane == 0) TW_TRACE(2);
                tc_fence_after();
                const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES);
                if (elect_one_sync()) {
                    const uint64_t a_desc = make_sw128_desc(sa);
                    const uint64_t b_desc = make_sw128_desc(sa + Cfg::A_BYTES);
#pragma unroll
                    for (int k = 0; k < TC_BK / 16; ++k) {
                        // advance 16 elements (32 bytes) along K inside the swizzle atom: +2 in the >>4 address field
                        const int a_idx = (kstep + k) % Cfg::NACC;
                        tc_mma_f16(d_tmem + a_idx * BN, a_desc + 2 * k, b_desc + 2 * k, idesc, (kstep + k >= Cfg::NACC) ? 1u : 0u);
                    }
                    tc_commit(smem_u32(&empty_bar[stage]));               // frees the smem slot when the MMAs retire
                    if (kb + 1 == kb1) tc_commit(smem_u32(&tmem_full[acc]));      // accumulator ready
                }
                __syncwarp();
                kstep += TC_BK / 16;
                if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
            }
            if (lane == 0) TW_TRACE(3);
        }
    } else {
        // ===================== epilogue (warps 2..5) =====================
        pdl_wait();
        const int quarter = warp & 3;              // TMEM lane quarter this warp may access
        int it = 0;
        int epi_box = 0;                           // boxes handed to TMA so far (TMA_EPI)
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
            const int ks = tile % ksplit, mn = tile / ksplit;
            const int m_blk = mn / tiles_n, n_blk = mn % tiles_n;
            const int acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1;
            mbar_wait(smem_u32(&tmem_full[acc]), acc_phase);
            if (warp == TC_EPI_WARP0 && lane == 0) TW_TRACE(4);
            tc_fence_after();
            const int row = m_blk * TC_BM + quarter * 32 + lane;
            if constexpr (Cfg::TMA_EPI) {
                // ---- TMA epilogue: registers -> swizzled smem box (128 rows x 128 B) -> tensor store / reduce-add.
                // Row / column tails are clipped by the tensor map, so no guards are needed on the stores.
                const bool out_bf16 = (epi.mode == EPI_STORE || epi.mode == EPI_GELU);
                const int cols_per_box = out_bf16 ? 64 : 32;
                const uint32_t t_row = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * (BN * Cfg::NACC);
                const int r_in_tile = quarter * 32 + lane;
                const bool issuer = (warp == TC_EPI_WARP0 && lane == 0);
                for (int cb = 0; cb < BN; cb += cols_per_box) {
                    const int nb0 = n_blk * BN + cb;
                    if (nb0 >= N) break;                                   // uniform over the 4 warps
                    unsigned char* buf = epi_smem + (epi_box & 1) * (TC_BM * 128);
                    if (epi_box >= 2) {                                    // the store issued from this buffer two boxes ago
                        if (issuer) tma_store_wait_read<1>();
                        epi_bar_sync();
                    }
                    unsigned char* rowp = buf + r_in_tile * 128;
#pragma unroll 1
                    for (int c0 = 0; c0 < cols_per_box; c0 += 32) {
                        const int n0 = nb0 + c0;
                        uint32_t r[32];
                        tmem_ld32(t_row + cb + c0, r);
                        tmem_ld_wait();
                        float v[32];
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
                        if (epi.bias && ks == 0) {
                            if (n0 + 32 <= N) {
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
                        if (epi.mode == EPI_GELU) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) v[j] = gelu_erf_fast(v[j]);
                        } else if (epi.mode == EPI_GELU_POS) {
                            const float* pp = epi.pos + (int64_t)(row % epi.pos_period) * N + n0;
#pragma unroll
                            for (int j = 0; j < 32; ++j) v[j] = gelu_erf_fast(v[j]) + ((n0 + j < N) ? __ldg(pp + j) : 0.0f);
                        }
                        if (out_bf16) {             // 32 columns = 64 bytes = 16-byte chunks (c0/8) .. (c0/8)+3 of the 128-byte row
#pragma unroll
                            for (int qd = 0; qd < 4; ++qd) {
                                __nv_bfloat162 h0 = __floats2bfloat162_rn(v[8 * qd], v[8 * qd + 1]), h1 = __floats2bfloat162_rn(v[8 * qd + 2], v[8 * qd + 3]);
                                __nv_bfloat162 h2 = __floats2bfloat162_rn(v[8 * qd + 4], v[8 * qd + 5]), h3 = __floats2bfloat162_rn(v[8 * qd + 6], v[8 * qd + 7]);
                                uint4 pk;
                                pk.x = *reinterpret_cast<uint32_t*>(&h0); pk.y = *reinterpret_cast<uint32_t*>(&h1);
                                pk.z = *reinterpret_cast<uint32_t*>(&h2); pk.w = *reinterpret_cast<uint32_t*>(&h3);
                                const int chunk = (c0 >> 3) + qd;
                                *reinterpret_cast<uint4*>(rowp + ((chunk ^ (r_in_tile & 7)) << 4)) = pk;
                            }
                        } else {                    // 32 fp32 columns = the whole 128-byte row
#pragma unroll
                            for (int qd = 0; qd < 8; ++qd)
                                *reinterpret_cast<float4*>(rowp + ((qd ^ (r_in_tile & 7)) << 4)) =
                                    make_float4(v[4 * qd], v[4 * qd + 1], v[4 * qd + 2], v[4 * qd + 3]);
                        }
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    epi_bar_sync();
                    if (issuer) {
                        if (epi.mode == EPI_RESID) tma_reduce_add_2d(&map_c, smem_u32(buf), nb0, m_blk * TC_BM);
                        else tma_store_2d(&map_c, smem_u32(buf), nb0, m_blk * TC_BM);
                        tma_store_commit();
                    }
                    ++epi_box;
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&tmem_empty[acc]));
                continue;
            }
            const bool row_ok = row < M;
            const uint32_t t_row = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * (BN * Cfg::NACC);
            // a short K range (split-K) may not have touched every accumulator
            const int kb_n = min(k_blocks_total, ks * kb_per_split + kb_per_split) - ks * kb_per_split;
            const int n_acc = min(Cfg::NACC, kb_n * (TC_BK / 16));
#pragma unroll 1
            for (int c0 = 0; c0 < BN; c0 += 32) {
                const int n0 = n_blk * BN + c0;
                if (n0 >= N) break;                // warp-uniform
                uint32_t r[32];
                tmem_ld32(t_row + c0, r);
                tmem_ld_wait();
                float v[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
                for (int a = 1; a < n_acc; ++a) {
                    tmem_ld32(t_row + a * BN + c0, r);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] += __uint_as_float(r[j]);
                }
                if (!row_ok) continue;
                const bool full = (n0 + 32 <= N);
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
                const int64_t o = (int64_t)row * epi.ldc + n0;
                if (epi.mode == EPI_STORE || epi.mode == EPI_GELU) {
                    if (epi.mode == EPI_GELU) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = gelu_erf_fast(v[j]);
                    }
                    __nv_bfloat16* cp = reinterpret_cast<__nv_bfloat16*>(epi.C) + o;
                    if (n0 >= epi.n_split)       // column split (fused QKV: K|V go straight into the cache row of this position)
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
                    float* cp = reinterpret_cast<float*>(epi.C) + o;
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
                        if (epi.mode == EPI_GELU_POS) {
                            const float* pp = epi.pos + (int64_t)(row % epi.pos_period) * N + n0;
#pragma unroll
                            for (int j = 0; j < 32; ++j) if (n0 + j < N) v[j] = gelu_erf_fast(v[j]) + __ldg(pp + j);
                        }
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
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&tmem_empty[acc]));
        }
    }

    if (warp == TC_EPI_WARP0 && lane == 0) TW_TRACE(5);
    if (Cfg::TMA_EPI && warp == TC_EPI_WARP0 && lane == 0) tma_store_wait_read<0>();   // smem must outlive the bulk stores
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(Cfg::TMEM_COLS) : "memory");
    }
    if (threadIdx.x == 0) TW_TRACE(6);
}

#ifdef TW_GEMM_TRACE
extern "C" __attribute__((visibility("default"))) void tw_debug_trace_copy(void* dst_dev) {
    void* p = nullptr;
    cudaGetSymbolAddress(&p, g_trace);
    cudaMemcpy(dst_dev, p, sizeof(unsigned long long) * 256 * 8, cudaMemcpyDeviceToDevice);
    cudaMemset(p, 0, sizeof(unsigned long long) * 256 * 8);
}
#endif

// ---- host side ---------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;

using MapKey = std::tuple<const void*, int64_t, int64_t, int64_t, int, int>;
static std::map<MapKey, CUtensorMap> g_maps;     // per process; a ctx is per device per process

int gemm_tc_init(tw_ctx* ctx) {
    if (!g_encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
        if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) {
            ctx->set_error(TW_E_CUDA, "cuTensorMapEncodeTiled not available from the driver");
            return TW_E_CUDA;
        }
        g_encode = reinterpret_cast<EncodeTiledFn>(fn);
    }
    TW_CUDA_OK(ctx, cudaFuncSetAttribute(gemm_tc_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfg<256>::SMEM_BYTES));
    TW_CUDA_OK(ctx, cudaFuncSetAttribute(gemm_tc_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfg<128>::SMEM_BYTES));
    TW_CUDA_OK(ctx, cudaFuncSetAttribute(gemm_tc_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfg<32>::SMEM_BYTES));
    TW_CUDA_OK(ctx, cudaFuncSetAttribute(gemm_tc_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfg<64>::SMEM_BYTES));
    return TW_OK;
}

// 2D row-major tensor map with 128-byte swizzle; esize 2 = bf16 (box 64 columns), 4 = f32 (box 32 columns)
static int get_map(tw_ctx* ctx, const void* ptr, int64_t rows, int64_t cols, int64_t ld, int box_rows, CUtensorMap* out, int esize = 2) {
    const MapKey key(ptr, rows, cols, ld, box_rows, esize);
    auto it = g_maps.find(key);
    if (it != g_maps.end()) {
        *out = it->second;
        return TW_OK;
    }
    CUtensorMap m;
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)ld * esize};
    const cuuint32_t box[2] = {(cuuint32_t)(128 / esize), (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode(&m, esize == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                          const_cast<void*>(ptr), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        ctx->set_error(TW_E_CUDA, "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ") rows=" + std::to_string(rows) +
                                      " cols=" + std::to_string(cols) + " ld=" + std::to_string(ld));
        return TW_E_CUDA;
    }
    if (g_maps.size() > 4096) g_maps.clear();
    g_maps[key] = m;
    *out = m;
    return TW_OK;
}

int gemm_tc(tw_ctx* ctx, const __nv_bfloat16* A, int64_t lda, const __nv_bfloat16* W, int64_t ldw, int M, int N, int K,
            const GemmEpi& epi, cudaStream_t st) {
    if (M <= 0 || N <= 0 || K <= 0) return TW_OK;
    if ((lda % 8) || (ldw % 8) || (reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(W) & 15)) {
        ctx->set_error(TW_E_UNSUPPORTED, "gemm_tc: operands must be 16-byte aligned with row pitch a multiple of 8 elements");
        return TW_E_UNSUPPORTED;
    }
    // skinny (decode, M <= 128): the GEMM streams W once; narrow N tiles spread it over many SMs
    // (N tiles of 32 unless that gives more tiles than SMs, then 64 so that one wave covers the matrix)
    const int BN = (M <= TC_BM) ? ((ceil_div(N, 32) > ctx->sm_count) ? 64 : 32) : ((N > 128) ? 256 : 128);
    CUtensorMap ma, mw, mc;
    TW_CHECK(get_map(ctx, A, M, K, lda, TC_BM, &ma));
    TW_CHECK(get_map(ctx, W, N, K, ldw, BN, &mw));
    mc = ma;
    if (BN >= 128) {        // TMA epilogue: output map (bf16 for the store / GELU epilogues, f32 otherwise)
        const int esize = (epi.mode == EPI_STORE || epi.mode == EPI_GELU) ? 2 : 4;
        if ((epi.ldc * esize) % 16 || (reinterpret_cast<uintptr_t>(epi.C) & 15)) {
            ctx->set_error(TW_E_UNSUPPORTED, "gemm_tc: output must be 16-byte aligned with a 16-byte multiple row pitch");
            return TW_E_UNSUPPORTED;
        }
        TW_CHECK(get_map(ctx, epi.C, M, N, epi.ldc, TC_BM, &mc, esize));
    }
    int tiles = ceil_div(M, TC_BM) * ceil_div(N, BN);
    int ksplit = 1;
    if (BN <= 64 && epi.mode == EPI_RESID) {
        // skinny residual GEMM: split K so that ~one wave of CTAs each streams a short K range
        const int kb = ceil_div(K, TC_BK);
        ksplit = ctx->sm_count / tiles;
        if (ksplit > kb / 4) ksplit = kb / 4;
        if (ksplit < 1) ksplit = 1;
        ksplit = ceil_div(kb, ceil_div(kb, ksplit));        // every split gets at least one K block
    }
    tiles *= ksplit;
    const int grid = tiles < ctx->sm_count ? tiles : ctx->sm_count;
    if (BN == 256)
        TW_CUDA_OK(ctx, launch_k(gemm_tc_kernel<256>, dim3(grid), dim3(TC_THREADS), TcCfg<256>::SMEM_BYTES, st, ma, mw, mc, M, N, K, ksplit, epi));
    else if (BN == 64)
        TW_CUDA_OK(ctx, launch_k(gemm_tc_kernel<64>, dim3(grid), dim3(TC_THREADS), TcCfg<64>::SMEM_BYTES, st, ma, mw, mc, M, N, K, ksplit, epi));
    else if (BN == 32)
        TW_CUDA_OK(ctx, launch_k(gemm_tc_kernel<32>, dim3(grid), dim3(TC_THREADS), TcCfg<32>::SMEM_BYTES, st, ma, mw, mc, M, N, K, ksplit, epi));
    else
        TW_CUDA_OK(ctx, launch_k(gemm_tc_kernel<128>, dim3(grid), dim3(TC_THREADS), TcCfg<128>::SMEM_BYTES, st, ma, mw, mc, M, N, K, ksplit, epi));
    TW_CUDA_OK(ctx, cudaGetLastError());
    return TW_OK;
}

}  // namespace tw
