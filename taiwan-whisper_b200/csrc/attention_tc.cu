// Encoder self-attention on 5th-gen tensor cores (sm_100a): softmax(Q K^T) V per (clip, head), non-causal,
// head_dim 64, bf16 operands, fp32 accumulate / softmax.  (HF WhisperAttention.forward,
// modeling_whisper.py:284-358; q is pre-scaled, no mask.)
//
// One CTA = one 128-query block of one (clip, head); two CTAs are co-resident per SM (256 TMEM columns,
// ~112 KB shared memory each) so one CTA's softmax overlaps the other's MMAs.
//   warp 4      TMA producer: Q tile once, then K / V tiles (128 keys x 64 dims, 128-byte swizzle) through
//               2-deep rings, straight out of the fused QKV activation matrix [B*S, 3d]
//   warp 5      MMA issuer: S = Q K^T (tcgen05.mma M128 N128 K16 x4, both operands K-major) into TMEM columns
//               [0,128); O_tile = P V (M128 N64 K16 x8, A = P from shared memory, B = V tile used MN-major)
//               into TMEM columns [128,192)
//   warps 0-3   softmax: thread = query row.  tcgen05.ld the scores, running max / sum in registers (no
//               shuffles: a thread owns its row), P = exp2(s*log2e - m*log2e) rounded to bf16 and written to
//               shared memory in the K-major 128-byte-swizzle layout the MMA expects, then O_reg = O_reg*scale
//               + O_tile read back from TMEM.  Final O / l stored as bf16 (128 contiguous bytes per row).
// Keys beyond the clip length (the 1536-padded tail, which TMA fills with the next clip's rows or zeros)
// are masked to -inf before the softmax.
#include <cuda.h>
#include <stdlib.h>

#include "kernels.cuh"

namespace tw {

constexpr int FA_BQ = 128, FA_BK = 128, FA_D = 64;
constexpr int FA_THREADS = 192;
constexpr int FA_TILE_BYTES = 128 * 64 * 2;      // one 128 x 64 bf16 tile = 16 KB
constexpr int FA_SMEM = 1024 + FA_TILE_BYTES * (1 + 2 + 1 + 2) + 256;   // Q, K x2, V x1, P (2 atoms): 2 CTAs / SM
constexpr int FA_TMEM_COLS = 256;

__device__ __forceinline__ uint32_t fa_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void fa_mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void fa_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void fa_mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void fa_mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "FA_WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra FA_WAIT_DONE;\n"
        "bra FA_WAIT_LOOP;\n"
        "FA_WAIT_DONE:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void fa_tma_load_2d(const CUtensorMap* map, uint32_t bar, uint32_t dst, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void fa_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void fa_mma(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void fa_tmem_ld32(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void fa_tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float fa_ex2(float x) {           // single MUFU.EX2 (exp2f adds range fix-ups we do not need)
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ void fa_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fa_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// shared-memory matrix descriptors, 128-byte swizzle, 8-row groups 1024 B apart (SBO); version 1 (sm_100)
__device__ __forceinline__ uint64_t fa_desc(uint32_t smem_addr) {
    return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// instruction descriptor: f32 accumulate, bf16 x bf16; b_mn_major selects an MN-major B operand (bit 16)
__host__ __device__ constexpr uint32_t fa_idesc(int M, int N, int b_mn_major) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)b_mn_major << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// VAR 1 (default): the TMEM load of the next 32 score columns is issued before the math of the current ones (+6 % on B200);
// VAR 0 (TWB200_FA_VARIANT=0): one load at a time.  Measured and dropped (profiles/r01_encoder_attention_ncu.md): a lazy
// running max (single pass over the scores), a share of the exponentials as an FMA-pipe polynomial, back-off in the
// producer / MMA-issuer waits — none of them faster.
template <int VAR>
__global__ void __launch_bounds__(FA_THREADS, 2)
encoder_attention_tc_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_kv,
                            __nv_bfloat16* __restrict__ out, int S, int Sk, int H, int q_col0, int k_col0, int v_col0, int causal) {
    extern __shared__ unsigned char fa_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(fa_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char* sQ = smem;
    unsigned char* sK = smem + FA_TILE_BYTES;           // 2 stages
    unsigned char* sV = smem + 3 * FA_TILE_BYTES;       // 1 stage (V(j) is only needed after softmax(j))
    unsigned char* sP = smem + 4 * FA_TILE_BYTES;       // 2 swizzle atoms (keys 0-63 | 64-127), 128 rows each
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 6 * FA_TILE_BYTES);
    uint64_t* q_full = bars;          // 1
    uint64_t* k_full = bars + 1;      // 2
    uint64_t* k_empty = bars + 3;     // 2
    uint64_t* v_full = bars + 5;      // 2
    uint64_t* v_empty = bars + 7;     // 2
    uint64_t* s_full = bars + 9;      // 1  QK^T done
    uint64_t* p_full = bars + 10;     // 1  P written (and S consumed)
    uint64_t* o_full = bars + 11;     // 1  PV done
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 12);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qb = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
    const int d = H * FA_D;
    const int q0 = qb * FA_BQ;
    // S queries per clip (rows of map_q), Sk keys / values per clip (rows of map_kv); causal: key index <= query index, so
    // query block qb only visits key tiles 0..qb (FA_BQ == FA_BK)
    const int n_tiles = causal ? min((Sk + FA_BK - 1) / FA_BK, qb + 1) : (Sk + FA_BK - 1) / FA_BK;
    const int row_base = b * S;                       // first query row of this clip
    const int kv_base = b * Sk;                       // first key / value row of this clip

    if (threadIdx.x == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_q) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_kv) : "memory");
        fa_mbar_init(fa_smem_u32(q_full), 1);
        for (int i = 0; i < 2; ++i) {
            fa_mbar_init(fa_smem_u32(&k_full[i]), 1);
            fa_mbar_init(fa_smem_u32(&k_empty[i]), 1);
            fa_mbar_init(fa_smem_u32(&v_full[i]), 1);
            fa_mbar_init(fa_smem_u32(&v_empty[i]), 1);
        }
        fa_mbar_init(fa_smem_u32(s_full), 1);
        fa_mbar_init(fa_smem_u32(p_full), 4);          // one arrive per softmax warp
        fa_mbar_init(fa_smem_u32(o_full), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 5) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(fa_smem_u32(tmem_ptr)), "r"(FA_TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fa_fence_before();
    __syncthreads();
    fa_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    const uint32_t tmem_S = tmem_base;            // columns [0,128)
    const uint32_t tmem_O = tmem_base + 128;      // columns [128,192)

    if (warp == 4) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            fa_mbar_expect_tx(fa_smem_u32(q_full), FA_TILE_BYTES);
            fa_tma_load_2d(&map_q, fa_smem_u32(q_full), fa_smem_u32(sQ), q_col0 + h * FA_D, row_base + q0);
            for (int j = 0; j < n_tiles; ++j) {
                const int st = j & 1;
                const uint32_t ph = (j >> 1) & 1;
                fa_mbar_wait(fa_smem_u32(&k_empty[st]), ph ^ 1);
                fa_mbar_expect_tx(fa_smem_u32(&k_full[st]), FA_TILE_BYTES);
                fa_tma_load_2d(&map_kv, fa_smem_u32(&k_full[st]), fa_smem_u32(sK + st * FA_TILE_BYTES), k_col0 + h * FA_D,
                               kv_base + j * FA_BK);
                fa_mbar_wait(fa_smem_u32(&v_empty[0]), (j & 1) ^ 1);
                fa_mbar_expect_tx(fa_smem_u32(&v_full[0]), FA_TILE_BYTES);
                fa_tma_load_2d(&map_kv, fa_smem_u32(&v_full[0]), fa_smem_u32(sV), v_col0 + h * FA_D, kv_base + j * FA_BK);
            }
        }
    } else if (warp == 5) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            constexpr uint32_t idesc_qk = fa_idesc(128, 128, 0);
            constexpr uint32_t idesc_pv = fa_idesc(128, 64, 1);
            const uint64_t q_desc = fa_desc(fa_smem_u32(sQ));
            fa_mbar_wait(fa_smem_u32(q_full), 0);
            for (int j = 0; j < n_tiles; ++j) {
                const int st = j & 1;
                const uint32_t ph = (j >> 1) & 1;
                // S(j) = Q K(j)^T   (S columns are free: the softmax warps arrived on p_full(j-1))
                fa_mbar_wait(fa_smem_u32(&k_full[st]), ph);
                if (j > 0) fa_mbar_wait(fa_smem_u32(p_full), (j - 1) & 1);
                fa_fence_after();
                const uint64_t k_desc = fa_desc(fa_smem_u32(sK + st * FA_TILE_BYTES));
#pragma unroll
                for (int k = 0; k < FA_D / 16; ++k) fa_mma(tmem_S, q_desc + 2 * k, k_desc + 2 * k, idesc_qk, k > 0 ? 1u : 0u);
                fa_commit(fa_smem_u32(&k_empty[st]));
                fa_commit(fa_smem_u32(s_full));
                if (j > 0) {
                    // O_tile(j-1) = P(j-1) V(j-1): P is in shared memory since p_full(j-1)
                    fa_mbar_wait(fa_smem_u32(&v_full[0]), (j - 1) & 1);
                    fa_fence_after();
                    const uint32_t v_addr = fa_smem_u32(sV);
#pragma unroll
                    for (int k = 0; k < FA_BK / 16; ++k) {
                        // A: P atom (k/4), 32-byte steps inside the 128-byte swizzle row; B: 16 key rows = 2048 bytes
                        const uint64_t p_desc = fa_desc(fa_smem_u32(sP + (k >> 2) * FA_TILE_BYTES)) + 2 * (k & 3);
                        const uint64_t v_desc = fa_desc(v_addr + k * 2048);
                        fa_mma(tmem_O, p_desc, v_desc, idesc_pv, k > 0 ? 1u : 0u);
                    }
                    fa_commit(fa_smem_u32(&v_empty[0]));
                    fa_commit(fa_smem_u32(o_full));
                }
            }
            {   // last P V
                const int j = n_tiles;
                fa_mbar_wait(fa_smem_u32(p_full), (j - 1) & 1);
                fa_mbar_wait(fa_smem_u32(&v_full[0]), (j - 1) & 1);
                fa_fence_after();
                const uint32_t v_addr = fa_smem_u32(sV);
#pragma unroll
                for (int k = 0; k < FA_BK / 16; ++k) {
                    const uint64_t p_desc = fa_desc(fa_smem_u32(sP + (k >> 2) * FA_TILE_BYTES)) + 2 * (k & 3);
                    const uint64_t v_desc = fa_desc(v_addr + k * 2048);
                    fa_mma(tmem_O, p_desc, v_desc, idesc_pv, k > 0 ? 1u : 0u);
                }
                fa_commit(fa_smem_u32(&v_empty[0]));
                fa_commit(fa_smem_u32(o_full));
            }
        }
    } else {
        // ===================== softmax warps 0..3: thread = query row =====================
        const int r = warp * 32 + lane;                 // row inside the tile == TMEM lane
        const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
        const float LOG2E = 1.4426950408889634f;
        float o_acc[FA_D];
#pragma unroll
        for (int i = 0; i < FA_D; ++i) o_acc[i] = 0.0f;
        float m_run = -INFINITY, l_run = 0.0f, scale_prev = 1.0f;
        // ---- the two passes over the 128 score columns of this row (TMEM lane), 32 columns at a time
        // pass A: raw row max (masked columns excluded)
        auto max_pass = [&](int valid, bool full_tile) -> float {
            float mx = -INFINITY;
            auto fold = [&](const uint32_t* v, int c) {
                if (full_tile) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(v[i]));
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i) mx = fmaxf(mx, (c + i < valid) ? __uint_as_float(v[i]) : -INFINITY);
                }
            };
            if (VAR & 1) {
                uint32_t va[32], vb[32];
                fa_tmem_ld32(tmem_S + lane_off, va);
                fa_tmem_wait_ld(); fa_tmem_ld32(tmem_S + lane_off + 32, vb); fold(va, 0);
                fa_tmem_wait_ld(); fa_tmem_ld32(tmem_S + lane_off + 64, va); fold(vb, 32);
                fa_tmem_wait_ld(); fa_tmem_ld32(tmem_S + lane_off + 96, vb); fold(va, 64);
                fa_tmem_wait_ld(); fold(vb, 96);
            } else {
#pragma unroll 1
                for (int c = 0; c < FA_BK; c += 32) {
                    uint32_t v[32];
                    fa_tmem_ld32(tmem_S + lane_off + c, v);
                    fa_tmem_wait_ld();
                    fold(v, c);
                }
            }
            return mx;
        };
        // pass B: P = exp2(s*log2e - m*log2e) -> bf16 -> shared memory (K-major, 128-byte swizzle); returns the row sum
        auto exp_pass = [&](float mneg, int valid, bool full_tile) -> float {
            float lsum0 = 0.0f, lsum1 = 0.0f;
            auto chunk = [&](const uint32_t* v, int c) {
                uint32_t pk[16];
                if (full_tile) {
#pragma unroll
                    for (int i = 0; i < 32; i += 2) {
                        const float s0 = __uint_as_float(v[i]), s1 = __uint_as_float(v[i + 1]);
                        const float p0 = fa_ex2(fmaf(s0, LOG2E, mneg));
                        const float p1 = fa_ex2(fmaf(s1, LOG2E, mneg));
                        lsum0 += p0;
                        lsum1 += p1;
                        __nv_bfloat162 hb = __floats2bfloat162_rn(p0, p1);
                        pk[i >> 1] = *reinterpret_cast<uint32_t*>(&hb);
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 32; i += 2) {
                        const bool ok0 = c + i < valid, ok1 = c + i + 1 < valid;
                        const float s0 = __uint_as_float(v[i]), s1 = __uint_as_float(v[i + 1]);
                        const float p0 = ok0 ? fa_ex2(fmaf(s0, LOG2E, mneg)) : 0.0f;
                        const float p1 = ok1 ? fa_ex2(fmaf(s1, LOG2E, mneg)) : 0.0f;
                        lsum0 += p0;
                        lsum1 += p1;
                        __nv_bfloat162 hb = __floats2bfloat162_rn(p0, p1);
                        pk[i >> 1] = *reinterpret_cast<uint32_t*>(&hb);
                    }
                }
                // 32 keys = 64 bytes = 4 chunks of 16 B; chunk index inside the 128-byte row: ((c % 64) / 8) + q
                unsigned char* prow = sP + (c >> 6) * FA_TILE_BYTES + r * 128;
#pragma unroll
                for (int qd = 0; qd < 4; ++qd) {
                    const int ch = ((c & 63) >> 3) + qd;
                    *reinterpret_cast<uint4*>(prow + ((ch ^ (r & 7)) << 4)) =
                        make_uint4(pk[4 * qd], pk[4 * qd + 1], pk[4 * qd + 2], pk[4 * qd + 3]);
                }
            };
            if (VAR & 1) {
                uint32_t va[32], vb[32];
                fa_tmem_ld32(tmem_S + lane_off, va);
                fa_tmem_wait_ld(); fa_tmem_ld32(tmem_S + lane_off + 32, vb); chunk(va, 0);
                fa_tmem_wait_ld(); fa_tmem_ld32(tmem_S + lane_off + 64, va); chunk(vb, 32);
                fa_tmem_wait_ld(); fa_tmem_ld32(tmem_S + lane_off + 96, vb); chunk(va, 64);
                fa_tmem_wait_ld(); chunk(vb, 96);
            } else {
#pragma unroll 1
                for (int c = 0; c < FA_BK; c += 32) {
                    uint32_t v[32];
                    fa_tmem_ld32(tmem_S + lane_off + c, v);
                    fa_tmem_wait_ld();
                    chunk(v, c);
                }
            }
            return lsum0 + lsum1;
        };

        for (int j = 0; j < n_tiles; ++j) {
            fa_mbar_wait(fa_smem_u32(s_full), j & 1);
            fa_fence_after();
            // keys >= valid are padding (the tail of the clip) or, under the causal mask, later than this row's query
            const int valid = causal ? min(Sk - j * FA_BK, q0 + r - j * FA_BK + 1) : Sk - j * FA_BK;
            // only the last key tile of a clip / the diagonal tile is masked (warp-uniform)
            const bool full_tile = (Sk - j * FA_BK >= FA_BK) && !(causal && j == qb);
            const float m_new = fmaxf(m_run, max_pass(valid, full_tile));
            const float scale = fa_ex2((m_run - m_new) * LOG2E);    // 0 on the first tile (m_run = -inf)
            // the previous P V must have finished reading P before it is overwritten, and O_tile(j-1) is folded in
            if (j > 0) {
                fa_mbar_wait(fa_smem_u32(o_full), (j - 1) & 1);
                fa_fence_after();
#pragma unroll
                for (int c = 0; c < FA_D; c += 32) {
                    uint32_t v[32];
                    fa_tmem_ld32(tmem_O + lane_off + c, v);
                    fa_tmem_wait_ld();
#pragma unroll
                    for (int i = 0; i < 32; ++i) o_acc[c + i] = fmaf(o_acc[c + i], scale_prev, __uint_as_float(v[i]));
                }
            }
            const float lsum = exp_pass(-m_new * LOG2E, valid, full_tile);
            l_run = l_run * scale + lsum;
            m_run = m_new;
            scale_prev = scale;
            // P visible to the tensor core (async proxy); S columns free again
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            fa_fence_before();
            __syncwarp();
            if (lane == 0) fa_mbar_arrive(fa_smem_u32(p_full));
        }
        // last O tile.  NB: scale_prev belongs to the tile whose P V is now finishing: O = O*scale + O_tile
        fa_mbar_wait(fa_smem_u32(o_full), (n_tiles - 1) & 1);
        fa_fence_after();
#pragma unroll
        for (int c = 0; c < FA_D; c += 32) {
            uint32_t v[32];
            fa_tmem_ld32(tmem_O + lane_off + c, v);
            fa_tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; ++i) o_acc[c + i] = fmaf(o_acc[c + i], scale_prev, __uint_as_float(v[i]));
        }
        const int q = q0 + r;
        if (q < S) {
            const float inv = 1.0f / l_run;
            __nv_bfloat16* orow = out + ((int64_t)(row_base + q)) * d + h * FA_D;
#pragma unroll
            for (int i = 0; i < FA_D; i += 8) {
                uint4 pk;
                __nv_bfloat162 h0 = __floats2bfloat162_rn(o_acc[i] * inv, o_acc[i + 1] * inv);
                __nv_bfloat162 h1 = __floats2bfloat162_rn(o_acc[i + 2] * inv, o_acc[i + 3] * inv);
                __nv_bfloat162 h2 = __floats2bfloat162_rn(o_acc[i + 4] * inv, o_acc[i + 5] * inv);
                __nv_bfloat162 h3 = __floats2bfloat162_rn(o_acc[i + 6] * inv, o_acc[i + 7] * inv);
                pk.x = *reinterpret_cast<uint32_t*>(&h0); pk.y = *reinterpret_cast<uint32_t*>(&h1);
                pk.z = *reinterpret_cast<uint32_t*>(&h2); pk.w = *reinterpret_cast<uint32_t*>(&h3);
                *reinterpret_cast<uint4*>(orow + i) = pk;
            }
        }
        fa_fence_before();
    }
    __syncthreads();
    if (warp == 5) {
        fa_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(FA_TMEM_COLS) : "memory");
    }
}

// ---- host ----------------------------------------------------------------------------------
typedef CUresult (*FaEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static FaEncodeTiledFn g_fa_encode = nullptr;

static int fa_map(tw_ctx* ctx, const __nv_bfloat16* ptr, int rows, int64_t ld, CUtensorMap* out) {
    struct Entry { const void* ptr; int rows; int64_t ld; CUtensorMap map; };
    static Entry cache[8];
    static int n_cached = 0, next = 0;
    for (int i = 0; i < n_cached; ++i)
        if (cache[i].ptr == ptr && cache[i].rows == rows && cache[i].ld == ld) { *out = cache[i].map; return TW_OK; }
    if ((reinterpret_cast<uintptr_t>(ptr) & 15) || (ld % 8)) {
        ctx->set_error(TW_E_UNSUPPORTED, "attention_tc: operands must be 16-byte aligned with a row pitch that is a multiple of 8 elements");
        return TW_E_UNSUPPORTED;
    }
    CUtensorMap m;
    const cuuint64_t dims[2] = {(cuuint64_t)ld, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
    const cuuint32_t box[2] = {64, 128};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = g_fa_encode(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<__nv_bfloat16*>(ptr), dims, strides, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        ctx->set_error(TW_E_CUDA, "attention_tc: cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
        return TW_E_CUDA;
    }
    Entry& e = cache[next];
    e.ptr = ptr; e.rows = rows; e.ld = ld; e.map = m;
    next = (next + 1) % 8;
    if (n_cached < 8) ++n_cached;
    *out = m;
    return TW_OK;
}

// General form: Sq query rows per clip in q (pitch q_ld, head h at column q_col0 + 64 h) against Sk key / value rows per clip
// in kv (pitch kv_ld, K at k_col0 + 64 h, V at v_col0 + 64 h); causal needs Sq == Sk.  out [B*Sq, 64 H].
int attention_tc(tw_ctx* ctx, const __nv_bfloat16* q, int64_t q_ld, int q_col0, const __nv_bfloat16* kv, int64_t kv_ld, int k_col0,
                 int v_col0, __nv_bfloat16* out, int B, int Sq, int Sk, int H, bool causal, cudaStream_t st) {
    if (!g_fa_encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
        if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) {
            ctx->set_error(TW_E_CUDA, "cuTensorMapEncodeTiled not available from the driver");
            return TW_E_CUDA;
        }
        g_fa_encode = reinterpret_cast<FaEncodeTiledFn>(fn);
        TW_CUDA_OK(ctx, cudaFuncSetAttribute(encoder_attention_tc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, FA_SMEM));
        TW_CUDA_OK(ctx, cudaFuncSetAttribute(encoder_attention_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, FA_SMEM));
    }
    if (causal && Sq != Sk) {
        ctx->set_error(TW_E_INVALID, "attention_tc: the causal mask needs as many queries as keys");
        return TW_E_INVALID;
    }
    CUtensorMap mq, mkv;
    TW_CHECK(fa_map(ctx, q, B * Sq, q_ld, &mq));
    TW_CHECK(fa_map(ctx, kv, B * Sk, kv_ld, &mkv));
    dim3 grid(ceil_div(Sq, FA_BQ), H, B);
    static const int variant = getenv("TWB200_FA_VARIANT") ? atoi(getenv("TWB200_FA_VARIANT")) : 1;
    if (variant == 0)
        encoder_attention_tc_kernel<0><<<grid, FA_THREADS, FA_SMEM, st>>>(mq, mkv, out, Sq, Sk, H, q_col0, k_col0, v_col0, causal ? 1 : 0);
    else
        encoder_attention_tc_kernel<1><<<grid, FA_THREADS, FA_SMEM, st>>>(mq, mkv, out, Sq, Sk, H, q_col0, k_col0, v_col0, causal ? 1 : 0);
    TW_CUDA_OK(ctx, cudaGetLastError());
    return TW_OK;
}

int encoder_attention_tc(tw_ctx* ctx, const __nv_bfloat16* qkv, __nv_bfloat16* out, int B, int S, int H, cudaStream_t st) {
    const int d = H * FA_D;
    return attention_tc(ctx, qkv, 3 * d, 0, qkv, 3 * d, d, 2 * d, out, B, S, S, H, false, st);
}

}  // namespace tw
