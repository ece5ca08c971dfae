// Encoder self-attention on 5th-gen tensor cores (sm_100a): softmax(Q K^T) V per (clip, head), non-causal,
// head_dim 64, bf16 operands, fp32 accumulate / softmax.  (HF WhisperAttention.forward,
// modeling_whisper.py:284-358; q is pre-scaled, no mask.)
//
// One CTA = one 128-query block of one (clip, head); two CTAs are co-resident per SM (208 of 256 TMEM columns,
// ~90 KB shared memory each).
//   warp 4      TMA producer: Q tile once, then K / V tiles (64 keys x 64 dims, 128-byte swizzle) through 4-deep rings,
//               straight out of the fused QKV activation matrix [B*S, 3d]
//   warp 5      MMA issuer (warp-uniform loop, one elected lane): S = Q K^T (M128 N64 K16 x4) into one of two TMEM score
//               buffers; O += P V and L += P 1 with P read from TMEM (M128 N64 / N16, K16 x4, V used MN-major)
//   warps 0-3   softmax: thread = query row; scores from TMEM to registers once, P back to TMEM — see the comment on the kernel
// Keys beyond the clip length (the padded tail, which TMA fills with the next clip's rows or zeros) and, under the causal
// mask, keys later than the query are given zero probability.
#include <cuda.h>
#include <stdlib.h>

#include "kernels.cuh"

namespace tw {

constexpr int FA_BQ = 128, FA_D = 64;
constexpr int FA_THREADS = 192;
constexpr int FA_TILE_BYTES = 128 * 64 * 2;      // one 128 x 64 bf16 tile = 16 KB
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
// one elected lane of a converged warp (see tc_ptx.cuh: the single-thread roles run warp-uniform and issue under this predicate)
__device__ __forceinline__ uint32_t fa_elect() {
    uint32_t pred = 0;
    asm volatile("{\n.reg .pred P1;\nelect.sync _|P1, 0xffffffff;\nselp.u32 %0, 1, 0, P1;\n}\n" : "=r"(pred));
    return pred;
}
// p = (col < valid) ? p : 0 as an opaque (volatile) select: it keeps the masked key-tile code path from being merged with
// the unmasked one — the compiler otherwise if-converts both into ONE path that runs a compare + select per score on every
// tile (25 M ISETP + 25 M FSEL per launch in profiles/r02_encoder_attention_ncu.md), although only the last key tile of
// a clip (or the causal diagonal) is masked
__device__ __forceinline__ float fa_mask0(float p, int col, int valid) {
    asm volatile("{\n.reg .pred q;\nsetp.lt.s32 q, %1, %2;\nselp.f32 %0, %0, 0f00000000, q;\n}\n" : "+f"(p) : "r"(col), "r"(valid));
    return p;
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

// ---- tensor-memory helpers of the kernel below: P V with the A operand in TMEM, tcgen05.st, 3-input max
constexpr float FA2_RESCALE_LOG2 = 8.0f;      // the reference max of a row moves only when the tile max exceeds it by 2^8

__device__ __forceinline__ void fa_mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
        "}\n" ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void fa_tmem_st16(uint32_t taddr, const uint32_t* r) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
          "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
__device__ __forceinline__ void fa_tmem_st32(uint32_t taddr, const uint32_t* r) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
          "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]),
          "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]),
          "r"(r[30]), "r"(r[31]) : "memory");
}
__device__ __forceinline__ void fa_tmem_ld1(uint32_t taddr, uint32_t& r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void fa_tmem_st1(uint32_t taddr, uint32_t r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr), "r"(r) : "memory");
}
__device__ __forceinline__ void fa_tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float fa_max3(float a, float b, float c) {
    float d;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}

// =====================================================================================================================
// The kernel.  Round 1 shipped a 128-key-tile kernel with P staged in shared memory and O folded into registers every tile
// (444 TFLOP/s); its profile showed neither MUFU, tensor nor memory throughput saturated.  Round 2 found the cause in two
// steps (profiles/r02_encoder_attention_ncu.md): (1) every tcgen05.mma was issued from `if (lane == 0)` code, which the
// compiler wraps in a divergence loop (ELECT / BRA.U.ANY) — tens of cycles per MMA, so the 12-20 MMAs of a key tile took
// longer to ISSUE than to execute and the softmax warps waited ~44 % of their time for S; (2) with warp-uniform issue the
// softmax of tile j can be overlapped with Q K^T of tile j+1 by double-buffering 64-key score tiles in tensor memory.
//   * P (bf16) is written back to TENSOR MEMORY over the score columns it was computed from (tcgen05.st) and P V runs with
//     the A operand in TMEM: no shared-memory P tile, no swizzled 16-byte stores, no proxy fence;
//   * O accumulates in TMEM across key tiles; the exponentials use a per-row REFERENCE max that only moves when the tile max
//     exceeds it by more than 2^8 (then O and the row sum are rescaled in TMEM, warp-uniformly skipped otherwise) — exact,
//     because numerator and denominator carry the same reference;
//   * the row sum comes from the tensor core: a second MMA of P against a tile of ones accumulates sum_k P[r,k] into a TMEM
//     column, from the same bf16-rounded P the numerator uses (no FADD per score); the row max uses the 3-input FMNMX;
//   * the masked key tile (clip tail / causal diagonal) has its own code path (fa_mask0), the others run select-free.
// TMEM: S0 [0,64) | S1 [64,128) | O [128,192) | L [192,208) = 208 columns, 2 CTAs / SM.
//   MMA issue order:  QK(0) QK(1) | PV(0) QK(2) | PV(1) QK(3) | ...     (P(j) overwrites the first 32 columns of S(j&1); QK(j+2)
//   is issued after PV(j), and the tensor pipe executes in order).  O / L are rescaled (rarely) only after pv_done(j-1).
// Measured (B200, 20 heads x 1500 keys, batch 8): 591 TFLOP/s; also tried and dropped: 128-key tiles without the double
// buffer (518), eight softmax warps per CTA sharing rows through shared memory (555), every 4th exponential as an FMA-pipe
// cubic instead of MUFU.EX2 (631 alone and test-green, but the pipelined batch loop hung with it on the 24-SM partition:
// profiles/r02_encoder_attention_ncu.md, addendum 2).
constexpr int FA3_BK = 64;
constexpr int FA3_KV_BYTES = FA3_BK * FA_D * 2;       // 8 KB
constexpr int FA3_STAGES = 4;
constexpr int FA3_SMEM = 1024 + FA_TILE_BYTES + FA3_KV_BYTES * (2 * FA3_STAGES + 1) + 512;    // Q, K x4, V x4, ones

__global__ void __launch_bounds__(FA_THREADS, 2)
encoder_attention_tc_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_kv,
                             __nv_bfloat16* __restrict__ out, int S, int Sk, int H, int q_col0, int k_col0, int v_col0, int causal) {
    extern __shared__ unsigned char fa_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(fa_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char* sQ = smem;
    unsigned char* sK = smem + FA_TILE_BYTES;
    unsigned char* sV = sK + FA3_STAGES * FA3_KV_BYTES;
    unsigned char* sOne = sV + FA3_STAGES * FA3_KV_BYTES;       // 64 x 64 bf16 ones
    uint64_t* bars = reinterpret_cast<uint64_t*>(sOne + FA3_KV_BYTES);
    uint64_t* q_full = bars;                       // 1
    uint64_t* k_full = bars + 1;                   // FA3_STAGES
    uint64_t* k_empty = k_full + FA3_STAGES;
    uint64_t* v_full = k_empty + FA3_STAGES;
    uint64_t* v_empty = v_full + FA3_STAGES;
    uint64_t* s_full = v_empty + FA3_STAGES;       // 2  QK^T into score buffer b done
    uint64_t* p_full = s_full + 2;                 // 2  P written over score buffer b
    uint64_t* pv_done = p_full + 2;                // 1  one phase per key tile: P V (j) has completed
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(pv_done + 1);

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
    const int qb = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
    const int d = H * FA_D;
    const int q0 = qb * FA_BQ;
    const int tiles_all = (Sk + FA3_BK - 1) / FA3_BK;
    const int n_tiles = causal ? min(tiles_all, 2 * qb + 2) : tiles_all;      // causal: tiles up to the block's last query
    const int row_base = b * S;
    const int kv_base = b * Sk;

    if (threadIdx.x == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_q) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_kv) : "memory");
        fa_mbar_init(fa_smem_u32(q_full), 1);
        for (int i = 0; i < FA3_STAGES; ++i) {
            fa_mbar_init(fa_smem_u32(&k_full[i]), 1);
            fa_mbar_init(fa_smem_u32(&k_empty[i]), 1);
            fa_mbar_init(fa_smem_u32(&v_full[i]), 1);
            fa_mbar_init(fa_smem_u32(&v_empty[i]), 1);
        }
        for (int i = 0; i < 2; ++i) {
            fa_mbar_init(fa_smem_u32(&s_full[i]), 1);
            fa_mbar_init(fa_smem_u32(&p_full[i]), 4);      // one arrive per softmax warp
        }
        fa_mbar_init(fa_smem_u32(pv_done), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 5) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(fa_smem_u32(tmem_ptr)), "r"(FA_TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (warp < 4) {       // ones tile (bf16 1.0 = 0x3F80)
        uint4* p1 = reinterpret_cast<uint4*>(sOne);
        for (int i = threadIdx.x; i < FA3_KV_BYTES / 16; i += 128) p1[i] = make_uint4(0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    fa_fence_before();
    __syncthreads();
    fa_fence_after();
    // programmatic dependent launch (encoder chain): barrier / TMEM / ones-tile setup above overlaps the previous kernel's tail
    pdl_trigger();
    pdl_wait();
    const uint32_t tmem_base = *tmem_ptr;
    const uint32_t tmem_S = tmem_base;            // two score buffers of 64 columns; P (bf16 pairs) over the first 32 of each
    const uint32_t tmem_O = tmem_base + 128;
    const uint32_t tmem_L = tmem_base + 192;

    if (warp == 4) {
        // ===================== TMA producer =====================
        if (fa_elect()) {
            fa_mbar_expect_tx(fa_smem_u32(q_full), FA_TILE_BYTES);
            fa_tma_load_2d(&map_q, fa_smem_u32(q_full), fa_smem_u32(sQ), q_col0 + h * FA_D, row_base + q0);
        }
        __syncwarp();
        for (int j = 0; j < n_tiles; ++j) {
            const int st = j % FA3_STAGES;
            const uint32_t ph = (j / FA3_STAGES) & 1;
            fa_mbar_wait(fa_smem_u32(&k_empty[st]), ph ^ 1);
            if (fa_elect()) {
                fa_mbar_expect_tx(fa_smem_u32(&k_full[st]), FA3_KV_BYTES);
                fa_tma_load_2d(&map_kv, fa_smem_u32(&k_full[st]), fa_smem_u32(sK + st * FA3_KV_BYTES), k_col0 + h * FA_D,
                               kv_base + j * FA3_BK);
            }
            __syncwarp();
            fa_mbar_wait(fa_smem_u32(&v_empty[st]), ph ^ 1);
            if (fa_elect()) {
                fa_mbar_expect_tx(fa_smem_u32(&v_full[st]), FA3_KV_BYTES);
                fa_tma_load_2d(&map_kv, fa_smem_u32(&v_full[st]), fa_smem_u32(sV + st * FA3_KV_BYTES), v_col0 + h * FA_D,
                               kv_base + j * FA3_BK);
            }
            __syncwarp();
        }
    } else if (warp == 5) {
        // ===================== MMA issuer =====================
        constexpr uint32_t idesc_qk = fa_idesc(128, FA3_BK, 0);
        constexpr uint32_t idesc_pv = fa_idesc(128, 64, 1);
        constexpr uint32_t idesc_l = fa_idesc(128, 16, 1);
        const uint32_t q_addr = fa_smem_u32(sQ);
        const uint32_t one_addr = fa_smem_u32(sOne);
        auto issue_qk = [&](int j) {
            const int st = j % FA3_STAGES;
            fa_mbar_wait(fa_smem_u32(&k_full[st]), (j / FA3_STAGES) & 1);
            fa_fence_after();
            if (fa_elect()) {
                const uint64_t q_desc = fa_desc(q_addr);
                const uint64_t k_desc = fa_desc(fa_smem_u32(sK + st * FA3_KV_BYTES));
#pragma unroll
                for (int k = 0; k < FA_D / 16; ++k)
                    fa_mma(tmem_S + (j & 1) * FA3_BK, q_desc + 2 * k, k_desc + 2 * k, idesc_qk, k > 0 ? 1u : 0u);
                fa_commit(fa_smem_u32(&k_empty[st]));
                fa_commit(fa_smem_u32(&s_full[j & 1]));
            }
            __syncwarp();
        };
        auto issue_pv = [&](int j) {
            const int st = j % FA3_STAGES;
            fa_mbar_wait(fa_smem_u32(&p_full[j & 1]), (j >> 1) & 1);
            fa_mbar_wait(fa_smem_u32(&v_full[st]), (j / FA3_STAGES) & 1);
            fa_fence_after();
            const uint32_t v_addr = fa_smem_u32(sV + st * FA3_KV_BYTES);
            const uint32_t p_tmem = tmem_S + (j & 1) * FA3_BK;
            if (fa_elect()) {
#pragma unroll
                for (int k = 0; k < FA3_BK / 16; ++k) {
                    fa_mma_ts(tmem_O, p_tmem + 8 * k, fa_desc(v_addr + k * 2048), idesc_pv, (j > 0 || k > 0) ? 1u : 0u);
                    fa_mma_ts(tmem_L, p_tmem + 8 * k, fa_desc(one_addr + k * 2048), idesc_l, (j > 0 || k > 0) ? 1u : 0u);
                }
                fa_commit(fa_smem_u32(&v_empty[st]));
                fa_commit(fa_smem_u32(pv_done));
            }
            __syncwarp();
        };
        fa_mbar_wait(fa_smem_u32(q_full), 0);
        issue_qk(0);
        if (n_tiles > 1) issue_qk(1);
        for (int j = 0; j < n_tiles; ++j) {
            issue_pv(j);
            if (j + 2 < n_tiles) {
                // Q K^T (j+2) overwrites the score buffer whose first 32 columns are P(j), the TMEM A operand of P V (j).  Issue
                // order alone does not protect it: consecutive tcgen05.mma overlap in the pipe, and nothing interlocks a D write
                // with an earlier instruction's A read from tensor memory — measured as rare wrong rows (3 of 180 runs of the
                // 20-head, 1500-key shape at 2 CTAs per SM; 0 of 450 with the wait).  A variant with the query tile in TMEM (Q K^T
                // as a TMEM-A instruction, +2 %) kept failing 8 of 450 and was dropped.  Wait for P V (j) to complete.
                fa_mbar_wait(fa_smem_u32(pv_done), j & 1);
                fa_fence_after();
                issue_qk(j + 2);
            }
        }
    } else {
        // ===================== softmax warps 0..3: thread = query row =====================
        const int r = warp * 32 + lane;
        const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
        const float LOG2E = 1.4426950408889634f;
        float m_ref = -INFINITY;
        for (int j = 0; j < n_tiles; ++j) {
            const uint32_t sbuf = tmem_S + lane_off + (j & 1) * FA3_BK;
            fa_mbar_wait(fa_smem_u32(&s_full[j & 1]), (j >> 1) & 1);
            fa_fence_after();
            const int valid = causal ? min(Sk - j * FA3_BK, q0 + r - j * FA3_BK + 1) : Sk - j * FA3_BK;
            const bool full_tile = (Sk - j * FA3_BK >= FA3_BK) && !(causal && j >= 2 * qb);
            // the 64 scores of this row stay in registers for both passes
            uint32_t va[32], vb[32];
            fa_tmem_ld32(sbuf, va);
            fa_tmem_ld32(sbuf + 32, vb);
            fa_tmem_wait_ld();
            float mx = -INFINITY;
            if (full_tile) {
                // four independent chains of 3-input max (a single chain is 32 dependent instructions per tile)
                float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll
                for (int i = 0; i < 16; i += 2) {
                    m0 = fa_max3(m0, __uint_as_float(va[i]), __uint_as_float(va[i + 1]));
                    m1 = fa_max3(m1, __uint_as_float(va[16 + i]), __uint_as_float(va[17 + i]));
                    m2 = fa_max3(m2, __uint_as_float(vb[i]), __uint_as_float(vb[i + 1]));
                    m3 = fa_max3(m3, __uint_as_float(vb[16 + i]), __uint_as_float(vb[17 + i]));
                }
                mx = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
            } else {
#pragma unroll
                for (int i = 0; i < 32; ++i) mx = fmaxf(mx, (i < valid) ? __uint_as_float(va[i]) : -INFINITY);
#pragma unroll
                for (int i = 0; i < 32; ++i) mx = fmaxf(mx, (32 + i < valid) ? __uint_as_float(vb[i]) : -INFINITY);
            }
            if (j == 0) {
                m_ref = mx;
            } else {
                const bool grow = (mx - m_ref) * LOG2E > FA2_RESCALE_LOG2;
                if (__any_sync(0xffffffffu, grow)) {
                    // O and L are being accumulated by P V (j-1): wait for it, then rescale this row in tensor memory
                    fa_mbar_wait(fa_smem_u32(pv_done), (j - 1) & 1);
                    fa_fence_after();
                    const float f = grow ? fa_ex2((m_ref - mx) * LOG2E) : 1.0f;
                    if (grow) m_ref = mx;
#pragma unroll
                    for (int c = 0; c < FA_D; c += 32) {
                        uint32_t v[32];
                        fa_tmem_ld32(tmem_O + lane_off + c, v);
                        fa_tmem_wait_ld();
#pragma unroll
                        for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * f);
                        fa_tmem_st32(tmem_O + lane_off + c, v);
                    }
                    uint32_t lv;
                    fa_tmem_ld1(tmem_L + lane_off, lv);
                    fa_tmem_wait_ld();
                    fa_tmem_st1(tmem_L + lane_off, __float_as_uint(__uint_as_float(lv) * f));
                    fa_tmem_wait_st();
                }
            }
            const float mneg = -m_ref * LOG2E;
            uint32_t pk[32];
            if (full_tile) {
#pragma unroll
                for (int i = 0; i < 32; i += 2) {
                    __nv_bfloat162 hb = __floats2bfloat162_rn(fa_ex2(fmaf(__uint_as_float(va[i]), LOG2E, mneg)),
                                                              fa_ex2(fmaf(__uint_as_float(va[i + 1]), LOG2E, mneg)));
                    pk[i >> 1] = *reinterpret_cast<uint32_t*>(&hb);
                }
#pragma unroll
                for (int i = 0; i < 32; i += 2) {
                    __nv_bfloat162 hb = __floats2bfloat162_rn(fa_ex2(fmaf(__uint_as_float(vb[i]), LOG2E, mneg)),
                                                              fa_ex2(fmaf(__uint_as_float(vb[i + 1]), LOG2E, mneg)));
                    pk[16 + (i >> 1)] = *reinterpret_cast<uint32_t*>(&hb);
                }
            } else {
#pragma unroll
                for (int i = 0; i < 32; i += 2) {
                    __nv_bfloat162 hb = __floats2bfloat162_rn(fa_mask0(fa_ex2(fmaf(__uint_as_float(va[i]), LOG2E, mneg)), i, valid),
                                                              fa_mask0(fa_ex2(fmaf(__uint_as_float(va[i + 1]), LOG2E, mneg)), i + 1, valid));
                    pk[i >> 1] = *reinterpret_cast<uint32_t*>(&hb);
                }
#pragma unroll
                for (int i = 0; i < 32; i += 2) {
                    __nv_bfloat162 hb = __floats2bfloat162_rn(fa_mask0(fa_ex2(fmaf(__uint_as_float(vb[i]), LOG2E, mneg)), 32 + i, valid),
                                                              fa_mask0(fa_ex2(fmaf(__uint_as_float(vb[i + 1]), LOG2E, mneg)), 33 + i, valid));
                    pk[16 + (i >> 1)] = *reinterpret_cast<uint32_t*>(&hb);
                }
            }
            fa_tmem_st32(sbuf, pk);              // P(j): 64 keys = 32 columns over the scores they came from
            fa_tmem_wait_st();
            fa_fence_before();
            __syncwarp();
            if (lane == 0) fa_mbar_arrive(fa_smem_u32(&p_full[j & 1]));
        }
        // ---- epilogue: O / L.  A parity wait only tells phases of equal parity apart if the barrier is at most one phase behind:
        // P V (n-2) may still be in flight here (nothing after tile n-2 waited for it), and then a wait for the parity of phase n-1
        // is satisfied by phase n-3 — O and L were read one or two tiles early in ~1 % of the launches at 2 CTAs per SM.  Wait for
        // phase n-2 first (phase n-3 is complete: Q K^T (n-1) was issued after it).
        if (n_tiles >= 2) fa_mbar_wait(fa_smem_u32(pv_done), (n_tiles - 2) & 1);
        fa_mbar_wait(fa_smem_u32(pv_done), (n_tiles - 1) & 1);
        fa_fence_after();
        uint32_t lv;
        fa_tmem_ld1(tmem_L + lane_off, lv);
        fa_tmem_wait_ld();
        const float inv = 1.0f / __uint_as_float(lv);
        const int q = q0 + r;
        __nv_bfloat16* orow = out + ((int64_t)(row_base + q)) * d + h * FA_D;
#pragma unroll
        for (int c = 0; c < FA_D; c += 32) {
            uint32_t v[32];
            fa_tmem_ld32(tmem_O + lane_off + c, v);
            fa_tmem_wait_ld();
            if (q < S) {
#pragma unroll
                for (int i = 0; i < 32; i += 8) {
                    uint4 pk4;
                    __nv_bfloat162 h0 = __floats2bfloat162_rn(__uint_as_float(v[i]) * inv, __uint_as_float(v[i + 1]) * inv);
                    __nv_bfloat162 h1 = __floats2bfloat162_rn(__uint_as_float(v[i + 2]) * inv, __uint_as_float(v[i + 3]) * inv);
                    __nv_bfloat162 h2 = __floats2bfloat162_rn(__uint_as_float(v[i + 4]) * inv, __uint_as_float(v[i + 5]) * inv);
                    __nv_bfloat162 h3 = __floats2bfloat162_rn(__uint_as_float(v[i + 6]) * inv, __uint_as_float(v[i + 7]) * inv);
                    pk4.x = *reinterpret_cast<uint32_t*>(&h0); pk4.y = *reinterpret_cast<uint32_t*>(&h1);
                    pk4.z = *reinterpret_cast<uint32_t*>(&h2); pk4.w = *reinterpret_cast<uint32_t*>(&h3);
                    *reinterpret_cast<uint4*>(orow + c + i) = pk4;
                }
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

static int fa_map(tw_ctx* ctx, const __nv_bfloat16* ptr, int rows, int64_t ld, CUtensorMap* out, int box_rows = 128) {
    struct Entry { const void* ptr; int rows; int64_t ld; int box_rows; CUtensorMap map; };
    static Entry cache[16];
    static int n_cached = 0, next = 0;
    for (int i = 0; i < n_cached; ++i)
        if (cache[i].ptr == ptr && cache[i].rows == rows && cache[i].ld == ld && cache[i].box_rows == box_rows) { *out = cache[i].map; return TW_OK; }
    if ((reinterpret_cast<uintptr_t>(ptr) & 15) || (ld % 8)) {
        ctx->set_error(TW_E_UNSUPPORTED, "attention_tc: operands must be 16-byte aligned with a row pitch that is a multiple of 8 elements");
        return TW_E_UNSUPPORTED;
    }
    CUtensorMap m;
    const cuuint64_t dims[2] = {(cuuint64_t)ld, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
    const cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = g_fa_encode(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<__nv_bfloat16*>(ptr), dims, strides, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        ctx->set_error(TW_E_CUDA, "attention_tc: cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
        return TW_E_CUDA;
    }
    Entry& e = cache[next];
    e.ptr = ptr; e.rows = rows; e.ld = ld; e.box_rows = box_rows; e.map = m;
    next = (next + 1) % 16;
    if (n_cached < 16) ++n_cached;
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
        TW_CUDA_OK(ctx, cudaFuncSetAttribute(encoder_attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FA3_SMEM));
    }
    if (causal && Sq != Sk) {
        ctx->set_error(TW_E_INVALID, "attention_tc: the causal mask needs as many queries as keys");
        return TW_E_INVALID;
    }
    CUtensorMap mq, mkv;
    TW_CHECK(fa_map(ctx, q, B * Sq, q_ld, &mq));
    TW_CHECK(fa_map(ctx, kv, B * Sk, kv_ld, &mkv, FA3_BK));
    dim3 grid(ceil_div(Sq, FA_BQ), H, B);
    TW_CUDA_OK(ctx, launch_k(encoder_attention_tc_kernel, grid, dim3(FA_THREADS), FA3_SMEM, st, mq, mkv, out, Sq, Sk, H, q_col0, k_col0, v_col0,
                             causal ? 1 : 0));
    TW_CUDA_OK(ctx, cudaGetLastError());
    return TW_OK;
}

int encoder_attention_tc(tw_ctx* ctx, const __nv_bfloat16* qkv, __nv_bfloat16* out, int B, int S, int H, cudaStream_t st) {
    const int d = H * FA_D;
    return attention_tc(ctx, qkv, 3 * d, 0, qkv, 3 * d, d, 2 * d, out, B, S, S, H, false, st);
}

}  // namespace tw
