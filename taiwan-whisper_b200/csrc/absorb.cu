// "Absorbed" decoder cross-attention on tensor cores (sm_100a): scores and values are taken straight from the encoder output
// E [1500, d] of a clip instead of from per-layer K|V projections of it.
//
//   HF WhisperAttention (modeling_whisper.py:284-358), cross-attention, one query per clip and head h:
//       s_k = q_h . K_h[k],  K_h = E Wk_h^T          =>  s_k = (Wk_h^T q_h) . E[k] = q~_h . E[k]          (q~_h in R^d)
//       o_h = sum_k p_k V_h[k],  V_h = E Wv_h^T + bv  =>  o_h = Wv_h (sum_k p_k E[k]) + bv_h = Wv_h c_h + bv_h
//   so one pass over E (2 d bytes per key, the same for every decoder layer) replaces the pass over the layer's K|V rows
//   (4 d bytes per key): half the streamed bytes, no K|V store, no per-window K|V projection.  q~ and o come from two small
//   grouped GEMMs (gemm_tc_skinny, group = head).
//
// The contraction is 2 x H x d MACs per key — tensor-core work.  A tile of TK keys (TK x d bf16, 160 KB at d = 1280) is brought
// into shared memory ONCE by TMA as d/64 boxes [TK keys x 64 features] (128-byte swizzle) and used twice:
//   pass 1   S^T[key, h]   = E_tile[TK keys, d] . Q~^T          tcgen05.mma M64 N24 K16: A = the boxes, K-major; B = Q~ (K-major)
//   softmax  over the keys = across TMEM lanes; the per-head reference max moves lazily (only when a score exceeds it by 2^8),
//            so the common tile needs one barrier-reduction and no shuffles; P^T (bf16) goes to shared memory
//   pass 2   C^T[feat, h] += E_tile^T[d, TK keys] . P^T         M128 N32 K16: A = the SAME boxes read MN-major; B = P^T (K-major)
// C^T (d/128 tiles x 32 columns, fp32) stays in tensor memory for a whole clip segment.  The ring of 128-feature slices is
// released slice by slice as pass 2 consumes it, and refilled at once with the next tile (at d = 1280 the ring holds exactly one
// tile: 61 KB of Q~ + 160 KB of keys).
//
// One persistent CTA per SM owns a contiguous range of the (active clips x 1500) key rows, as the K|V stream kernel does;
// per clip segment it emits one partial record (m[h], l[h], C[h][d]) and a second kernel merges the records of a clip.
// `rev` walks the CTA's tiles backwards: consecutive decoder layers alternate the direction so that the tail of what layer l
// read is still in L2 when layer l + 1 starts there (E is the same for every layer).
//   warp 4   TMA producer      warp 5   MMA issuer (warp-uniform, elect.sync)      warps 0-3   softmax / epilogue
#include <cuda.h>
#include <stdlib.h>

#include <map>
#include <tuple>

#include "tc_ptx.cuh"

namespace tw {

constexpr int AB_THREADS = 192;
constexpr int AB_NQ = 24;                 // rows of the Q~ box = MMA N of pass 1 (heads padded; H <= 24)
constexpr int AB_NP = 32;                 // MMA N of pass 2 (multiple of 16 at M = 128) = rows of P^T = TMEM tile spacing
constexpr int AB_MAX_NS = 20;             // ring slots (barrier array size)
constexpr int AB_QT_KB = AB_NQ * 128;     // bytes of Q~ per 64-feature K block
constexpr int AB_P_BYTES = AB_NP * 128;   // P^T: 32 head rows x 64 keys bf16, one K-major swizzle atom set
constexpr float AB_RESCALE_LOG2 = 8.0f;
constexpr int AB_SMEM_MAX = 227 * 1024;

__device__ __forceinline__ float ab_ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ void ab_tmem_st32(uint32_t taddr, const uint32_t* r) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
          "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]),
          "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]),
          "r"(r[30]), "r"(r[31]) : "memory");
}
__device__ __forceinline__ void ab_tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// barrier of the 128 softmax threads that also ORs a predicate over them
__device__ __forceinline__ bool ab_bar_or(bool q) {
    uint32_t r;
    asm volatile(
        "{\n.reg .pred p, q;\nsetp.ne.u32 q, %1, 0;\nbar.red.or.pred p, 1, 128, q;\nselp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(r) : "r"((uint32_t)q) : "memory");
    return r != 0;
}
__device__ __forceinline__ void ab_bar() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

// MN-major, 128-byte swizzle shared-memory matrix descriptor (cute::UMMA canonical layout ((8,n),(8,k)):((1,LBO),(8,SBO)) in
// 16-byte units): the operand's MN index runs along the 128-byte rows (64 elements per swizzle atom), its K index down the
// rows.  SBO = 1024 B between 8-row (K) groups; LBO = distance between 64-element atoms along MN (the slice's second box).
// desc_swap = 1 exchanges the two fields (bring-up knob, TWB200_AB_DESC).
__device__ __forceinline__ uint64_t ab_desc_mn(uint32_t smem_addr, uint32_t lbo_bytes, int desc_swap) {
    const uint64_t lbo = desc_swap ? (1024u >> 4) : (lbo_bytes >> 4);
    const uint64_t sbo = desc_swap ? (lbo_bytes >> 4) : (1024u >> 4);
    return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | (lbo << 16) | (sbo << 32) | (1ull << 46) | (2ull << 61);
}
// instruction descriptor with an MN-major A operand (bit 15) and a K-major B operand
__host__ __device__ constexpr uint32_t ab_idesc_amn(int M, int N) { return make_idesc(M, N) | (1u << 15); }

// rows per CTA for a stream of total_rows split over `grid` CTAs (same formula in the merge kernel)
__host__ __device__ inline int ab_rows_per_cta(int64_t total_rows, int grid, int tk) {
    int64_t g = (total_rows + tk - 1) / tk;
    if (g > grid) g = grid;
    if (g < 1) g = 1;
    return (int)((total_rows + g - 1) / g);
}
// partial record: m[AB_NP], l[AB_NP], C[H][d] (fp32)
__host__ __device__ inline size_t ab_rec_floats(int H, int d) { return (size_t)2 * AB_NP + (size_t)H * d; }

// shared-memory layout (from a 1024-aligned base): Q~ | ring | P^T | sMax | barriers | tmem pointer
__host__ __device__ inline size_t ab_fixed_bytes(int d) { return (size_t)(d / 64) * AB_QT_KB + AB_P_BYTES + 4 * AB_NP * 4 + (2 * AB_MAX_NS + 8) * 8 + 16; }

// The CTA's share of the key rows, cut at clip boundaries into segments and into tiles of TK keys; processing order reversed by rev.
struct AbRange {
    int64_t row_begin, row_end;
    int Tk, slot_first, nseg;
    __device__ __forceinline__ AbRange(int64_t total, int R, int Tk_) : Tk(Tk_) {
        row_begin = (int64_t)blockIdx.x * R;
        row_end = min(total, row_begin + R);
        if (row_begin >= total) { nseg = 0; slot_first = 0; return; }
        slot_first = (int)(row_begin / Tk);
        nseg = (int)((row_end - 1) / Tk) - slot_first + 1;
    }
    // segment k of the processing order: slot (index into the active list), first key inside the clip, number of keys
    __device__ __forceinline__ void seg(int k, int rev, int& slot, int& key0, int& nkeys) const {
        const int kk = rev ? nseg - 1 - k : k;
        slot = slot_first + kk;
        const int64_t a = max(row_begin, (int64_t)slot * Tk), b = min(row_end, (int64_t)(slot + 1) * Tk);
        key0 = (int)(a - (int64_t)slot * Tk);
        nkeys = (int)(b - a);
    }
};

template <int TK>
__global__ void __launch_bounds__(AB_THREADS, 1)
absorbed_attention_kernel(const __grid_constant__ CUtensorMap map_e, const __grid_constant__ CUtensorMap map_qt, int Tk, int B, int H,
                          int d, int NS, float* __restrict__ partial, const int32_t* __restrict__ active,
                          const int32_t* __restrict__ n_active, int rev, int desc_swap, long long* __restrict__ trace) {
    // bring-up timeline (TWB200_AB_TRACE): clock64 of CTA 0's pipeline events, 64 slots per tile
#define AB_TRACE(tile, slot) do { if (trace && blockIdx.x == 0 && (tile) < 16 && lane == 0) trace[(tile) * 64 + (slot)] = clock64(); } while (0)
    constexpr int BOX = TK * 128;             // one TMA box: TK keys x 64 features
    constexpr int SLICE = 2 * BOX;            // one ring slot: 128 features of a tile
    extern __shared__ unsigned char ab_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(ab_raw) + 1023) & ~(uintptr_t)1023);
    const int kbt = d / 64;                  // 64-feature K blocks of pass 1
    const int npair = d / 128;               // slices per tile = 128-feature M tiles of pass 2
    unsigned char* sQt = smem;                                        // kbt x (24 rows x 128 B)
    unsigned char* sRing = smem + (size_t)kbt * AB_QT_KB;             // NS slots (1024-aligned)
    unsigned char* sP = sRing + (size_t)NS * SLICE;                   // P^T
    float* sMax = reinterpret_cast<float*>(sP + AB_P_BYTES);          // [4 warps][AB_NP]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sMax + 4 * AB_NP);
    uint64_t* st_full = bars;                      // AB_MAX_NS
    uint64_t* st_empty = bars + AB_MAX_NS;         // AB_MAX_NS
    uint64_t* qt_full = bars + 2 * AB_MAX_NS;      // 1: Q~ of the clip segment has landed
    uint64_t* s_full = qt_full + 1;                // 2: scores of tile parity b are complete
    uint64_t* p_full = s_full + 2;                 // 1: P^T written (4 softmax warps)
    uint64_t* pv_done = p_full + 1;                // 1: pass 2 of a tile has completed (one phase per tile)
    uint64_t* seg_done = pv_done + 1;              // 1: every MMA of a clip segment has completed (Q~ may be replaced)
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(seg_done + 2);

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
    pdl_trigger();
    if (threadIdx.x == 0) {
        tma_prefetch_desc(&map_e);
        tma_prefetch_desc(&map_qt);
        for (int i = 0; i < NS; ++i) {
            mbar_init(smem_u32(&st_full[i]), 1);
            mbar_init(smem_u32(&st_empty[i]), 1);
        }
        mbar_init(smem_u32(qt_full), 1);
        mbar_init(smem_u32(&s_full[0]), 1);
        mbar_init(smem_u32(&s_full[1]), 1);
        mbar_init(smem_u32(p_full), 4);
        mbar_init(smem_u32(pv_done), 1);
        mbar_init(smem_u32(seg_done), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 5) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (warp < 4) {       // P^T rows of the padding heads stay zero
        uint4* p4 = reinterpret_cast<uint4*>(sP);
        for (int i = threadIdx.x; i < AB_P_BYTES / 16; i += 128) p4[i] = make_uint4(0, 0, 0, 0);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    const uint32_t tmem_S = tmem_base;              // two score buffers of AB_NP columns (lanes 0-15 of every 32: M = 64)
    const uint32_t tmem_C = tmem_base + 2 * AB_NP;  // npair tiles of AB_NP columns: C^T[feature lane][head]

    // the active-clip list of this step was written by the previous step's advance kernel (full dependency at the step's start)
    if (active) B = *n_active;
    const int64_t total = (int64_t)B * Tk;
    const AbRange rg(total, ab_rows_per_cta(total, gridDim.x, TK), Tk);

    if (warp == 4) {
        // ===================== TMA producer =====================
        uint32_t c = 0;                           // slices issued so far (ring position c % NS, fill round c / NS)
        auto load_slice = [&](int row, int p) {
            const int slot = c % NS;
            mbar_wait(smem_u32(&st_empty[slot]), ((c / NS) & 1) ^ 1);
            AB_TRACE((int)(c / npair), (int)(c % npair));
            if (elect_one_sync()) {
                const uint32_t fb = smem_u32(&st_full[slot]);
                const uint32_t dst = smem_u32(sRing + (size_t)slot * SLICE);
                mbar_expect_tx(fb, SLICE);
                tma_load_2d(&map_e, fb, dst, p * 128, row);
                tma_load_2d(&map_e, fb, dst + BOX, p * 128 + 64, row);
            }
            __syncwarp();
            ++c;
        };
        uint32_t skip = 0;                        // slices of the first segment already issued ahead of the dependency wait
        for (int k = 0; k < rg.nseg; ++k) {
            int slot, key0, nkeys;
            rg.seg(k, rev, slot, key0, nkeys);
            const int clip = active ? active[slot] : slot;
            const int nt = (nkeys + TK - 1) / TK;
            const int row0 = clip * Tk + key0;
            if (k == 0) {
                // E does not depend on the previous kernel: fill the ring before griddepcontrol.wait; Q~ does
                const uint32_t ahead = min((uint32_t)NS, (uint32_t)(nt * npair));
                for (uint32_t i = 0; i < ahead; ++i) {
                    const int t = i / npair, tt = rev ? nt - 1 - t : t;
                    load_slice(row0 + tt * TK, i % npair);
                }
                skip = ahead;
                pdl_wait();
            } else {
                mbar_wait(smem_u32(seg_done), (k - 1) & 1);        // previous segment's MMAs no longer read Q~
            }
            if (elect_one_sync()) {
                mbar_expect_tx(smem_u32(qt_full), (uint32_t)kbt * AB_QT_KB);
                for (int kb = 0; kb < kbt; ++kb) tma_load_2d(&map_qt, smem_u32(qt_full), smem_u32(sQt + kb * AB_QT_KB), kb * 64, clip * H);
            }
            __syncwarp();
            for (uint32_t i = (k == 0 ? skip : 0); i < (uint32_t)(nt * npair); ++i) {
                const int t = i / npair, tt = rev ? nt - 1 - t : t;
                load_slice(row0 + tt * TK, i % npair);
            }
        }
        if (rg.nseg == 0) pdl_wait();
    } else if (warp == 5) {
        // ===================== MMA issuer =====================
        constexpr uint32_t idesc_qk = make_idesc(64, AB_NQ);
        constexpr uint32_t idesc_pv = ab_idesc_amn(128, AB_NP);
        uint32_t c1 = 0, c2 = 0;                   // slices issued by pass 1 / pass 2 (ring position c % NS, fill round c / NS)
        int g = 0;                                 // tiles completed by this CTA (score-buffer / barrier phases)
        for (int k = 0; k < rg.nseg; ++k) {
            int slot, key0, nkeys;
            rg.seg(k, rev, slot, key0, nkeys);
            const int nt = (nkeys + TK - 1) / TK;
            mbar_wait(smem_u32(qt_full), k & 1);
            tc_fence_after();
            int p1_tile = 0, p1_slice = 0;         // progress of pass 1 inside the segment
            auto pass1_slice = [&]() {             // S^T(buffer) (+)= E_slice[TK keys, 128 features] . Q~^T
                const int ring = c1 % NS;
                mbar_wait(smem_u32(&st_full[ring]), (c1 / NS) & 1);
                tc_fence_after();
                AB_TRACE(g + p1_tile, 10 + p1_slice);
                const uint32_t d_tmem = tmem_S + ((g + p1_tile) & 1) * AB_NP;
                if (elect_one_sync()) {
                    const uint32_t sa = smem_u32(sRing + (size_t)ring * SLICE);
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                        const uint64_t a_desc = make_sw128_desc(sa + half * BOX);
                        const uint64_t b_desc = make_sw128_desc(smem_u32(sQt + (2 * p1_slice + half) * AB_QT_KB));
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk)
                            tc_mma_f16(d_tmem, a_desc + 2 * kk, b_desc + 2 * kk, idesc_qk, (p1_slice > 0 || half > 0 || kk > 0) ? 1u : 0u);
                    }
                    if (p1_slice == npair - 1) tc_commit(smem_u32(&s_full[(g + p1_tile) & 1]));
                }
                __syncwarp();
                ++c1;
                if (++p1_slice == npair) { p1_slice = 0; ++p1_tile; }
            };
            for (int t = 0; t < nt; ++t) {
                while (p1_tile <= t) pass1_slice();
                // look ahead into tile t + 1 as far as the ring allows without waiting for slots that pass 2 of tile t releases
                while (p1_tile == t + 1 && p1_tile < nt && c1 < c2 + (uint32_t)NS) pass1_slice();
                // pass 2: C^T(tile p) += E_slice^T . P^T
                AB_TRACE(g + t, 20);
                mbar_wait(smem_u32(p_full), (g + t) & 1);
                tc_fence_after();
                AB_TRACE(g + t, 21);
                for (int p = 0; p < npair; ++p) {
                    const int ring = c2 % NS;
                    mbar_wait(smem_u32(&st_full[ring]), (c2 / NS) & 1);
                    tc_fence_after();
                    if (elect_one_sync()) {
                        const uint32_t sa = smem_u32(sRing + (size_t)ring * SLICE);
                        const uint64_t b_base = make_sw128_desc(smem_u32(sP));
#pragma unroll
                        for (int kk = 0; kk < TK / 16; ++kk)
                            tc_mma_f16(tmem_C + p * AB_NP, ab_desc_mn(sa + kk * 2048, BOX, desc_swap), b_base + 2 * kk, idesc_pv,
                                       (t > 0 || kk > 0) ? 1u : 0u);
                        tc_commit(smem_u32(&st_empty[ring]));
                        if (p == npair - 1) {
                            tc_commit(smem_u32(pv_done));
                            if (t == nt - 1) tc_commit(smem_u32(seg_done));
                        }
                    }
                    __syncwarp();
                    ++c2;
                }
                AB_TRACE(g + t, 22);
            }
            g += nt;
        }
    } else {
        // ===================== softmax / epilogue warps 0..3 =====================
        // pass 1 (M = 64): key 16 w + i of the tile sits in TMEM lane 32 w + i (i < 16); epilogue (M = 128): lane = feature
        const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
        const int kk = warp * 16 + lane;                                  // this thread's key inside a tile (lane < 16)
        const float LOG2E = 1.4426950408889634f;
        int g = 0;
        for (int k = 0; k < rg.nseg; ++k) {
            int slot, key0, nkeys;
            rg.seg(k, rev, slot, key0, nkeys);
            const int nt = (nkeys + TK - 1) / TK;
            float m_ref[AB_NQ], l_part[AB_NQ];
#pragma unroll
            for (int h = 0; h < AB_NQ; ++h) { m_ref[h] = -INFINITY; l_part[h] = 0.0f; }
            for (int t = 0; t < nt; ++t, ++g) {
                const int tt = rev ? nt - 1 - t : t;
                const bool valid = lane < 16 && kk < TK && tt * TK + kk < nkeys;     // this lane holds a key of the segment
                if (warp == 0) AB_TRACE(g, 29);
                mbar_wait(smem_u32(&s_full[g & 1]), (g >> 1) & 1);
                tc_fence_after();
                if (warp == 0) AB_TRACE(g, 30);
                uint32_t sv[32];
                tmem_ld32(tmem_S + lane_off + (g & 1) * AB_NP, sv);
                tmem_ld_wait();
                if (warp == 0) AB_TRACE(g, 31);
                bool slow = (t == 0);
                if (t > 0) {
                    bool grow = false;
#pragma unroll
                    for (int h = 0; h < AB_NQ; ++h)
                        grow |= valid && h < H && (__uint_as_float(sv[h]) - m_ref[h]) * LOG2E > AB_RESCALE_LOG2;
                    slow = ab_bar_or(grow);
                }
                // P^T (single buffer) is read and C^T accumulated by pass 2 of the previous tile
                if (warp == 0) AB_TRACE(g, 32);
                if (g > 0) {
                    mbar_wait(smem_u32(pv_done), (g - 1) & 1);
                    tc_fence_after();
                }
                if (warp == 0) AB_TRACE(g, 33);
                if (slow) {
                    // exact per-head maximum of the tile: 16 lanes by shuffles, the four warps through shared memory
                    float mx[AB_NQ];
#pragma unroll
                    for (int h = 0; h < AB_NQ; ++h) {
                        float v = valid ? __uint_as_float(sv[h]) : -INFINITY;
#pragma unroll
                        for (int o = 8; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
                        mx[h] = v;
                    }
                    if (lane == 0) {
#pragma unroll
                        for (int h = 0; h < AB_NQ; ++h) sMax[warp * AB_NP + h] = mx[h];
                    }
                    ab_bar();
                    float f[AB_NQ];
                    bool any_f = false;
#pragma unroll
                    for (int h = 0; h < AB_NQ; ++h) {
                        const float m4 = fmaxf(fmaxf(sMax[h], sMax[AB_NP + h]), fmaxf(sMax[2 * AB_NP + h], sMax[3 * AB_NP + h]));
                        const float mn = fmaxf(m_ref[h], m4);
                        f[h] = (mn > m_ref[h] && m_ref[h] > -INFINITY) ? ab_ex2((m_ref[h] - mn) * LOG2E) : 1.0f;
                        any_f |= f[h] != 1.0f;
                        m_ref[h] = mn;
                        l_part[h] *= f[h];
                    }
                    ab_bar();                      // sMax may be rewritten
                    if (t > 0 && any_f) {          // uniform over the CTA: every thread holds the same maxima
                        for (int p = 0; p < npair; ++p) {
                            uint32_t cv[32];
                            tmem_ld32(tmem_C + lane_off + p * AB_NP, cv);
                            tmem_ld_wait();
#pragma unroll
                            for (int h = 0; h < AB_NQ; ++h) cv[h] = __float_as_uint(__uint_as_float(cv[h]) * f[h]);
                            ab_tmem_st32(tmem_C + lane_off + p * AB_NP, cv);
                        }
                        ab_tmem_wait_st();
                    }
                }
                // P^T[h][key] = exp2((s - m_ref) log2e) as bf16, K-major 128-byte swizzle (one 64-key atom per 8 heads)
                if (lane < 16) {
#pragma unroll
                    for (int h = 0; h < AB_NQ; ++h) {
                        if (h < H) {
                            // a fully masked segment-first tile cannot happen (nkeys >= 1 and t == 0 sets m_ref from valid keys)
                            const float p = valid ? ab_ex2((__uint_as_float(sv[h]) - m_ref[h]) * LOG2E) : 0.0f;
                            const __nv_bfloat16 pb = __float2bfloat16_rn(p);
                            l_part[h] += __bfloat162float(pb);
                            *reinterpret_cast<__nv_bfloat16*>(sP + h * 128 + ((((kk >> 3) ^ (h & 7))) << 4) + (kk & 7) * 2) = pb;
                        }
                    }
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(p_full));
                if (warp == 0) AB_TRACE(g, 34);
            }
            // ---- end of the clip segment: partial record (m, l, C) of this (CTA, clip)
            mbar_wait(smem_u32(pv_done), (g - 1) & 1);
            tc_fence_after();
            float* rec = partial + ((size_t)blockIdx.x + slot) * ab_rec_floats(H, d);
#pragma unroll
            for (int h = 0; h < AB_NQ; ++h) l_part[h] = warp_sum(l_part[h]);
            if (lane == 0) {
#pragma unroll
                for (int h = 0; h < AB_NQ; ++h) sMax[warp * AB_NP + h] = l_part[h];
            }
            ab_bar();
            if (threadIdx.x < AB_NQ) {
                const int h = threadIdx.x;
                float mine = m_ref[0];
#pragma unroll
                for (int j = 1; j < AB_NQ; ++j) mine = (h == j) ? m_ref[j] : mine;
                rec[h] = mine;
                rec[AB_NP + h] = sMax[h] + sMax[AB_NP + h] + sMax[2 * AB_NP + h] + sMax[3 * AB_NP + h];
            }
            ab_bar();
            for (int p = 0; p < npair; ++p) {
                uint32_t cv[32];
                tmem_ld32(tmem_C + lane_off + p * AB_NP, cv);
                tmem_ld_wait();
                float* dst = rec + 2 * AB_NP + p * 128 + warp * 32 + lane;    // C[h][feature]: consecutive lanes -> consecutive floats
#pragma unroll
                for (int h = 0; h < AB_NQ; ++h)
                    if (h < H) dst[(size_t)h * d] = __uint_as_float(cv[h]);
            }
            tc_fence_before();
        }
        if (rg.nseg == 0) pdl_wait();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 5) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

// merge the partial records of a clip: ctx[clip][h*d + c] = sum_r w_r C_r[h][c] / sum_r w_r l_r[h], w_r = 2^((m_r - m) log2e)
template <typename T>
__global__ void __launch_bounds__(256)
absorbed_attention_combine(const float* __restrict__ partial, int Tk, int B, int grid, int tk, int H, int d, T* __restrict__ ctx,
                           const int32_t* __restrict__ active, const int32_t* __restrict__ n_active) {
    pdl_trigger();
    pdl_wait();
    const int slot = blockIdx.y, h = blockIdx.x;
    int clip = slot;
    if (active) {
        B = *n_active;
        if (slot >= B) return;
        clip = active[slot];
    }
    const float LOG2E = 1.4426950408889634f;
    const int R = ab_rows_per_cta((int64_t)B * Tk, grid, tk);
    const int c_first = (int)(((int64_t)slot * Tk) / R), c_last = (int)((((int64_t)slot + 1) * Tk - 1) / R);
    const size_t rf = ab_rec_floats(H, d);
    float m = -INFINITY;
    for (int c = c_first; c <= c_last; ++c) m = fmaxf(m, partial[(size_t)(c + slot) * rf + h]);
    float l = 0.0f;
    for (int c = c_first; c <= c_last; ++c) {
        const float* rec = partial + (size_t)(c + slot) * rf;
        l += rec[AB_NP + h] * ab_ex2((rec[h] - m) * LOG2E);
    }
    const float inv = 1.0f / l;
    for (int i = threadIdx.x; i < d; i += blockDim.x) {
        float o = 0.0f;
        for (int c = c_first; c <= c_last; ++c) {
            const float* rec = partial + (size_t)(c + slot) * rf;
            o += rec[2 * AB_NP + (size_t)h * d + i] * ab_ex2((rec[h] - m) * LOG2E);
        }
        ctx[((size_t)clip * H + h) * d + i] = from_f32<T>(o * inv);
    }
}

// ---- host ----------------------------------------------------------------------------------------
typedef CUresult (*AbEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static AbEncodeFn g_ab_encode = nullptr;
using AbKey = std::tuple<const void*, int64_t, int64_t, int>;
static std::map<AbKey, CUtensorMap> g_ab_maps;

static int ab_map(tw_ctx* ctx, const void* ptr, int64_t rows, int64_t cols, int box_rows, CUtensorMap* out) {
    const AbKey key(ptr, rows, cols, box_rows);
    auto it = g_ab_maps.find(key);
    if (it != g_ab_maps.end()) { *out = it->second; return TW_OK; }
    CUtensorMap m;
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
    const cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = g_ab_encode(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        ctx->set_error(TW_E_CUDA, "absorbed_attention: cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
        return TW_E_CUDA;
    }
    if (g_ab_maps.size() > 1024) g_ab_maps.clear();
    g_ab_maps[key] = m;
    *out = m;
    return TW_OK;
}

static int ab_tile_keys() {
    static const int tk = getenv("TWB200_AB_TK") ? atoi(getenv("TWB200_AB_TK")) : 64;
    return (tk == 48 || tk == 32) ? tk : 64;
}
// ring slots that fit next to Q~ (at least one whole tile)
static int ab_ring_slots(int d, int tk) {
    const int64_t avail = (int64_t)AB_SMEM_MAX - 1024 - (int64_t)ab_fixed_bytes(d);
    int64_t ns = avail / (2 * tk * 128);
    if (ns > AB_MAX_NS) ns = AB_MAX_NS;
    return (int)ns;
}
static size_t ab_smem_bytes(int d, int tk, int ns) { return 1024 + ab_fixed_bytes(d) + (size_t)ns * 2 * tk * 128; }

bool absorbed_attention_supported(int H, int d) {
    return H >= 1 && H <= AB_NQ && d % 128 == 0 && d >= 128 && d <= 1280 && ab_ring_slots(d, ab_tile_keys()) >= d / 128;
}

constexpr int AB_MAX_CTAS = 148;          // persistent CTAs (one per SM of a B200); sizes the partial-record workspace
size_t absorbed_attention_partial_floats(int B, int H, int d) { return (size_t)(AB_MAX_CTAS + B) * ab_rec_floats(H, d); }

int absorbed_attention_init(tw_ctx* ctx) {
    if (!g_ab_encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
        if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) {
            ctx->set_error(TW_E_CUDA, "cuTensorMapEncodeTiled not available from the driver");
            return TW_E_CUDA;
        }
        g_ab_encode = reinterpret_cast<AbEncodeFn>(fn);
    }
    TW_CUDA_OK(ctx, cudaFuncSetAttribute(absorbed_attention_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, AB_SMEM_MAX));
    TW_CUDA_OK(ctx, cudaFuncSetAttribute(absorbed_attention_kernel<48>, cudaFuncAttributeMaxDynamicSharedMemorySize, AB_SMEM_MAX));
    TW_CUDA_OK(ctx, cudaFuncSetAttribute(absorbed_attention_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, AB_SMEM_MAX));
    return TW_OK;
}

// qt: [B*H + AB_NQ rows allocated, d] bf16 — row clip*H + h holds q~ of that clip and head; enc: [B*Tk, d] bf16;
// ctx_out: [B, H*d] bf16 (A operand of the grouped value projection)
static_assert(ABSORB_QT_PAD == AB_NQ, "the Q~ matrix is padded by one TMA box of rows");
int absorbed_attention(tw_ctx* ctx, const __nv_bfloat16* qt, const __nv_bfloat16* enc, int Tk, int B, int H, int d, float* partial,
                       __nv_bfloat16* ctx_out, cudaStream_t st, const int32_t* active, const int32_t* n_active, int rev, cudaEvent_t ev0,
                       cudaEvent_t ev1, long long* trace) {
    if (!absorbed_attention_supported(H, d) || !g_ab_encode) {
        ctx->set_error(TW_E_UNSUPPORTED, "absorbed_attention: needs H <= 24, d a multiple of 128 up to 1280 and absorbed_attention_init");
        return TW_E_UNSUPPORTED;
    }
    const int tk = ab_tile_keys();
    const int ns = ab_ring_slots(d, tk);
    CUtensorMap me, mq;
    TW_CHECK(ab_map(ctx, enc, (int64_t)B * Tk, d, tk, &me));
    TW_CHECK(ab_map(ctx, qt, (int64_t)B * H + AB_NQ, d, AB_NQ, &mq));
    static const int desc_swap = getenv("TWB200_AB_DESC") ? atoi(getenv("TWB200_AB_DESC")) : 0;
    const int G = ctx->sm_count < AB_MAX_CTAS ? ctx->sm_count : AB_MAX_CTAS;
    if (ev0) cudaEventRecord(ev0, st);
    if (tk == 64)
        TW_CUDA_OK(ctx, launch_k(absorbed_attention_kernel<64>, dim3(G), dim3(AB_THREADS), ab_smem_bytes(d, tk, ns), st, me, mq, Tk, B, H, d, ns,
                                 partial, active, n_active, rev, desc_swap, trace));
    else if (tk == 32)
        TW_CUDA_OK(ctx, launch_k(absorbed_attention_kernel<32>, dim3(G), dim3(AB_THREADS), ab_smem_bytes(d, tk, ns), st, me, mq, Tk, B, H, d, ns,
                                 partial, active, n_active, rev, desc_swap, trace));
    else
        TW_CUDA_OK(ctx, launch_k(absorbed_attention_kernel<48>, dim3(G), dim3(AB_THREADS), ab_smem_bytes(d, tk, ns), st, me, mq, Tk, B, H, d, ns,
                                 partial, active, n_active, rev, desc_swap, trace));
    if (ev1) cudaEventRecord(ev1, st);
    TW_CUDA_OK(ctx, launch_k(absorbed_attention_combine<__nv_bfloat16>, dim3(H, B), dim3(256), 0, st, (const float*)partial, Tk, B, G, tk, H, d,
                             ctx_out, active, n_active));
    TW_CUDA_OK(ctx, cudaGetLastError());
    return TW_OK;
}

}  // namespace tw
