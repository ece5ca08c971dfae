// K1 — fused Whisper log-mel front end for sm_100a.
//
// One CTA turns 32 consecutive STFT frames of one clip into 32 columns of the log-mel matrix:
//   stage the 5360 input samples (reflect-padded, int16 -> f32 dequantised, vectorised loads) in
//   shared memory -> periodic-Hann window -> 400-point real FFT done as a 200-point complex FFT
//   (200 = 8 x 5 x 5, mixed radix, all butterflies in registers, one shared-memory transpose) ->
//   real-FFT split -> |X|^2 -> sparse slaney mel filterbank held in shared memory -> log10 clamp.
// The per-clip `max - 8` floor of the reference is a clip-wide reduction: pass 1 writes the
// clamped log10 values and atomically maxes one float per clip; pass 2 (logmel_finalize) applies
// max(., clipmax-8) and (x+4)/4 in place while the tile is still L2-resident.
//
// Arithmetic restated from transformers/models/whisper/feature_extraction_whisper.py:135-164
// (torch.stft n_fft=400 hop=160 center/reflect, drop last frame, mel_filters.T @ |stft|^2,
// clamp 1e-10, log10, per-clip max-8, (x+4)/4) and its in-reference twin
// ref: training/flax/distil_whisper/pipeline.py:40-58; filterbank from
// transformers/audio_utils.py:263-333,356-375,453-545 (slaney scale, slaney norm, 0-8000 Hz).
#include <math.h>
#include <stdlib.h>

#include <vector>

#include "common.cuh"

namespace tw {

constexpr int LM_THREADS = 256;
constexpr int LM_NFFT = 400;
constexpr int LM_HOP = 160;
// FR = frames per tile.  32: 2 CTAs / SM (109 KB); 16: the power spectrum re-uses the sample staging buffer and the CTA
// shrinks to ~49 KB, so 3 CTAs (24 warps) fit an SM — the kernel is bound by issue slots and dependency stalls
// (profiles/r02_logmel_ncu.md), which more resident warps hide
template <int FR> struct LmCfg {
    static constexpr int NX = FR * LM_HOP + (LM_NFFT - LM_HOP);   // staged samples (5360 for 32 frames)
    static constexpr int CTAS = FR == 32 ? 2 : 3;
};
constexpr int LM_NBIN = 201;
constexpr int LM_MAX_NNZ = 1024;

struct cpx {
    float r, i;
};
__device__ __forceinline__ cpx operator+(cpx a, cpx b) { return {a.r + b.r, a.i + b.i}; }
__device__ __forceinline__ cpx operator-(cpx a, cpx b) { return {a.r - b.r, a.i - b.i}; }
__device__ __forceinline__ cpx cmul(cpx a, cpx b) { return {a.r * b.r - a.i * b.i, a.r * b.i + a.i * b.r}; }
__device__ __forceinline__ cpx mul_neg_i(cpx a) { return {a.i, -a.r}; }   // a * (-i)
__device__ __forceinline__ cpx mul_pos_i(cpx a) { return {-a.i, a.r}; }   // a * (+i)

// forward 4-point DFT
__device__ __forceinline__ void dft4(cpx a0, cpx a1, cpx a2, cpx a3, cpx& A0, cpx& A1, cpx& A2, cpx& A3) {
    cpx s0 = a0 + a2, s1 = a0 - a2, s2 = a1 + a3, s3 = a1 - a3;
    A0 = s0 + s2;
    A2 = s0 - s2;
    A1 = s1 + mul_neg_i(s3);
    A3 = s1 + mul_pos_i(s3);
}

// forward 8-point DFT, in place
__device__ __forceinline__ void dft8(cpx* x) {
    cpx E0, E1, E2, E3, O0, O1, O2, O3;
    dft4(x[0], x[2], x[4], x[6], E0, E1, E2, E3);
    dft4(x[1], x[3], x[5], x[7], O0, O1, O2, O3);
    const float h = 0.70710678118654752440f;
    cpx t1 = {h * (O1.r + O1.i), h * (O1.i - O1.r)};      // O1 * (h, -h)
    cpx t2 = mul_neg_i(O2);
    cpx t3 = {h * (O3.i - O3.r), -h * (O3.r + O3.i)};     // O3 * (-h, -h)
    x[0] = E0 + O0; x[4] = E0 - O0;
    x[1] = E1 + t1; x[5] = E1 - t1;
    x[2] = E2 + t2; x[6] = E2 - t2;
    x[3] = E3 + t3; x[7] = E3 - t3;
}

// forward 5-point DFT, in place on x[0], x[s], x[2s], x[3s], x[4s]
template <int S>
__device__ __forceinline__ void dft5(cpx* x) {
    const float c1 = 3.090169944e-01f, c2 = -8.090169944e-01f, s1 = 9.510565163e-01f, s2 = 5.877852523e-01f;
    cpx x0 = x[0], x1 = x[S], x2 = x[2 * S], x3 = x[3 * S], x4 = x[4 * S];
    cpx t1 = x1 + x4, t2 = x2 + x3, t3 = x1 - x4, t4 = x2 - x3;
    cpx m1 = {x0.r + c1 * t1.r + c2 * t2.r, x0.i + c1 * t1.i + c2 * t2.i};
    cpx m2 = {x0.r + c2 * t1.r + c1 * t2.r, x0.i + c2 * t1.i + c1 * t2.i};
    cpx n1 = {s1 * t3.r + s2 * t4.r, s1 * t3.i + s2 * t4.i};
    cpx n2 = {s2 * t3.r - s1 * t4.r, s2 * t3.i - s1 * t4.i};
    x[0] = {x0.r + t1.r + t2.r, x0.i + t1.i + t2.i};
    x[S] = m1 + mul_neg_i(n1);
    x[4 * S] = m1 + mul_pos_i(n1);
    x[2 * S] = m2 + mul_neg_i(n2);
    x[3 * S] = m2 + mul_pos_i(n2);
}

__device__ __constant__ float c_cos25[17] = {
    1.000000000e+00f, 9.685831611e-01f, 8.763066800e-01f, 7.289686274e-01f, 5.358267950e-01f, 3.090169944e-01f,
    6.279051953e-02f, -1.873813146e-01f, -4.257792916e-01f, -6.374239897e-01f, -8.090169944e-01f, -9.297764859e-01f,
    -9.921147013e-01f, -9.921147013e-01f, -9.297764859e-01f, -8.090169944e-01f, -6.374239897e-01f};
__device__ __constant__ float c_sin25[17] = {
    0.000000000e+00f, 2.486898872e-01f, 4.817536741e-01f, 6.845471059e-01f, 8.443279255e-01f, 9.510565163e-01f,
    9.980267284e-01f, 9.822872507e-01f, 9.048270525e-01f, 7.705132428e-01f, 5.877852523e-01f, 3.681245527e-01f,
    1.253332336e-01f, -1.253332336e-01f, -3.681245527e-01f, -5.877852523e-01f, -7.705132428e-01f};

// tables built on the host in double precision (see build_tables below)
struct LogmelTables {
    float win[LM_NFFT];        // periodic Hann
    float2 w200[200];          // exp(-2 pi i m / 200)
    float2 w400[LM_NBIN];      // exp(-2 pi i k / 400)
};
__device__ LogmelTables g_tables;

__device__ __forceinline__ void atomic_max_float(float* addr, float v) {
    if (v >= 0.0f) atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
    else atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}

template <typename SampleT> __device__ __forceinline__ float sample_to_f32(SampleT v);
template <> __device__ __forceinline__ float sample_to_f32<int16_t>(int16_t v) { return (float)v * (1.0f / 32768.0f); }
template <> __device__ __forceinline__ float sample_to_f32<float>(float v) { return v; }

template <int FR> struct LmSmem {
    static constexpr int XPW = (LmCfg<FR>::NX + 16) > FR * LM_NBIN ? (LmCfg<FR>::NX + 16) : FR * LM_NBIN;
    // staged samples x[NX] (stage 0-1) and the power spectrum pw[FR][201] (stage 3-4) share one buffer for FR = 16
    float xpw[FR == 32 ? (LmCfg<FR>::NX + 16) + FR * LM_NBIN : XPW];
    float2 z[FR][200];
    float win[LM_NFFT];
    float2 w200[200];
    float2 w400[LM_NBIN];
    float fbw[LM_MAX_NNZ];
    int fb_start[128], fb_count[128], fb_off[128];
    unsigned short zpos[LM_NBIN];      // where bin k of the 200-point FFT lives inside a frame's z[] (digit-reversed order)
    float red[LM_THREADS / 32];
};

template <typename SampleT, int LM_FR>
__global__ void __launch_bounds__(LM_THREADS, LmCfg<LM_FR>::CTAS)
logmel_kernel(const SampleT* __restrict__ pcm, int64_t pcm_stride, const int32_t* __restrict__ n_valid_arr, int B,
              int n_mel, const int* __restrict__ fb_start, const int* __restrict__ fb_count,
              const int* __restrict__ fb_off, const float* __restrict__ fb_w, int fb_nnz,
              float* __restrict__ out, float* __restrict__ clip_max) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int LM_NX = LmCfg<LM_FR>::NX;
    LmSmem<LM_FR>& s = *reinterpret_cast<LmSmem<LM_FR>*>(smem_raw);
    float* const sx = s.xpw;                                                           // staged samples
    float (*const spw)[LM_NBIN] = reinterpret_cast<float (*)[LM_NBIN]>(LM_FR == 32 ? s.xpw + LM_NX + 16 : s.xpw);   // power spectrum
    const int tid = threadIdx.x;

    // ---- tables -> smem, once per (persistent) CTA
    for (int i = tid; i < LM_NFFT; i += LM_THREADS) s.win[i] = g_tables.win[i];
    for (int i = tid; i < 200; i += LM_THREADS) s.w200[i] = g_tables.w200[i];
    for (int i = tid; i < LM_NBIN; i += LM_THREADS) {
        s.w400[i] = g_tables.w400[i];
        const int k = (i == 200) ? 0 : i, q = k >> 3;
        s.zpos[i] = (unsigned short)(25 * (k & 7) + 5 * (q % 5) + q / 5);
    }
    for (int i = tid; i < fb_nnz; i += LM_THREADS) s.fbw[i] = fb_w[i];
    for (int i = tid; i < n_mel; i += LM_THREADS) {
        s.fb_start[i] = fb_start[i];
        s.fb_count[i] = fb_count[i];
        s.fb_off[i] = fb_off[i];
    }

    const int tiles_per_clip = (TW_N_FRAMES + LM_FR - 1) / LM_FR;
    const int n_tiles = B * tiles_per_clip;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int b = tile / tiles_per_clip;
    const int f0 = (tile - b * tiles_per_clip) * LM_FR;
    const SampleT* x = pcm + (int64_t)b * pcm_stride;
    // n_valid is authoritative when given (rows may overlap: long-form windows are rows of pitch < 480000);
    // without it a row is pcm_stride samples long
    int64_t nv = n_valid_arr ? (int64_t)n_valid_arr[b] : pcm_stride;
    if (nv > TW_N_SAMPLES) nv = TW_N_SAMPLES;
    if (nv < 0) nv = 0;
    const int n_valid = (int)nv;
    __syncthreads();            // previous tile fully consumed (and the tables are in place)

    // ---- stage samples: smem x[i] = xp[160 f0 + i], xp = reflect-padded (200) zero-extended clip
    const int base = LM_HOP * f0 - LM_NFFT / 2;    // clip index of x[0]; multiple of 8 samples
    constexpr int VEC = 16 / sizeof(SampleT);      // samples per 16-byte load
    const bool row_aligned = ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
    for (int v = tid; v < LM_NX / VEC; v += LM_THREADS) {
        const int i0 = v * VEC;
        const int j0 = base + i0;
        if (row_aligned && j0 >= 0 && j0 + VEC <= n_valid) {
            const int4 raw = __ldg(reinterpret_cast<const int4*>(x + j0));
            const SampleT* e = reinterpret_cast<const SampleT*>(&raw);
#pragma unroll
            for (int q = 0; q < VEC; ++q) sx[i0 + q] = sample_to_f32<SampleT>(e[q]);
        } else {
#pragma unroll
            for (int q = 0; q < VEC; ++q) {
                int j = j0 + q;
                if (j < 0) j = -j;
                else if (j >= TW_N_SAMPLES) j = 2 * (TW_N_SAMPLES - 1) - j;
                float val = 0.0f;
                if (j >= 0 && j < n_valid) val = sample_to_f32<SampleT>(x[j]);
                sx[i0 + q] = val;
            }
        }
    }
    __syncthreads();

    // ---- stage 1: window, pack z[n] = x[2n] + i x[2n+1], 8-point DFTs over n1 (n = 25 n1 + n2), twiddle
    for (int t = tid; t < LM_FR * 25; t += LM_THREADS) {
        const int f = t / 25, n2 = t - f * 25;
        cpx v[8];
#pragma unroll
        for (int n1 = 0; n1 < 8; ++n1) {
            const int n = 25 * n1 + n2;
            const float2 xv = *reinterpret_cast<const float2*>(&sx[LM_HOP * f + 2 * n]);
            const float2 wv = *reinterpret_cast<const float2*>(&s.win[2 * n]);
            v[n1] = {xv.x * wv.x, xv.y * wv.y};
        }
        dft8(v);
#pragma unroll
        for (int k1 = 0; k1 < 8; ++k1) {
            const float2 w = s.w200[n2 * k1];     // n2*k1 <= 168 < 200
            const cpx y = cmul(v[k1], cpx{w.x, w.y});
            s.z[f][25 * k1 + n2] = make_float2(y.r, y.i);
        }
    }
    __syncthreads();

    // ---- stage 2: 25-point DFT over n2 for each (frame, k1), in registers (5 x 5)
    for (int t = tid; t < LM_FR * 8; t += LM_THREADS) {
        const int f = t >> 3, k1 = t & 7;
        cpx y[25];
#pragma unroll
        for (int j = 0; j < 25; ++j) {
            const float2 q = s.z[f][25 * k1 + j];
            y[j] = {q.x, q.y};
        }
        // n2 = 5a + b ; 5-point DFT over a (stride 5) -> T[c][b] at y[5c + b], times W25^(b c)
#pragma unroll
        for (int bb = 0; bb < 5; ++bb) dft5<5>(&y[bb]);
#pragma unroll
        for (int c = 1; c < 5; ++c)
#pragma unroll
            for (int bb = 1; bb < 5; ++bb)
                y[5 * c + bb] = cmul(y[5 * c + bb], cpx{c_cos25[bb * c], -c_sin25[bb * c]});
        // 5-point DFT over b (stride 1) -> Z2[c + 5d] at y[5c + d]
#pragma unroll
        for (int c = 0; c < 5; ++c) dft5<1>(&y[5 * c]);
#pragma unroll
        for (int j = 0; j < 25; ++j) s.z[f][25 * k1 + j] = make_float2(y[j].r, y[j].i);
    }
    __syncthreads();

    // ---- real-FFT split + power, two bins per task: with E = (Z[k] + conj Z[200-k]) / 2, O = (Z[k] - conj Z[200-k]) / 2i
    // and T = W400^k O:  |X[k]|^2 = |E + T|^2 and |X[200-k]|^2 = |E - T|^2
    for (int t = tid; t < LM_FR * 101; t += LM_THREADS) {
        const int f = t / 101, k = t - f * 101;
        const float2 za = s.z[f][s.zpos[k]];
        const float2 zb = s.z[f][s.zpos[200 - k]];
        const float er = 0.5f * (za.x + zb.x), ei = 0.5f * (za.y - zb.y);
        const float orr = 0.5f * (za.y + zb.y), oi = -0.5f * (za.x - zb.x);
        const float2 w = s.w400[k];
        const float tr = w.x * orr - w.y * oi, ti = w.x * oi + w.y * orr;
        const float ar = er + tr, ai = ei + ti, br = er - tr, bi = ei - ti;
        spw[f][k] = ar * ar + ai * ai;
        spw[f][200 - k] = br * br + bi * bi;        // k = 100 writes the same value twice
    }
    __syncthreads();

    // ---- sparse mel filterbank + log10 clamp; thread -> (frame = tid % FR, mel = tid / FR + (256 / FR) it)
    const int f = tid % LM_FR;
    const int frame = f0 + f;
    float lmax = -INFINITY;
    float* out_b = out + (int64_t)b * n_mel * TW_N_FRAMES;
    for (int m = tid / LM_FR; m < n_mel; m += LM_THREADS / LM_FR) {
        const int st = s.fb_start[m], cnt = s.fb_count[m], off = s.fb_off[m];
        float acc = 0.0f;
        for (int j = 0; j < cnt; ++j) acc = fmaf(s.fbw[off + j], spw[f][st + j], acc);
        const float lv = __log10f(fmaxf(acc, 1e-10f));      // MUFU.LG2 path: |err| ~1e-7, the contract is 1e-4
        if (frame < TW_N_FRAMES) {
            out_b[(int64_t)m * TW_N_FRAMES + frame] = lv;
            lmax = fmaxf(lmax, lv);
        }
    }
    lmax = warp_max(lmax);
    if ((tid & 31) == 0) s.red[tid >> 5] = lmax;
    __syncthreads();
    if (tid == 0) {
        float m = s.red[0];
#pragma unroll
        for (int w = 1; w < LM_THREADS / 32; ++w) m = fmaxf(m, s.red[w]);
        atomic_max_float(&clip_max[b], m);
    }
    }   // tile loop
}

__global__ void logmel_init_max(float* clip_max, int B) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < B) clip_max[i] = -INFINITY;
}

// pass 2: x = (max(x, clipmax - 8) + 4) / 4, float4-vectorised (n_mel*3000 is a multiple of 4)
__global__ void __launch_bounds__(256)
logmel_finalize(float* __restrict__ out, const float* __restrict__ clip_max, int per_clip4) {
    const int b = blockIdx.y;
    const float floor_v = clip_max[b] - 8.0f;
    float4* p = reinterpret_cast<float4*>(out) + (int64_t)b * per_clip4;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < per_clip4; i += gridDim.x * blockDim.x) {
        float4 v = p[i];
        v.x = (fmaxf(v.x, floor_v) + 4.0f) * 0.25f;
        v.y = (fmaxf(v.y, floor_v) + 4.0f) * 0.25f;
        v.z = (fmaxf(v.z, floor_v) + 4.0f) * 0.25f;
        v.w = (fmaxf(v.w, floor_v) + 4.0f) * 0.25f;
        p[i] = v;
    }
}

// ---------------------------------------------------------------------------------------------
// host: tables (double precision, cast to f32 exactly as HF stores its filterbank)
static double hz_to_mel(double f) {
    const double logstep = 27.0 / log(6.4);
    return f >= 1000.0 ? 15.0 + log(f / 1000.0) * logstep : 3.0 * f / 200.0;
}
static double mel_to_hz(double m) {
    const double logstep = log(6.4) / 27.0;
    return m >= 15.0 ? 1000.0 * exp(logstep * (m - 15.0)) : 200.0 * m / 3.0;
}

static int build_bank(tw_ctx* ctx, tw_ctx::MelBank& bank, int n_mel) {
    std::vector<double> ff(n_mel + 2);
    const double m0 = hz_to_mel(0.0), m1 = hz_to_mel(8000.0);
    for (int i = 0; i < n_mel + 2; ++i) ff[i] = mel_to_hz(m0 + (m1 - m0) * (double)i / (double)(n_mel + 1));
    std::vector<int> start(n_mel), count(n_mel), off(n_mel);
    std::vector<float> w;
    for (int m = 0; m < n_mel; ++m) {
        const double enorm = 2.0 / (ff[m + 2] - ff[m]);
        int first = -1, last = -1;
        std::vector<float> row(LM_NBIN);
        for (int k = 0; k < LM_NBIN; ++k) {
            const double fk = 8000.0 * (double)k / 200.0;     // np.linspace(0, 8000, 201)
            const double down = -(ff[m] - fk) / (ff[m + 1] - ff[m]);
            const double up = (ff[m + 2] - fk) / (ff[m + 2] - ff[m + 1]);
            double v = down < up ? down : up;
            if (v < 0.0) v = 0.0;
            row[k] = (float)(v * enorm);
            if (row[k] != 0.0f) {
                if (first < 0) first = k;
                last = k;
            }
        }
        start[m] = first < 0 ? 0 : first;
        count[m] = first < 0 ? 0 : last - first + 1;
        off[m] = (int)w.size();
        for (int k = 0; k < count[m]; ++k) w.push_back(row[start[m] + k]);
    }
    if ((int)w.size() > LM_MAX_NNZ) {
        ctx->set_error(TW_E_INVALID, "mel filterbank too dense");
        return TW_E_INVALID;
    }
    bank.n_mel = n_mel;
    bank.nnz = (int)w.size();
    TW_CUDA_OK(ctx, cudaMalloc(&bank.d_start, n_mel * sizeof(int)));
    TW_CUDA_OK(ctx, cudaMalloc(&bank.d_count, n_mel * sizeof(int)));
    TW_CUDA_OK(ctx, cudaMalloc(&bank.d_offset, n_mel * sizeof(int)));
    TW_CUDA_OK(ctx, cudaMalloc(&bank.d_w, w.size() * sizeof(float)));
    TW_CUDA_OK(ctx, cudaMemcpy(bank.d_start, start.data(), n_mel * sizeof(int), cudaMemcpyHostToDevice));
    TW_CUDA_OK(ctx, cudaMemcpy(bank.d_count, count.data(), n_mel * sizeof(int), cudaMemcpyHostToDevice));
    TW_CUDA_OK(ctx, cudaMemcpy(bank.d_offset, off.data(), n_mel * sizeof(int), cudaMemcpyHostToDevice));
    TW_CUDA_OK(ctx, cudaMemcpy(bank.d_w, w.data(), w.size() * sizeof(float), cudaMemcpyHostToDevice));
    return TW_OK;
}

int logmel_init(tw_ctx* ctx) {
    static LogmelTables h;
    const double PI = 3.14159265358979323846;
    for (int n = 0; n < LM_NFFT; ++n) h.win[n] = (float)(0.5 - 0.5 * cos(2.0 * PI * n / LM_NFFT));
    for (int m = 0; m < 200; ++m) h.w200[m] = make_float2((float)cos(2.0 * PI * m / 200.0), (float)-sin(2.0 * PI * m / 200.0));
    for (int k = 0; k < LM_NBIN; ++k) h.w400[k] = make_float2((float)cos(2.0 * PI * k / 400.0), (float)-sin(2.0 * PI * k / 400.0));
    TW_CUDA_OK(ctx, cudaMemcpyToSymbol(g_tables, &h, sizeof(h)));
    TW_CHECK(build_bank(ctx, ctx->banks[0], 80));
    TW_CHECK(build_bank(ctx, ctx->banks[1], 128));
    TW_CUDA_OK(ctx, cudaFuncSetAttribute(logmel_kernel<int16_t, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(LmSmem<32>)));
    TW_CUDA_OK(ctx, cudaFuncSetAttribute(logmel_kernel<float, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(LmSmem<32>)));
    TW_CUDA_OK(ctx, cudaFuncSetAttribute(logmel_kernel<int16_t, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(LmSmem<16>)));
    TW_CUDA_OK(ctx, cudaFuncSetAttribute(logmel_kernel<float, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(LmSmem<16>)));
    return TW_OK;
}

void logmel_destroy(tw_ctx* ctx) {
    for (auto& b : ctx->banks) {
        cudaFree(b.d_start); cudaFree(b.d_count); cudaFree(b.d_offset); cudaFree(b.d_w);
        b = tw_ctx::MelBank();
    }
    cudaFree(ctx->d_clip_max);
    ctx->d_clip_max = nullptr;
    ctx->clip_max_cap = 0;
}

int logmel_run(tw_ctx* ctx, const void* pcm, int pcm_dtype, int64_t pcm_stride, const int32_t* n_valid, int B, int n_mel,
               float* out, cudaStream_t st, bool finalize, const float** clip_max_out) {
    if (B <= 0) return TW_OK;
    if (n_mel != 80 && n_mel != 128) {
        ctx->set_error(TW_E_INVALID, "tw_logmel: n_mel must be 80 or 128");
        return TW_E_INVALID;
    }
    if (pcm_dtype != TW_I16 && pcm_dtype != TW_F32) {
        ctx->set_error(TW_E_INVALID, "tw_logmel: pcm dtype must be int16 or float32");
        return TW_E_INVALID;
    }
    if (!pcm || !out || pcm_stride <= 0) {
        ctx->set_error(TW_E_INVALID, "tw_logmel: null buffer or bad stride");
        return TW_E_INVALID;
    }
    if (B > ctx->clip_max_cap) {
        if (ctx->d_clip_max) TW_CUDA_OK(ctx, cudaFree(ctx->d_clip_max));
        const int cap = B < 256 ? 256 : B;
        TW_CUDA_OK(ctx, cudaMalloc(&ctx->d_clip_max, cap * sizeof(float)));
        ctx->clip_max_cap = cap;
    }
    const tw_ctx::MelBank& bank = ctx->banks[n_mel == 80 ? 0 : 1];
    logmel_init_max<<<ceil_div(B, 256), 256, 0, st>>>(ctx->d_clip_max, B);
    static const int fr = getenv("TWB200_LM_FR") ? atoi(getenv("TWB200_LM_FR")) : 16;      // tuning knob: frames per tile (16 | 32)
#define TW_LM_LAUNCH(ST, FR)                                                                                                   \
    do {                                                                                                                       \
        const int n_tiles = ceil_div(TW_N_FRAMES, FR) * B;                                                                     \
        const int per_sm = LmCfg<FR>::CTAS;                                                                                    \
        const int grid = n_tiles < per_sm * ctx->sm_count ? n_tiles : per_sm * ctx->sm_count;                                  \
        logmel_kernel<ST, FR><<<grid, LM_THREADS, sizeof(LmSmem<FR>), st>>>((const ST*)pcm, pcm_stride, n_valid, B, n_mel, bank.d_start, \
                                                                             bank.d_count, bank.d_offset, bank.d_w, bank.nnz, out,        \
                                                                             ctx->d_clip_max);                                           \
    } while (0)
    if (pcm_dtype == TW_I16) {
        if (fr == 32) TW_LM_LAUNCH(int16_t, 32); else TW_LM_LAUNCH(int16_t, 16);
    } else {
        if (fr == 32) TW_LM_LAUNCH(float, 32); else TW_LM_LAUNCH(float, 16);
    }
#undef TW_LM_LAUNCH
    if (clip_max_out) *clip_max_out = ctx->d_clip_max;
    if (finalize) {
        const int per_clip4 = n_mel * TW_N_FRAMES / 4;
        dim3 g2(ceil_div(per_clip4, 256 * 4), B);
        logmel_finalize<<<g2, 256, 0, st>>>(out, ctx->d_clip_max, per_clip4);
        ctx->launches += 1;
    }
    ctx->launches += 2;
    TW_CUDA_OK(ctx, cudaGetLastError());
    return TW_OK;
}

}  // namespace tw
