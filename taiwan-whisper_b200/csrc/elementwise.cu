// LayerNorm, conv im2col gathers, token embedding, KV append.
#include "kernels.cuh"

namespace tw {

// ---- LayerNorm over the fp32 residual stream: one warp per row, two-pass (mean, then centred
// variance) in fp32, biased variance, eps 1e-5 (torch nn.LayerNorm; ref flax layers.py:759-815).
template <typename T>
__global__ void __launch_bounds__(256)
layernorm_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                 T* __restrict__ out, int M, int d) {
    pdl_trigger();
    pdl_wait();
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= M) return;
    const float* xr = x + (int64_t)row * d;
    // d <= 1280 here: keep the row in registers (float4 per lane, up to 10 chunks)
    constexpr int MAXC = 10;
    float4 v[MAXC];
    const int nchunk = d >> 2;          // d % 4 == 0
    float s = 0.0f;
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
        const int i = lane + 32 * c;
        if (i < nchunk) {
            v[c] = *reinterpret_cast<const float4*>(xr + 4 * i);
            s += (v[c].x + v[c].y) + (v[c].z + v[c].w);
        }
    }
    const float mean = warp_sum(s) / (float)d;
    float q = 0.0f;
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
        const int i = lane + 32 * c;
        if (i < nchunk) {
            const float a = v[c].x - mean, b = v[c].y - mean, e = v[c].z - mean, f = v[c].w - mean;
            q += (a * a + b * b) + (e * e + f * f);
        }
    }
    const float rstd = rsqrtf(warp_sum(q) / (float)d + 1e-5f);
    T* orow = out + (int64_t)row * d;
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
        const int i = lane + 32 * c;
        if (i < nchunk) {
            const float4 g = *reinterpret_cast<const float4*>(gamma + 4 * i);
            const float4 bb = *reinterpret_cast<const float4*>(beta + 4 * i);
            const float o0 = (v[c].x - mean) * rstd * g.x + bb.x;
            const float o1 = (v[c].y - mean) * rstd * g.y + bb.y;
            const float o2 = (v[c].z - mean) * rstd * g.z + bb.z;
            const float o3 = (v[c].w - mean) * rstd * g.w + bb.w;
            if constexpr (sizeof(T) == 4) {
                *reinterpret_cast<float4*>(reinterpret_cast<float*>(orow) + 4 * i) = make_float4(o0, o1, o2, o3);
            } else {
                __nv_bfloat162 p0 = __floats2bfloat162_rn(o0, o1), p1 = __floats2bfloat162_rn(o2, o3);
                uint2 pk;
                pk.x = *reinterpret_cast<uint32_t*>(&p0);
                pk.y = *reinterpret_cast<uint32_t*>(&p1);
                *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(orow) + 4 * i) = pk;
            }
        }
    }
}

// Decode-step variant (M = batch rows): one CTA per row, one float4 per thread, two block reductions — the
// warp-per-row kernel above would run 8 CTAs and is latency-bound at this size.
template <typename T>
__global__ void __launch_bounds__(320)
layernorm_row_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                     T* __restrict__ out, int d) {
    __shared__ float s_part[2][10];
    pdl_trigger();
    const int row = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = blockDim.x >> 5;
    const bool act = 4 * tid < d;
    // gamma / beta do not depend on the previous kernel: fetch them before the dependency wait
    float4 g = make_float4(0, 0, 0, 0), bb = make_float4(0, 0, 0, 0);
    if (act) {
        g = *reinterpret_cast<const float4*>(gamma + 4 * tid);
        bb = *reinterpret_cast<const float4*>(beta + 4 * tid);
    }
    pdl_wait();
    float4 v = make_float4(0, 0, 0, 0);
    if (act) v = *reinterpret_cast<const float4*>(x + (int64_t)row * d + 4 * tid);
    float s = warp_sum((v.x + v.y) + (v.z + v.w));
    if (lane == 0) s_part[0][warp] = s;
    __syncthreads();
    float tot = 0.0f;
    for (int w = 0; w < nwarp; ++w) tot += s_part[0][w];
    const float mean = tot / (float)d;
    const float a = act ? v.x - mean : 0.0f, b = act ? v.y - mean : 0.0f, c = act ? v.z - mean : 0.0f, e = act ? v.w - mean : 0.0f;
    float q = warp_sum((a * a + b * b) + (c * c + e * e));
    if (lane == 0) s_part[1][warp] = q;
    __syncthreads();
    float qt = 0.0f;
    for (int w = 0; w < nwarp; ++w) qt += s_part[1][w];
    const float rstd = rsqrtf(qt / (float)d + 1e-5f);
    if (act) {
        const float o0 = a * rstd * g.x + bb.x, o1 = b * rstd * g.y + bb.y, o2 = c * rstd * g.z + bb.z, o3 = e * rstd * g.w + bb.w;
        T* orow = out + (int64_t)row * d;
        if constexpr (sizeof(T) == 4) {
            *reinterpret_cast<float4*>(reinterpret_cast<float*>(orow) + 4 * tid) = make_float4(o0, o1, o2, o3);
        } else {
            __nv_bfloat162 p0 = __floats2bfloat162_rn(o0, o1), p1 = __floats2bfloat162_rn(o2, o3);
            uint2 pk;
            pk.x = *reinterpret_cast<uint32_t*>(&p0);
            pk.y = *reinterpret_cast<uint32_t*>(&p1);
            *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(orow) + 4 * tid) = pk;
        }
    }
}

template <typename T>
void layernorm(const float* x, const float* gamma, const float* beta, T* out, int M, int d, cudaStream_t st) {
    if (M <= 0) return;
    if (M <= 512 && d <= 1280) {
        const int threads = ((d / 4 + 31) / 32) * 32;
        launch_k(layernorm_row_kernel<T>, dim3(M), dim3(threads), 0, st, x, gamma, beta, out, d);
        return;
    }
    launch_k(layernorm_kernel<T>, dim3(ceil_div(M, 8)), dim3(256), 0, st, x, gamma, beta, out, M, d);
}
template void layernorm<float>(const float*, const float*, const float*, float*, int, int, cudaStream_t);
template void layernorm<__nv_bfloat16>(const float*, const float*, const float*, __nv_bfloat16*, int, int, cudaStream_t);

// ---- conv1 im2col: out[(b,t)][tap*n_mel + c] = mel[b][c][t + tap - 1]  (0 outside), cast to T —
// the same rounding as the reference's `input_features.to(torch_dtype)`
// (ref: training/run_pseudo_labelling.py:918).  Tile transpose through shared memory:
// reads are contiguous in t, writes contiguous in c.
template <typename T>
__global__ void __launch_bounds__(256)
im2col_conv1_kernel(const float* __restrict__ mel, T* __restrict__ out, int n_mel, const float* __restrict__ clip_max) {
    __shared__ float tile[32][33 + 2];      // [c][t - t0 + 1], halo of 1 each side
    const int b = blockIdx.z, t0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    // clip_max given: mel holds the un-normalised log10 values of the log-mel kernel's first pass; the per-clip floor and
    // the (x + 4) / 4 scaling of logmel_finalize are applied on the way in (the conv padding stays exactly 0)
    const float floor_v = clip_max ? clip_max[b] - 8.0f : 0.0f;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
    const float* src = mel + (int64_t)b * n_mel * TW_N_FRAMES;
    for (int cc = ty; cc < 32; cc += 8) {
        const int c = c0 + cc;
        for (int tt = tx; tt < 34; tt += 32) {
            const int t = t0 + tt - 1;
            float v = 0.0f;
            if (c < n_mel && t >= 0 && t < TW_N_FRAMES) {
                v = src[(int64_t)c * TW_N_FRAMES + t];
                if (clip_max) v = (fmaxf(v, floor_v) + 4.0f) * 0.25f;
            }
            tile[cc][tt] = v;
        }
    }
    __syncthreads();
    const int K = 3 * n_mel;
    for (int tt = ty; tt < 32; tt += 8) {
        const int t = t0 + tt;
        if (t >= TW_N_FRAMES) continue;
        T* orow = out + ((int64_t)b * TW_N_FRAMES + t) * K;
        const int c = c0 + tx;
        if (c < n_mel) {
#pragma unroll
            for (int tap = 0; tap < 3; ++tap) orow[tap * n_mel + c] = from_f32<T>(tile[tx][tt + tap]);
        }
    }
}

template <typename T>
void im2col_conv1(const float* mel, T* out, int B, int n_mel, cudaStream_t st, const float* clip_max) {
    dim3 grid(ceil_div(TW_N_FRAMES, 32), ceil_div(n_mel, 32), B);
    im2col_conv1_kernel<T><<<grid, 256, 0, st>>>(mel, out, n_mel, clip_max);
}
template void im2col_conv1<float>(const float*, float*, int, int, cudaStream_t, const float*);
template void im2col_conv1<__nv_bfloat16>(const float*, __nv_bfloat16*, int, int, cudaStream_t, const float*);

// ---- conv2 im2col (stride 2, pad 1): out[(b,t')][tap*d + c] = h0[(b, 2t'+tap-1)][c]; 16-byte copies
template <typename T>
__global__ void __launch_bounds__(256)
im2col_conv2_kernel(const T* __restrict__ h0, T* __restrict__ out, int d, int64_t total_vec) {
    constexpr int VEC = 16 / sizeof(T);
    const int dv = d / VEC;                        // d % 8 == 0
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total_vec; i += (int64_t)gridDim.x * blockDim.x) {
        const int cv = (int)(i % dv);
        const int64_t r = i / dv;                  // (b*1500 + t')*3 + tap
        const int tap = (int)(r % 3);
        const int64_t bt = r / 3;
        const int tp = (int)(bt % TW_N_CTX);
        const int64_t b = bt / TW_N_CTX;
        const int t = 2 * tp + tap - 1;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (t >= 0 && t < TW_N_FRAMES)
            v = *reinterpret_cast<const uint4*>(h0 + ((int64_t)b * TW_N_FRAMES + t) * d + (int64_t)cv * VEC);
        *reinterpret_cast<uint4*>(out + (bt * 3 + tap) * d + (int64_t)cv * VEC) = v;
    }
}

template <typename T>
void im2col_conv2(const T* h0, T* out, int B, int d, cudaStream_t st) {
    constexpr int VEC = 16 / sizeof(T);
    const int64_t total = (int64_t)B * TW_N_CTX * 3 * (d / VEC);
    int blocks = (int)((total + 255) / 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    im2col_conv2_kernel<T><<<blocks, 256, 0, st>>>(h0, out, d, total);
}
template void im2col_conv2<float>(const float*, float*, int, int, cudaStream_t);
template void im2col_conv2<__nv_bfloat16>(const __nv_bfloat16*, __nv_bfloat16*, int, int, cudaStream_t);

// ---- decoder input: x[b] = E[tok[b]] + P[pos]  (HF WhisperDecoder.forward :738-760)
template <typename T>
__global__ void embed_kernel(const int32_t* __restrict__ tok, const T* __restrict__ E, const T* __restrict__ P,
                             const int32_t* __restrict__ d_step, float* __restrict__ x, int d) {
    pdl_trigger();
    pdl_wait();
    const int b = blockIdx.x;
    const int t = tok[b];
    const int pos = d_step[0];
    for (int i = threadIdx.x; i < d; i += blockDim.x)
        x[(int64_t)b * d + i] = to_f32(E[(int64_t)t * d + i]) + to_f32(P[(int64_t)pos * d + i]);
}
template <typename T>
void embed_tokens(const int32_t* tok, const T* E, const T* P, const int32_t* d_step, float* x, int B, int d, cudaStream_t st) {
    launch_k(embed_kernel<T>, dim3(B), dim3(256), 0, st, tok, E, P, d_step, x, d);
}
template void embed_tokens<float>(const int32_t*, const float*, const float*, const int32_t*, float*, int, int, cudaStream_t);
template void embed_tokens<__nv_bfloat16>(const int32_t*, const __nv_bfloat16*, const __nv_bfloat16*, const int32_t*, float*, int, int,
                                          cudaStream_t);

// ---- full token sequences (teacher-forced decoder pass): row = clip * Tn + position; ids outside the vocabulary (the -100 of
// HF label tensors must be replaced by the caller) are clamped so that the gather stays in bounds
template <typename T>
__global__ void embed_seq_kernel(const int32_t* __restrict__ tok, const T* __restrict__ E, const T* __restrict__ P,
                                 float* __restrict__ x, int Tn, int d, int vocab) {
    const int row = blockIdx.x;
    int t = tok[row];
    t = t < 0 ? 0 : (t >= vocab ? vocab - 1 : t);
    const int pos = row % Tn;
    for (int i = threadIdx.x; i < d; i += blockDim.x)
        x[(int64_t)row * d + i] = to_f32(E[(int64_t)t * d + i]) + to_f32(P[(int64_t)pos * d + i]);
}
template <typename T>
void embed_tokens_seq(const int32_t* tok, const T* E, const T* P, float* x, int B, int Tn, int d, int vocab, cudaStream_t st) {
    embed_seq_kernel<T><<<B * Tn, 256, 0, st>>>(tok, E, P, x, Tn, d, vocab);
}
template void embed_tokens_seq<float>(const int32_t*, const float*, const float*, float*, int, int, int, int, cudaStream_t);
template void embed_tokens_seq<__nv_bfloat16>(const int32_t*, const __nv_bfloat16*, const __nv_bfloat16*, float*, int, int, int, int,
                                              cudaStream_t);

// ---- self-attention KV cache append: cache[b][pos][0:2d] = qkv[b][d:3d]
template <typename T>
__global__ void kv_append_kernel(const T* __restrict__ qkv, T* __restrict__ cache, const int32_t* __restrict__ d_step, int d,
                                 int max_len, const int32_t* __restrict__ page_table, int pt_stride) {
    pdl_trigger();
    pdl_wait();
    const int b = blockIdx.x;
    const int pos = d_step[0];
    const T* src = qkv + (int64_t)b * 3 * d + d;
    T* dst = cache + (page_table ? kv_page_row(page_table, pt_stride, b, pos) : (int64_t)b * max_len + pos) * 2 * d;
    for (int i = threadIdx.x; i < 2 * d; i += blockDim.x) dst[i] = src[i];
}
template <typename T>
void kv_append(const T* qkv, T* cache, const int32_t* d_step, int B, int d, int max_len, cudaStream_t st, const int32_t* page_table,
               int pt_stride) {
    launch_k(kv_append_kernel<T>, dim3(B), dim3(256), 0, st, qkv, cache, d_step, d, max_len, page_table, pt_stride);
}
template void kv_append<float>(const float*, float*, const int32_t*, int, int, int, cudaStream_t, const int32_t*, int);
template void kv_append<__nv_bfloat16>(const __nv_bfloat16*, __nv_bfloat16*, const int32_t*, int, int, int, cudaStream_t, const int32_t*,
                                       int);

// one warp: position += 1; active[] = ids of the clips whose finished flag is clear, in clip order (ballot + popc prefix)
__global__ void advance_step_kernel(int32_t* d_step, DecodeState S, int B) {
    pdl_trigger();
    pdl_wait();
    const int lane = threadIdx.x;
    if (lane == 0) d_step[0] += 1;
    int n = 0;
    for (int base = 0; base < B; base += 32) {
        const int b = base + lane;
        const bool live = b < B && S.finished[b] == 0;
        const unsigned m = __ballot_sync(0xffffffffu, live);
        if (live) S.active[n + __popc(m & ((1u << lane) - 1u))] = b;
        n += __popc(m);
    }
    if (lane == 0) *S.n_active = n;
}
void advance_step(int32_t* d_step, const DecodeState& S, int B, cudaStream_t st) {
    launch_k(advance_step_kernel, dim3(1), dim3(32), 0, st, d_step, S, B);
}

__global__ void copy_f32_kernel(const float4* __restrict__ s, float4* __restrict__ d, int64_t n4) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) d[i] = s[i];
}
void copy_f32(const float* src, float* dst, int64_t n, cudaStream_t st) {
    const int64_t n4 = n / 4;     // callers pass multiples of 4
    int blocks = (int)((n4 + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (blocks > 0) copy_f32_kernel<<<blocks, 256, 0, st>>>(reinterpret_cast<const float4*>(src), reinterpret_cast<float4*>(dst), n4);
}

}  // namespace tw
