// Shared declarations for libtwb200 (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "../../include/twb200.h"

namespace tw {

struct Error {
    int code;
    std::string msg;
};

// Every internal entry point returns 0 or a TW_E_* code and fills ctx->err.
#define TW_CUDA_OK(ctx, expr)                                                                   \
    do {                                                                                        \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess) {                                                                \
            (ctx)->set_error(TW_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));    \
            return TW_E_CUDA;                                                                   \
        }                                                                                       \
    } while (0)

#define TW_CHECK(call)               \
    do {                             \
        int _r = (call);             \
        if (_r != TW_OK) return _r;  \
    } while (0)

template <typename T> struct TypeTag;
template <> struct TypeTag<float> { static constexpr int id = TW_F32; };
template <> struct TypeTag<__nv_bfloat16> { static constexpr int id = TW_BF16; };

__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ float gelu_erf(float x) {
    // exact GELU: 0.5 x (1 + erf(x / sqrt 2))   (HF activation "gelu"; ref flax :977)
    return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

}  // namespace tw

// ---------------------------------------------------------------------------------------------
// The context: one per device per host thread (not thread-safe, no global state).
struct tw_ctx {
    int device = 0;
    int sm_count = 148;
    std::string err;
    int err_code = 0;
    // log-mel tables (device): sparse slaney filterbank per n_mel in {80,128}
    struct MelBank {
        int n_mel = 0;
        int nnz = 0;
        int* d_start = nullptr;   // [n_mel] first fft bin
        int* d_count = nullptr;   // [n_mel]
        int* d_offset = nullptr;  // [n_mel] offset into weights
        float* d_w = nullptr;     // [nnz]
    } banks[2];
    float* d_clip_max = nullptr;  // [cap] scratch for the per-clip max
    int clip_max_cap = 0;
    uint64_t launches = 0;        // kernels launched through this ctx (bench's gpu_launches)

    void set_error(int code, const std::string& m) {
        err_code = code;
        err = m;
    }
};
