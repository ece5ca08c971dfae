// Shared declarations for libtwb200 (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <utility>

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

// GELU for the bf16 tensor-core epilogues: erf by Abramowitz-Stegun 7.1.26 (|error| <= 1.5e-7, far below bf16
// resolution), one MUFU reciprocal + one MUFU exp2 and ~12 FMAs instead of erff's ~45 instructions.  The fp32 check mode keeps erff.
__device__ __forceinline__ float gelu_erf_fast(float x) {
    const float z = fabsf(x) * 0.70710678118654752440f;
    float t;                                           // MUFU.RCP (1 ulp); __frcp_rn would be a ~60-instruction routine
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.0f)));
    float p = fmaf(1.061405429f, t, -1.453152027f);
    p = fmaf(p, t, 1.421413741f);
    p = fmaf(p, t, -0.284496736f);
    p = fmaf(p, t, 0.254829592f);
    p *= t;
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-z * z * 1.4426950408889634f));
    const float erf_abs = 1.0f - p * e;
    const float erf_v = copysignf(erf_abs, x);
    return 0.5f * x * (1.0f + erf_v);
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

// ---- programmatic dependent launch (PDL) for the decode-step kernel chain.  Kernels launched through launch_k
// while g_pdl is set may start before their predecessor has finished: each of them calls pdl_trigger() on entry
// (lets its own successor be scheduled) and pdl_wait() before touching anything the predecessor produces, so
// barrier/TMEM setup and weight prefetch overlap the predecessor's tail.  Without the launch attribute both
// instructions are no-ops.
extern thread_local bool g_pdl;
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = g_pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
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
