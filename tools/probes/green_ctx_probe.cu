// Probe: do CUDA green contexts (driver API, SM partitions) confine kernels launched through the RUNTIME API on a stream created with
// cuGreenCtxStreamCreate — plain launches, programmatic dependent launches and replayed CUDA graphs?
//   nvcc -gencode arch=compute_100a,code=sm_100a -o green_ctx_probe green_ctx_probe.cu && ./green_ctx_probe [n_sms_small]
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <set>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 1; } } while (0)
#define DK(x) do { CUresult r = (x); if (r != CUDA_SUCCESS) { printf("driver error %d at %s:%d (%s)\n", (int)r, __FILE__, __LINE__, #x); return 1; } } while (0)

__global__ void smid_kernel(unsigned* out, long long spin_ns) {
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    if (threadIdx.x == 0) out[blockIdx.x] = smid;
    long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    do { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1)); } while (t1 - t0 < spin_ns);
}

template <typename F> static F entry(const char* name) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint(name, &fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) return nullptr;
    return reinterpret_cast<F>(fn);
}

static int report(const char* what, unsigned* d_out, int n) {
    std::vector<unsigned> h(n);
    cudaMemcpy(h.data(), d_out, n * sizeof(unsigned), cudaMemcpyDeviceToHost);
    std::set<unsigned> s(h.begin(), h.end());
    printf("%-40s %d CTAs ran on %zu distinct SMs: min %u max %u\n", what, n, s.size(), *s.begin(), *s.rbegin());
    return (int)s.size();
}

int main(int argc, char** argv) {
    const int n_small = argc > 1 ? atoi(argv[1]) : 24;
    CK(cudaSetDevice(0));
    CK(cudaFree(0));
    auto pGetRes = entry<CUresult (*)(CUdevice, CUdevResource*, CUdevResourceType)>("cuDeviceGetDevResource");
    auto pSplit = entry<CUresult (*)(CUdevResource*, unsigned int*, const CUdevResource*, CUdevResource*, unsigned int, unsigned int)>("cuDevSmResourceSplitByCount");
    auto pDesc = entry<CUresult (*)(CUdevResourceDesc*, CUdevResource*, unsigned int)>("cuDevResourceGenerateDesc");
    auto pCreate = entry<CUresult (*)(CUgreenCtx*, CUdevResourceDesc, CUdevice, unsigned int)>("cuGreenCtxCreate");
    auto pStream = entry<CUresult (*)(CUstream*, CUgreenCtx, unsigned int, int)>("cuGreenCtxStreamCreate");
    if (!pGetRes || !pSplit || !pDesc || !pCreate || !pStream) { printf("green context entry points missing\n"); return 1; }
    CUdevResource all, small, rest;
    DK(pGetRes(0, &all, CU_DEV_RESOURCE_TYPE_SM));
    printf("device SM resource: %u SMs\n", all.sm.smCount);
    unsigned nb = 1;
    DK(pSplit(&small, &nb, &all, &rest, 0, (unsigned)n_small));
    printf("split: %u group(s) of %u SMs, remaining %u SMs\n", nb, small.sm.smCount, rest.sm.smCount);
    CUdevResourceDesc dsmall, drest;
    DK(pDesc(&dsmall, &small, 1));
    DK(pDesc(&drest, &rest, 1));
    CUgreenCtx gsmall, grest;
    DK(pCreate(&gsmall, dsmall, 0, CU_GREEN_CTX_DEFAULT_STREAM));
    DK(pCreate(&grest, drest, 0, CU_GREEN_CTX_DEFAULT_STREAM));
    CUstream ss, sr, sr2;
    DK(pStream(&ss, gsmall, CU_STREAM_NON_BLOCKING, 0));
    DK(pStream(&sr, grest, CU_STREAM_NON_BLOCKING, 0));
    DK(pStream(&sr2, grest, CU_STREAM_NON_BLOCKING, 0));
    unsigned *o1, *o2;
    const int N = 1024;
    CK(cudaMalloc(&o1, N * 4));
    CK(cudaMalloc(&o2, N * 4));
    // plain runtime launches
    smid_kernel<<<N, 64, 0, (cudaStream_t)ss>>>(o1, 20000);
    smid_kernel<<<N, 64, 0, (cudaStream_t)sr>>>(o2, 20000);
    CK(cudaDeviceSynchronize());
    report("runtime launch, small partition:", o1, N);
    report("runtime launch, rest partition:", o2, N);
    // concurrency: 1 ms on each partition at once
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    smid_kernel<<<small.sm.smCount, 64, 0, (cudaStream_t)ss>>>(o1, 1000);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0, (cudaStream_t)sr));
    smid_kernel<<<small.sm.smCount * 8, 1024, 0, (cudaStream_t)ss>>>(o1, 1000000);      // fills the small partition for ~4 ms (2 CTAs / SM)
    smid_kernel<<<rest.sm.smCount, 64, 0, (cudaStream_t)sr>>>(o2, 1000000);
    CK(cudaEventRecord(e1, (cudaStream_t)sr));
    CK(cudaDeviceSynchronize());
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    printf("1 ms kernel on the rest partition while the small one is saturated: %.2f ms (concurrent if ~1)\n", ms);
    // launch with the programmatic-serialization attribute
    {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(N); cfg.blockDim = dim3(64); cfg.stream = (cudaStream_t)sr;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        CK(cudaLaunchKernelEx(&cfg, smid_kernel, o2, (long long)20000));
        CK(cudaDeviceSynchronize());
        report("PDL launch, rest partition:", o2, N);
    }
    // graph captured on one green-context stream, replayed on another stream of the same green context
    {
        cudaGraph_t g;
        cudaGraphExec_t ge;
        CK(cudaStreamBeginCapture((cudaStream_t)sr2, cudaStreamCaptureModeThreadLocal));
        smid_kernel<<<N, 64, 0, (cudaStream_t)sr2>>>(o2, 20000);
        CK(cudaStreamEndCapture((cudaStream_t)sr2, &g));
        CK(cudaGraphInstantiate(&ge, g, 0));
        CK(cudaMemset(o2, 0xff, N * 4));
        CK(cudaGraphLaunch(ge, (cudaStream_t)sr));
        CK(cudaDeviceSynchronize());
        report("graph (captured + launched in rest):", o2, N);
        // the same graph launched on the SMALL partition's stream
        CK(cudaMemset(o2, 0xff, N * 4));
        cudaError_t e = cudaGraphLaunch(ge, (cudaStream_t)ss);
        if (e != cudaSuccess) printf("graph captured in rest, launched in small: %s\n", cudaGetErrorString(e));
        else { CK(cudaDeviceSynchronize()); report("graph (captured rest, launched small):", o2, N); }
        cudaGetLastError();
        // graph captured on an ordinary stream, launched into the rest partition
        cudaStream_t plain;
        CK(cudaStreamCreateWithFlags(&plain, cudaStreamNonBlocking));
        CK(cudaStreamBeginCapture(plain, cudaStreamCaptureModeThreadLocal));
        smid_kernel<<<N, 64, 0, plain>>>(o2, 20000);
        CK(cudaStreamEndCapture(plain, &g));
        CK(cudaGraphInstantiate(&ge, g, 0));
        CK(cudaMemset(o2, 0xff, N * 4));
        e = cudaGraphLaunch(ge, (cudaStream_t)sr);
        if (e != cudaSuccess) printf("graph captured on a plain stream, launched in rest: %s\n", cudaGetErrorString(e));
        else { CK(cudaDeviceSynchronize()); report("graph (captured plain, launched rest):", o2, N); }
    }
    // memcpy + events across partitions
    CK(cudaEventRecord(e0, (cudaStream_t)ss));
    CK(cudaStreamWaitEvent((cudaStream_t)sr, e0, 0));
    CK(cudaDeviceSynchronize());
    printf("cross-partition event wait: ok\n");
    return 0;
}
