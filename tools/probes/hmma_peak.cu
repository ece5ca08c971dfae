// probe: peak of the legacy mma.sync.m16n8k16 bf16 path on B200 (needed to size the "absorbed" cross-attention kernel)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(float* out, int iters) {
    float c[8][4];
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
    unsigned a0 = threadIdx.x, a1 = a0 * 3, a2 = a0 * 5, a3 = a0 * 7, b0 = a0 * 11, b1 = a0 * 13;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
    float s = 0; for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    float* d; cudaMalloc(&d, 148 * 8 * 1024 * 4);
    for (int warps = 4; warps <= 32; warps *= 2) {
        int iters = 20000;
        k<<<148, warps * 32>>>(d, 10); cudaDeviceSynchronize();
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0); k<<<148, warps * 32>>>(d, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double flops = 148.0 * warps * iters * 8 * (2.0 * 16 * 8 * 16);
        printf("warps/SM %d: %.1f TFLOP/s\n", warps, flops / ms / 1e9);
    }
    return 0;
}
