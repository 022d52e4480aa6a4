// Throughput of fma.rn.f32x2 vs fma.rn.f32 on sm_100a (is the packed form full rate?)
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float *out, int iters) {
    float a0 = threadIdx.x * 1e-3f, a1 = a0 + 1.f, a2 = a0 + 2.f, a3 = a0 + 3.f, a4 = a0 + 4.f, a5 = a0 + 5.f, a6 = a0 + 6.f, a7 = a0 + 7.f;
    float b0 = 1.0001f, b1 = 0.9999f;
    unsigned long long p0, p1, p2, p3, q, c;
    asm("mov.b64 %0, {%1, %2};" : "=l"(p0) : "f"(a0), "f"(a1));
    asm("mov.b64 %0, {%1, %2};" : "=l"(p1) : "f"(a2), "f"(a3));
    asm("mov.b64 %0, {%1, %2};" : "=l"(p2) : "f"(a4), "f"(a5));
    asm("mov.b64 %0, {%1, %2};" : "=l"(p3) : "f"(a6), "f"(a7));
    asm("mov.b64 %0, {%1, %2};" : "=l"(q) : "f"(b0), "f"(b1));
    asm("mov.b64 %0, {%1, %2};" : "=l"(c) : "f"(1e-7f), "f"(2e-7f));
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            if (MODE == 0) {
                a0 = fmaf(a0, b0, 1e-7f); a1 = fmaf(a1, b1, 1e-7f); a2 = fmaf(a2, b0, 1e-7f); a3 = fmaf(a3, b1, 1e-7f);
                a4 = fmaf(a4, b0, 1e-7f); a5 = fmaf(a5, b1, 1e-7f); a6 = fmaf(a6, b0, 1e-7f); a7 = fmaf(a7, b1, 1e-7f);
            } else {
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p0) : "l"(q), "l"(c));
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p1) : "l"(q), "l"(c));
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p2) : "l"(q), "l"(c));
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p3) : "l"(q), "l"(c));
            }
        }
    }
    if (MODE == 0) out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    else { float x, y; asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(p0 ^ p1 ^ p2 ^ p3)); out[blockIdx.x * blockDim.x + threadIdx.x] = x + y; }
}
int main() {
    float *out; cudaMalloc(&out, 148 * 8 * 512 * sizeof(float));
    const int iters = 4096;
    for (int mode = 0; mode < 2; ++mode) {
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0);
            if (mode == 0) k<0><<<148 * 4, 512>>>(out, iters); else k<1><<<148 * 4, 512>>>(out, iters);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            double fma = 148.0 * 4 * 512 * iters * 16 * 8;
            if (rep == 2) printf("%s: %.3f ms, %.1f TFMA/s (%.1f TFLOP/s)\n", mode == 0 ? "fma.f32 x8" : "fma.f32x2 x4", ms, fma / ms * 1e-9, 2 * fma / ms * 1e-9);
        }
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
