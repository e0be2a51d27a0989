// Probe: vector fp64 issue rate on this GPU (DFMA and separately rounded DMUL + DADD), lanes per clock per SM.
// Build: nvcc -O2 -gencode arch=compute_100a,code=sm_100a tools/fp64_rate.cu -o tools/bin/fp64_rate
#include <cstdio>
#include <cuda_runtime.h>

template<int MODE>
__global__ void __launch_bounds__(1024) rate_kernel(double *out, int iters, double a, double b) {
    double x[8];
#pragma unroll
    for (int i = 0; i < 8; i++) x[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (MODE == 0) x[i] = fma(x[i], a, b);
            else x[i] = __dadd_rn(__dmul_rn(x[i], a), b);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += x[i];
    if (s == 12345.678) out[0] = s;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    double *d; cudaMalloc(&d, 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    for (int mode = 0; mode < 2; mode++) {
        for (int rep = 0; rep < 2; rep++) {
            cudaEventRecord(e0);
            if (mode == 0) rate_kernel<0><<<p.multiProcessorCount * 2, 1024>>>(d, iters, 1.0000001, 1e-9);
            else rate_kernel<1><<<p.multiProcessorCount * 2, 1024>>>(d, iters, 1.0000001, 1e-9);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            const double ops = (double) p.multiProcessorCount * 2 * 1024 * iters * 8 * (mode == 0 ? 1 : 2);
            int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
            printf("%s: %.3f ms, %.2f T instr-lanes/s, %.1f lanes/clk/SM at %d MHz (%s)\n", mode == 0 ? "DFMA      " : "DMUL+DADD ", ms,
                   ops / ms / 1e9, ops / (ms * 1e-3) / p.multiProcessorCount / (clk * 1e3), clk / 1000, p.name);
        }
    }
    return 0;
}
