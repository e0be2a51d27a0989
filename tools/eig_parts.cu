// Where do cuSOLVER's 12 ms at n = 1000 go?  Times Dsytrd, Dormtr and Dsyevd (vectors / no vectors) separately.
// Build: nvcc -O2 -gencode arch=compute_100a,code=sm_100a tools/eig_parts.cu -lcusolver -o tools/bin/eig_parts
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include <cusolverDn.h>
#define CK(x) do { auto r_ = (x); if (r_ != 0) { printf("FAIL %s -> %d line %d\n", #x, (int) r_, __LINE__); exit(1); } } while (0)
int main(int argc, char **argv) {
    const int n = argc > 1 ? atoi(argv[1]) : 1000;
    std::vector<double> A((size_t) n * n);
    srand(1);
    for (int i = 0; i < n; i++) for (int j = 0; j <= i; j++) {
        double v = 0.01 * ((rand() / (double) RAND_MAX) - 0.5);
        if (i == j) v += 1.0 + (i % 7) * 0.3;
        A[(size_t) i * n + j] = A[(size_t) j * n + i] = v;
    }
    double *dA, *dA0, *dW, *dE, *dTau, *dQ;
    int *dinfo;
    CK(cudaMalloc(&dA, 8ull * n * n)); CK(cudaMalloc(&dA0, 8ull * n * n)); CK(cudaMalloc(&dQ, 8ull * n * n));
    CK(cudaMalloc(&dW, 8ull * n)); CK(cudaMalloc(&dE, 8ull * n)); CK(cudaMalloc(&dTau, 8ull * n)); CK(cudaMalloc(&dinfo, 4));
    CK(cudaMemcpy(dA0, A.data(), 8ull * n * n, cudaMemcpyHostToDevice));
    cusolverDnHandle_t h; CK(cusolverDnCreate(&h));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms;
    int lw1 = 0, lw2 = 0, lw3 = 0;
    CK(cusolverDnDsytrd_bufferSize(h, CUBLAS_FILL_MODE_LOWER, n, dA, n, dW, dE, dTau, &lw1));
    CK(cusolverDnDormtr_bufferSize(h, CUBLAS_SIDE_LEFT, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, n, n, dA, n, dTau, dQ, n, &lw2));
    CK(cusolverDnDsyevd_bufferSize(h, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n, dA, n, dW, &lw3));
    int lw = lw1 > lw2 ? lw1 : lw2; if (lw3 > lw) lw = lw3;
    double *work; CK(cudaMalloc(&work, 8ull * lw));
    for (int r = 0; r < 4; r++) {
        CK(cudaMemcpy(dA, dA0, 8ull * n * n, cudaMemcpyDeviceToDevice));
        cudaEventRecord(e0);
        CK(cusolverDnDsytrd(h, CUBLAS_FILL_MODE_LOWER, n, dA, n, dW, dE, dTau, work, lw, dinfo));
        cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
        printf("Dsytrd  n=%d rep %d: %.3f ms\n", n, r, ms);
        CK(cudaMemset(dQ, 0, 8ull * n * n));
        cudaEventRecord(e0);
        CK(cusolverDnDormtr(h, CUBLAS_SIDE_LEFT, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, n, n, dA, n, dTau, dQ, n, work, lw, dinfo));
        cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
        printf("Dormtr  n=%d rep %d: %.3f ms\n", n, r, ms);
        for (int mode = 0; mode < 2; mode++) {
            CK(cudaMemcpy(dA, dA0, 8ull * n * n, cudaMemcpyDeviceToDevice));
            cudaEventRecord(e0);
            CK(cusolverDnDsyevd(h, mode ? CUSOLVER_EIG_MODE_VECTOR : CUSOLVER_EIG_MODE_NOVECTOR, CUBLAS_FILL_MODE_LOWER, n, dA, n, dW, work, lw, dinfo));
            cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
            printf("Dsyevd %s n=%d rep %d: %.3f ms\n", mode ? "vectors  " : "novectors", n, r, ms);
        }
    }
    return 0;
}
