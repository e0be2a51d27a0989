// cuSOLVER syevj on matrices like B^T C B of consecutive CMA-ES generations: diag(1 + spread) + eps * symmetric noise.
// Build: nvcc -O2 -gencode arch=compute_100a,code=sm_100a tools/eig_jacobi_probe.cu -lcusolver -o tools/bin/eig_jacobi_probe
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include <cuda_runtime.h>
#include <cusolverDn.h>
#define CK(x) do { auto r_ = (x); if (r_ != 0) { printf("FAIL %s -> %d line %d\n", #x, (int) r_, __LINE__); exit(1); } } while (0)
int main(int argc, char **argv) {
    const int n = argc > 1 ? atoi(argv[1]) : 1000;
    cusolverDnHandle_t h; CK(cusolverDnCreate(&h));
    double *dA, *dW; int *dinfo;
    CK(cudaMalloc(&dA, 8ull * n * n)); CK(cudaMalloc(&dW, 8ull * n)); CK(cudaMalloc(&dinfo, 4));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const double spreads[] = { 1e-2, 1e-3, 1e-1 }, epss[] = { 7e-5, 1e-6, 1e-3 };
    for (double spread : spreads) for (double eps : epss) for (double tol : { 1e-14, 1e-10 }) {
        std::vector<double> A((size_t) n * n);
        srand(1);
        for (int i = 0; i < n; i++) for (int j = 0; j <= i; j++) {
            double v = eps * ((rand() / (double) RAND_MAX) - 0.5);
            if (i == j) v = 1.0 + spread * (rand() / (double) RAND_MAX);
            A[(size_t) i * n + j] = A[(size_t) j * n + i] = v;
        }
        syevjInfo_t ji; CK(cusolverDnCreateSyevjInfo(&ji));
        CK(cusolverDnXsyevjSetTolerance(ji, tol)); CK(cusolverDnXsyevjSetMaxSweeps(ji, 30));
        int lwork = 0; CK(cusolverDnDsyevj_bufferSize(h, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n, dA, n, dW, &lwork, ji));
        double *work; CK(cudaMalloc(&work, 8ull * lwork));
        float ms = 0;
        for (int r = 0; r < 2; r++) {
            CK(cudaMemcpy(dA, A.data(), 8ull * n * n, cudaMemcpyHostToDevice));
            cudaEventRecord(e0);
            CK(cusolverDnDsyevj(h, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n, dA, n, dW, work, lwork, dinfo, ji));
            cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
        }
        int sweeps = 0; cusolverDnXsyevjGetSweeps(h, ji, &sweeps);
        double res = 0; cusolverDnXsyevjGetResidual(h, ji, &res);
        printf("spread %.0e offdiag %.0e tol %.0e: %.3f ms, %d sweeps, residual %.2e\n", spread, eps, tol, ms, sweeps, res);
        cudaFree(work); cusolverDnDestroySyevjInfo(ji);
    }
    return 0;
}
