// Probe: which cuSOLVER symmetric eigensolver is fastest for the CMA-ES decomposition shape (n = 1000,
// fp64, vectors)?  Dsyevd (what the engine uses) vs Xsyevd vs XsyevBatched(batch=1) vs Dsyevj.
// Build: nvcc -O2 -gencode arch=compute_100a,code=sm_100a tools/eig_probe.cu -lcusolver -o tools/bin/eig_probe
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include <cusolverDn.h>

#define CK(x) do { auto r_ = (x); if (r_ != 0) { printf("FAIL %s -> %d line %d\n", #x, (int) r_, __LINE__); exit(1); } } while (0)

int main(int argc, char **argv) {
    const int n = argc > 1 ? atoi(argv[1]) : 1000;
    const int reps = 5;
    std::vector<double> A((size_t) n * n);
    srand(1);
    // SPD with a CMA-like spectrum: I + low-amplitude symmetric noise + a few large directions
    for (int i = 0; i < n; i++)
        for (int j = 0; j <= i; j++) {
            double v = 0.01 * ((rand() / (double) RAND_MAX) - 0.5);
            if (i == j) v += 1.0 + (i % 7) * 0.3;
            A[(size_t) i * n + j] = A[(size_t) j * n + i] = v;
        }
    double *dA, *dA0, *dW;
    int *dinfo;
    CK(cudaMalloc(&dA, sizeof(double) * n * n)); CK(cudaMalloc(&dA0, sizeof(double) * n * n)); CK(cudaMalloc(&dW, sizeof(double) * n));
    CK(cudaMalloc(&dinfo, 4));
    CK(cudaMemcpy(dA0, A.data(), sizeof(double) * n * n, cudaMemcpyHostToDevice));
    cusolverDnHandle_t h; CK(cusolverDnCreate(&h));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms;
    // ---- Dsyevd
    {
        int lwork = 0; CK(cusolverDnDsyevd_bufferSize(h, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n, dA, n, dW, &lwork));
        double *work; CK(cudaMalloc(&work, sizeof(double) * lwork));
        for (int r = 0; r < reps; r++) {
            CK(cudaMemcpy(dA, dA0, sizeof(double) * n * n, cudaMemcpyDeviceToDevice));
            cudaEventRecord(e0);
            CK(cusolverDnDsyevd(h, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n, dA, n, dW, work, lwork, dinfo));
            cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
            printf("Dsyevd         n=%d rep %d: %.3f ms\n", n, r, ms);
        }
        cudaFree(work);
    }
    cusolverDnParams_t params; CK(cusolverDnCreateParams(&params));
    // ---- Xsyevd
    {
        size_t wd = 0, wh = 0;
        CK(cusolverDnXsyevd_bufferSize(h, params, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n, CUDA_R_64F, dA, n, CUDA_R_64F, dW, CUDA_R_64F, &wd, &wh));
        void *bd; CK(cudaMalloc(&bd, wd)); std::vector<char> bh(wh + 1);
        for (int r = 0; r < reps; r++) {
            CK(cudaMemcpy(dA, dA0, sizeof(double) * n * n, cudaMemcpyDeviceToDevice));
            cudaEventRecord(e0);
            CK(cusolverDnXsyevd(h, params, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n, CUDA_R_64F, dA, n, CUDA_R_64F, dW, CUDA_R_64F, bd, wd, bh.data(), wh, dinfo));
            cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
            printf("Xsyevd         n=%d rep %d: %.3f ms\n", n, r, ms);
        }
        cudaFree(bd);
    }
    // ---- XsyevBatched, batch 1
    {
        size_t wd = 0, wh = 0;
        cusolverStatus_t st = cusolverDnXsyevBatched_bufferSize(h, params, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n, CUDA_R_64F, dA, n, CUDA_R_64F, dW, CUDA_R_64F, &wd, &wh, 1);
        if (st != 0) printf("XsyevBatched_bufferSize -> %d (unsupported)\n", (int) st);
        else {
            void *bd; CK(cudaMalloc(&bd, wd)); std::vector<char> bh(wh + 1);
            for (int r = 0; r < reps; r++) {
                CK(cudaMemcpy(dA, dA0, sizeof(double) * n * n, cudaMemcpyDeviceToDevice));
                cudaEventRecord(e0);
                st = cusolverDnXsyevBatched(h, params, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n, CUDA_R_64F, dA, n, CUDA_R_64F, dW, CUDA_R_64F, bd, wd, bh.data(), wh, dinfo, 1);
                cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
                printf("XsyevBatched   n=%d rep %d: %.3f ms (status %d)\n", n, r, ms, (int) st);
            }
            cudaFree(bd);
        }
    }
    // ---- Dsyevj
    {
        syevjInfo_t ji; CK(cusolverDnCreateSyevjInfo(&ji));
        CK(cusolverDnXsyevjSetTolerance(ji, 1e-14)); CK(cusolverDnXsyevjSetMaxSweeps(ji, 30));
        int lwork = 0; CK(cusolverDnDsyevj_bufferSize(h, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n, dA, n, dW, &lwork, ji));
        double *work; CK(cudaMalloc(&work, sizeof(double) * lwork));
        for (int r = 0; r < 2; r++) {
            CK(cudaMemcpy(dA, dA0, sizeof(double) * n * n, cudaMemcpyDeviceToDevice));
            cudaEventRecord(e0);
            CK(cusolverDnDsyevj(h, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n, dA, n, dW, work, lwork, dinfo, ji));
            cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
            int sweeps = 0; cusolverDnXsyevjGetSweeps(h, ji, &sweeps);
            printf("Dsyevj         n=%d rep %d: %.3f ms (%d sweeps)\n", n, r, ms, sweeps);
        }
        cudaFree(work);
    }
    return 0;
}
