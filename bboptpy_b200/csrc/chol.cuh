// Cholesky-CMA-ES on the device (reference: cma/cholesky_cmaes.cpp, Krause et al. 2016).
//
// The reference keeps a lower-triangular factor A (C = A A^T), samples x = m + sigma A z (:120-135) and, per
// generation, applies mu + 1 sequential rank-one updates to sqrt(1 - c1 - cmu) A (:85-90, rank1Update :163-177):
// O(mu n^2) scalar work with a dependency chain through every column -- 8 GFLOP of latency at n = 1000, lambda = 4096.
// The result of that chain is THE Cholesky factor (positive diagonal, hence unique) of
//     M = (1 - c1 - cmu) A A^T + c1 pc pc^T + cmu sum_{i < mu} w_i y_i y_i^T,   y_i = (x_i - m_new) / sigma
// so the device forms M with ONE tensor-core SYRK over the stacked operand [A^T ; Y] (n + mu rows) and factors it with
// a blocked right-looking Cholesky whose trailing updates run on the same DMMA kernel.  Same factor to rounding
// (parity at 1e-10, tests/test_gpu_chol.py); quirks kept: the y_i are the FIRST mu candidates by index, not by rank,
// against the NEW mean; p_sigma uses the OLD factor (:93-98); there is no hsig.
//
// Layout: At = A^T lives in rows [0, n) of the stacked operand Ystk (row j = column j of A): the k-major form both the
// sampling GEMM (B operand) and the SYRK want; rows [n, n + mu) hold the steps y_i.
#pragma once

#include "common.cuh"

namespace bb {

constexpr int CH_NB = 64; // panel width

// pc <- (1 - cc) pc + ccc dmean   (cholesky_cmaes.cpp:76-79; no hsig), and the SYRK coefficients:
// coef[k] = acoeff^2 for the n rows of A^T, cmu w_i for the mu steps, 0 in the pad
__global__ void __launch_bounds__(256) chol_paths_kernel(int n, int mu, double *__restrict__ pc, const double *__restrict__ dm,
        double cc, double ccc, double acoeff, double cmu, const double *__restrict__ w, double *__restrict__ coef, int coef_len) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) pc[i] = __dadd_rn(__dmul_rn(1.0 - cc, pc[i]), __dmul_rn(ccc, dm[i]));
    if (i < coef_len) coef[i] = i < n ? __dmul_rn(acoeff, acoeff) : (i < n + mu ? __dmul_rn(cmu, w[i - n]) : 0.0);
}

// Y[r][:] = (X[r][:] - m_new) / sigma for the first mu candidates BY INDEX (cholesky_cmaes.cpp:86-89)
__global__ void __launch_bounds__(256) chol_steps_kernel(double *__restrict__ Y, const double *__restrict__ X, int64_t ld, int n,
        const double *__restrict__ m, double sigma) {
    const int r = blockIdx.x;
    const double *x = X + (int64_t) r * ld;
    double *y = Y + (int64_t) r * ld;
    for (int i = threadIdx.x; i < n; i += blockDim.x) y[i] = __ddiv_rn(__dsub_rn(x[i], m[i]), sigma);
}

// One panel of the blocked Cholesky of M (lower, row-major, pitch ld): every CTA factors the nb x nb diagonal block at k0
// redundantly in shared memory (no exchange between CTAs), CTA b then solves its 64 rows of the panel below,
// X L11^T = M21, and writes them twice: in place (row-major) and transposed into Lt (k-major, the SYRK operand of the
// trailing update).  CTA 0 writes L11 back.  info: first non-positive pivot (1-based), 0 = ok.
__global__ void __launch_bounds__(256) chol_panel_kernel(double *__restrict__ M, int64_t ld, int n, int k0, int nb,
        double *__restrict__ Lt, int *__restrict__ info) {
    __shared__ double S[CH_NB][CH_NB + 1];
    __shared__ int bad;
    const int tid = threadIdx.x;
    if (tid == 0) bad = 0;
    for (int q = tid; q < CH_NB * CH_NB; q += 256) {
        const int i = q / CH_NB, j = q % CH_NB;
        S[i][j] = (i < nb && j <= i) ? M[(int64_t) (k0 + i) * ld + k0 + j] : 0.0;
    }
    __syncthreads();
    const int tx = tid & 15, ty = tid >> 4;
    for (int j = 0; j < nb; j++) {
        if (tid == 0) {
            const double p = S[j][j];
            if (!(p > 0.0)) bad = bad ? bad : j + 1;
            const double dj = sqrt(p > 0.0 ? p : 1.0);
            S[j][j] = dj;
        }
        __syncthreads();
        if (tid > j && tid < nb) S[tid][j] = S[tid][j] / S[j][j];
        __syncthreads();
        // trailing lower triangle: (i, t), j < t <= i < nb
        for (int i = j + 1 + ty; i < nb; i += 16) {
            const double lij = S[i][j];
            for (int t = j + 1 + tx; t <= i; t += 16) S[i][t] -= lij * S[t][j];
        }
        __syncthreads();
    }
    if (bad && blockIdx.x == 0 && tid == 0) atomicCAS(info, 0, k0 + bad);
    if (blockIdx.x == 0) {
        for (int q = tid; q < CH_NB * CH_NB; q += 256) {
            const int i = q / CH_NB, j = q % CH_NB;
            if (i < nb && j <= i) M[(int64_t) (k0 + i) * ld + k0 + j] = S[i][j];
        }
    }
    // the panel below: thread r owns row r0 + r, its nb entries in registers; column-oriented substitution (after x_c is
    // known the remaining entries are updated by independent FMAs)
    const int r = k0 + nb + blockIdx.x * CH_NB + tid;
    if (tid < CH_NB && r < n) {
        double x[CH_NB];
        double *row = M + (int64_t) r * ld + k0;
#pragma unroll
        for (int c = 0; c < CH_NB; c++) x[c] = c < nb ? row[c] : 0.0;
#pragma unroll
        for (int c = 0; c < CH_NB; c++) {
            const double xc = c < nb ? x[c] / S[c][c] : 0.0;
            x[c] = xc;
#pragma unroll
            for (int t = c + 1; t < CH_NB; t++) x[t] -= xc * S[t][c];
        }
#pragma unroll
        for (int c = 0; c < CH_NB; c++) {
            if (c < nb) {
                row[c] = x[c];
                Lt[(int64_t) c * ld + r] = x[c];
            }
        }
    }
}

// At <- L^T (upper triangle of the k-major operand, zeros below) from the factored M, and the factorisation status into
// the report (its eig_info slot: 0 = ok, j = pivot j was not positive)
__global__ void __launch_bounds__(256) chol_store_kernel(const double *__restrict__ L, double *__restrict__ At, int n, int64_t ld,
        int *__restrict__ info, bbgpu_cma_report *rep) {
    __shared__ double tile[32][33];
    const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int r = ty; r < 32; r += 8) {
        const int i = by + r, j = bx + tx;
        tile[r][tx] = (i < n && j < n && j <= i) ? L[(int64_t) i * ld + j] : 0.0;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int i = bx + r, j = by + tx;
        if (i < n && j < n) At[(int64_t) i * ld + j] = tile[tx][r];
    }
    if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) {
        rep->eig_info = *info;
        *info = 0;
    }
}

// v = A^-1 dmean with the factor stored as At (row j = column j of A), then p_sigma <- (1 - cs) p_sigma + csc v and
// ||p_sigma|| into the report (cholesky_cmaes.cpp:93-104, base_cmaes.cpp:179).  One CTA: 32-column blocks, the diagonal
// triangle by warp 0, the rest of the right-hand side by all threads (coalesced along the rows of At).
__global__ void __launch_bounds__(1024) chol_ps_kernel(const double *__restrict__ At, int64_t ld, int n, const double *__restrict__ dm,
        double *__restrict__ ps, double cs, double csc, bbgpu_cma_report *rep) {
    extern __shared__ double rhs[]; // n, then 32 solved values, then 32 reduction slots
    double *vblk = rhs + n, *red = vblk + 32;
    const int tid = threadIdx.x, lane = tid & 31;
    for (int i = tid; i < n; i += blockDim.x) rhs[i] = dm[i];
    __syncthreads();
    for (int j0 = 0; j0 < n; j0 += 32) {
        const int nb = min(32, n - j0);
        if (tid < 32) {
            double rl = lane < nb ? rhs[j0 + lane] : 0.0;
            for (int c = 0; c < nb; c++) {
                const double dcc = At[(int64_t) (j0 + c) * ld + j0 + c];
                const double vc = __shfl_sync(0xffffffffu, rl, c) / dcc;
                if (lane == c) rl = vc;
                else if (lane > c && lane < nb) rl -= At[(int64_t) (j0 + c) * ld + j0 + lane] * vc;
            }
            if (lane < nb) { vblk[lane] = rl; rhs[j0 + lane] = rl; }
        }
        __syncthreads();
        for (int i = j0 + 32 + tid; i < n; i += blockDim.x) {
            double s = rhs[i];
            for (int c = 0; c < nb; c++) s -= At[(int64_t) (j0 + c) * ld + i] * vblk[c];
            rhs[i] = s;
        }
        __syncthreads();
    }
    double ss = 0.0;
    for (int i = tid; i < n; i += blockDim.x) {
        const double p = __dadd_rn(__dmul_rn(1.0 - cs, ps[i]), __dmul_rn(csc, rhs[i]));
        ps[i] = p;
        ss = fma(p, p, ss);
    }
    ss = block_sum(ss, red);
    if (tid == 0) {
        rep->pslen = sqrt(ss);
        rep->hsig = 1;
    }
}

// converged(): ||x_k||_2 of every candidate (cholesky_cmaes.cpp:146-155 uses dnrm2 row by row), one warp per row
__global__ void __launch_bounds__(256) row_norms_kernel(const double *__restrict__ X, int64_t ld, int n, int rows,
        double *__restrict__ out) {
    const int k = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (k >= rows) return;
    const double *x = X + (int64_t) k * ld;
    double s = 0.0;
    for (int i = lane; i < n; i += 32) s = fma(x[i], x[i], s);
    s = warp_sum(s);
    if (lane == 0) out[k] = sqrt(s);
}

// mean and M2 = sum (r_k - mean)^2 of the row norms (the quantity Welford's recurrence accumulates, :148-154)
__global__ void __launch_bounds__(1024) radius_stats_kernel(const double *__restrict__ r, int rows, double *__restrict__ out2) {
    __shared__ double red[32];
    double s = 0.0;
    for (int k = threadIdx.x; k < rows; k += blockDim.x) s += r[k];
    s = block_sum(s, red);
    const double mean = s / rows;
    __syncthreads();
    double m2 = 0.0;
    for (int k = threadIdx.x; k < rows; k += blockDim.x) {
        const double dlt = r[k] - mean;
        m2 = fma(dlt, dlt, m2);
    }
    m2 = block_sum(m2, red);
    if (threadIdx.x == 0) { out2[0] = mean; out2[1] = m2; }
}

} // namespace bb
