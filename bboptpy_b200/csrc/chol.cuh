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

// 1 / p for a positive, normal p from the 20-bit hardware seed and two Newton steps (~1 ulp): the IEEE division is ~30
// dependent instructions, and it sits on the column-to-column chain of the factorisation
__device__ __forceinline__ double rcp_pos(double p) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(p));
    r = fma(r, fma(-p, r, 1.0), r);
    r = fma(r, fma(-p, r, 1.0), r);
    return r;
}

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

// One panel of the blocked Cholesky of M (lower, row-major, pitch ld), 256 threads per CTA.  Every CTA factors the nb x nb
// diagonal block at k0 redundantly (no exchange between CTAs), CTA b then solves its 64 rows of the panel below,
// X L11^T = M21, and writes them twice: in place (row-major) and transposed into Lt (k-major, the SYRK operand of the
// trailing update).  CTA 0 writes L11 back.  info: first non-positive pivot (1-based), 0 = ok.
// Both stages keep the rows in registers -- four threads per row, thread (i, q) owns columns 4u + q -- and are
// right-looking, so the work between two dependent steps is a run of independent FMAs (the first version updated the
// trailing block in shared memory: 78 us per panel of dependent shared-memory round trips; one thread per row with the
// 64 x 64 loops unrolled: 85 us of straight-line code).  Factorisation, column j: the owners publish their still
// unscaled entries a_i[j] (row j: the pivot p), ONE barrier, then a_i[t] -= a_i[j] a_t[j] / p for the rest of the row;
// the scaling L[i][j] = a_i[j] / sqrt(p) happens once at the end, off the chain.
__global__ void __launch_bounds__(256) chol_panel_kernel(double *__restrict__ M, int64_t ld, int n, int k0, int nb,
        double *__restrict__ Lt, int *__restrict__ info) {
    __shared__ double S[CH_NB][CH_NB + 1]; // the factored diagonal block (lower)
    __shared__ double col[2][CH_NB];
    __shared__ double piv[CH_NB], invd[CH_NB];
    __shared__ int bad;
    const int tid = threadIdx.x, i = tid >> 2, q = tid & 3;
    constexpr int NU = CH_NB / 4;
    double a[NU];
    if (tid == 0) bad = 0;
    {
        const double *row = M + (int64_t) (k0 + min(i, nb - 1)) * ld + k0;
#pragma unroll
        for (int u = 0; u < NU; u++) {
            const int t = 4 * u + q;
            a[u] = (i < nb && t <= i) ? row[t] : (t == i ? 1.0 : 0.0); // identity beyond nb
        }
    }
    __syncthreads();
#pragma unroll
    for (int u0 = 0; u0 < NU; u0++) {
#pragma unroll
        for (int q0 = 0; q0 < 4; q0++) {
            const int j = 4 * u0 + q0;
            double *cj = col[j & 1];
            if (q == q0) cj[i] = a[u0];
            __syncthreads();
            const double p = cj[j];
            if (tid == 0) {
                if (!(p > 0.0) && bad == 0) bad = j + 1;
                piv[j] = p > 0.0 ? p : 1.0;
            }
            const double aij = cj[i] * rcp_pos(p > 0.0 ? p : 1.0);
#pragma unroll
            for (int u = u0; u < NU; u++) {
                const int t = 4 * u + q;
                if (t > j) a[u] -= aij * cj[t];
            }
        }
    }
    __syncthreads();
    if (tid < CH_NB) {
        const double sq = sqrt(piv[tid]);
        piv[tid] = sq;          // now the diagonal of the factor
        invd[tid] = 1.0 / sq;
    }
    __syncthreads();
    if (bad && blockIdx.x == 0 && tid == 0) atomicCAS(info, 0, k0 + bad);
#pragma unroll
    for (int u = 0; u < NU; u++) {
        const int t = 4 * u + q;
        const double l = t < i ? a[u] * invd[t] : (t == i ? piv[t] : 0.0);
        S[i][t] = l;
        if (blockIdx.x == 0 && i < nb && t <= i) M[(int64_t) (k0 + i) * ld + k0 + t] = l;
    }
    __syncthreads();
    // the panel below: the four threads of row r hold x; x_c comes from its owner by a shuffle inside the quad, then every
    // thread updates its own entries with independent FMAs
    const int r = k0 + nb + blockIdx.x * CH_NB + i;
    const bool live = r < n;
    double *row = M + (int64_t) (live ? r : 0) * ld + k0;
#pragma unroll
    for (int u = 0; u < NU; u++) {
        const int c = 4 * u + q;
        a[u] = (live && c < nb) ? row[c] : 0.0;
    }
    const int quad = (tid & 31) & ~3;
#pragma unroll
    for (int u0 = 0; u0 < NU; u0++) {
#pragma unroll
        for (int q0 = 0; q0 < 4; q0++) {
            const int c = 4 * u0 + q0;
            const double mine = a[u0] * invd[c];
            const double xc = __shfl_sync(0xffffffffu, mine, quad | q0);
            if (q == q0) a[u0] = xc;
#pragma unroll
            for (int u = u0; u < NU; u++) {
                const int t = 4 * u + q;
                if (t > c) a[u] -= xc * S[t][c];
            }
        }
    }
    if (live) {
#pragma unroll
        for (int u = 0; u < NU; u++) {
            const int c = 4 * u + q;
            if (c < nb) {
                row[c] = a[u];
                Lt[(int64_t) c * ld + r] = a[u];
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
__global__ void __launch_bounds__(512) chol_ps_kernel(const double *__restrict__ At, int64_t ld, int n, const double *__restrict__ dm,
        double *__restrict__ ps, double cs, double csc, bbgpu_cma_report *rep) {
    extern __shared__ double rhs[]; // n, then 32 solved values, then 32 reduction slots
    double *vblk = rhs + n, *red = vblk + 32;
    const int tid = threadIdx.x, lane = tid & 31;
    for (int i = tid; i < n; i += blockDim.x) rhs[i] = dm[i];
    __syncthreads();
    for (int j0 = 0; j0 < n; j0 += 32) {
        const int nb = min(32, n - j0);
        if (tid < 32) {
            // lane l keeps A[j0+l][j0+c], c = 0..31 (row c of the At block, element l): 32 independent loads up front, so the
            // substitution chain below runs on registers (with the loads inside it every step paid an L2 round trip: 1 ms)
            double acol[32];
#pragma unroll
            for (int c = 0; c < 32; c++) acol[c] = (c < nb && lane < nb) ? At[(int64_t) (j0 + c) * ld + j0 + lane] : (c == lane ? 1.0 : 0.0);
            double rl = lane < nb ? rhs[j0 + lane] : 0.0;
            double dgl = 1.0; // this lane's diagonal entry A[j0+lane][j0+lane]
#pragma unroll
            for (int c = 0; c < 32; c++) dgl = c == lane ? acol[c] : dgl;
            const double idg = 1.0 / dgl; // one division per lane, off the chain: row l of D^-1 A v = D^-1 rhs
            rl *= idg;
#pragma unroll
            for (int c = 0; c < 32; c++) acol[c] *= idg;
#pragma unroll
            for (int c = 0; c < 32; c++) {
                const double vc = __shfl_sync(0xffffffffu, rl, c);
                if (lane > c) rl -= acol[c] * vc;
            }
            if (lane < nb) { vblk[lane] = rl; rhs[j0 + lane] = rl; }
        }
        __syncthreads();
        for (int i = j0 + 32 + tid; i < n; i += blockDim.x) { // nb == 32 here (a shorter last block has nothing below it)
            double ai[32];
#pragma unroll
            for (int c = 0; c < 32; c++) ai[c] = At[(int64_t) (j0 + c) * ld + i]; // 32 loads in flight, one L2 round trip
            double s = rhs[i];
#pragma unroll
            for (int c = 0; c < 32; c++) s -= ai[c] * vblk[c];
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
