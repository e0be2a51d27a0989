// xNES on the device (reference: nes/xnes.cpp, Glasmachers et al. 2010: exponential natural evolution strategies).
//
// State: the mean mu, the step size sigma (host scalar) and the transform B with det 1 (x = mu + sigma B z); np = 4 +
// floor(3 ln n) points per generation with fixed rank utilities u_r (sum 0).  Per generation (xnes.cpp:110-191):
//   sample (the sampling GEMM with B^T as operand), evaluate, rank;
//   G_delta = sum_r u_r z_(r);  G = sum_r u_r (z_(r) z_(r)^T - I);  G_sigma = tr(G) / n;  G_B = G - G_sigma I;
//   mu += eta_mu sigma B G_delta;  sigma *= exp(eta_sigma G_sigma / 2);  B <- B expm(eta_B G_B / 2).
// The matrix exponential is the reference's: eigendecomposition, expm(C) = V exp(L) V^T (:227-253) -- basis independent, so
// any valid decomposition gives the same matrix to rounding (own solver / cuSOLVER through eig_core).  Everything O(n^3)
// is two tensor-core GEMMs: E = V^T exp(L) V and B^T <- E B^T.
#pragma once

#include "common.cuh"

namespace bb {

// G_delta[d] = sum_r u_r Z[idx_r][d], r ascending like the reference (:131-136)
__global__ void __launch_bounds__(256) xnes_gdelta_kernel(double *__restrict__ g, const double *__restrict__ Z, int64_t ld, int n, int np,
        const int *__restrict__ rank_idx, const double *__restrict__ u) {
    const int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= n) return;
    double s = 0.0;
    for (int r = 0; r < np; r++) s = __dadd_rn(s, __dmul_rn(u[r], Z[(int64_t) rank_idx[r] * ld + d]));
    g[d] = s;
}

// G[d1][d2] = sum_r u_r (z_r[d1] z_r[d2] - delta), r ascending (:139-149); 32 x 32 tiles, the np rows of the two column
// blocks staged in shared memory
constexpr int XN_MAXNP = 64;
__global__ void __launch_bounds__(256) xnes_gm_kernel(double *__restrict__ G, const double *__restrict__ Z, int64_t ld, int n, int np,
        const int *__restrict__ rank_idx, const double *__restrict__ u) {
    __shared__ double za[XN_MAXNP][33], zb[XN_MAXNP][33], us[XN_MAXNP];
    const int b1 = blockIdx.y * 32, b2 = blockIdx.x * 32;
    for (int q = threadIdx.x; q < np * 32; q += 256) {
        const int r = q / 32, c = q % 32;
        const double *row = Z + (int64_t) rank_idx[r] * ld;
        za[r][c] = b1 + c < n ? row[b1 + c] : 0.0;
        zb[r][c] = b2 + c < n ? row[b2 + c] : 0.0;
    }
    if (threadIdx.x < np) us[threadIdx.x] = u[threadIdx.x];
    __syncthreads();
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int i = ty; i < 32; i += 8) {
        const int d1 = b1 + i, d2 = b2 + tx;
        if (d1 >= n || d2 >= n) continue;
        const double eye = d1 == d2 ? 1.0 : 0.0;
        double s = 0.0;
        for (int r = 0; r < np; r++) s = __dadd_rn(s, __dmul_rn(us[r], __dsub_rn(__dmul_rn(za[r][i], zb[r][tx]), eye)));
        G[(int64_t) d1 * ld + d2] = s;
    }
}

// G_sigma = sum_d G[d][d] / n (:152-155) into out[0]; one CTA
__global__ void __launch_bounds__(1024) xnes_trace_kernel(const double *__restrict__ G, int64_t ld, int n, double *__restrict__ out) {
    __shared__ double red[32];
    double s = 0.0;
    for (int d = threadIdx.x; d < n; d += blockDim.x) s += G[(int64_t) d * ld + d] / n;
    s = block_sum(s, red);
    if (threadIdx.x == 0) out[0] = s;
}

// C = (eta_B / 2) (G - G_sigma I) in place (:158-176)
__global__ void __launch_bounds__(256) xnes_scale_kernel(double *__restrict__ G, int64_t ld, int n, const double *__restrict__ gsigma,
        double half_etab) {
    const int d2 = blockIdx.x * blockDim.x + threadIdx.x, d1 = blockIdx.y;
    if (d2 >= n) return;
    double g = G[(int64_t) d1 * ld + d2];
    if (d1 == d2) g = __dsub_rn(g, *gsigma);
    G[(int64_t) d1 * ld + d2] = __dmul_rn(half_etab, g);
}

// mu[d] += sum_j ((eta_mu sigma) B[d][j]) G_delta[j], j ascending (:164-168); Bt[j][d] = B[d][j]: coalesced over d
__global__ void __launch_bounds__(256) xnes_mean_kernel(double *__restrict__ mu, const double *__restrict__ Bt, int64_t ld, int n,
        const double *__restrict__ g, double etamu_sigma) {
    const int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= n) return;
    double m = mu[d];
    for (int j = 0; j < n; j++) m = __dadd_rn(m, __dmul_rn(__dmul_rn(etamu_sigma, Bt[(int64_t) j * ld + d]), g[j]));
    mu[d] = m;
}

// S[j][:] = exp(W[j]) Vt[j][:]
__global__ void __launch_bounds__(256) xnes_expscale_kernel(const double *__restrict__ Vt, const double *__restrict__ W,
        double *__restrict__ S, int n, int64_t ld) {
    const int j = blockIdx.x;
    const double e = exp(W[j]);
    for (int i = threadIdx.x; i < n; i += blockDim.x) S[(int64_t) j * ld + i] = e * Vt[(int64_t) j * ld + i];
}

// status of the eigensolver behind the matrix exponential into the report
__global__ void xnes_info_kernel(const int *__restrict__ info, bbgpu_cma_report *rep) {
    if (threadIdx.x == 0 && blockIdx.x == 0) rep->eig_info = *info;
}

} // namespace bb
