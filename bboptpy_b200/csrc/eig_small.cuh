// One-warp symmetric eigensolver for n <= 64 (EISPACK tred2 + tql2, the reference's own algorithm), shared by the batched
// small-n CMA-ES engine (cmab.cu) and the leaves of the tridiagonal divide & conquer (stedc.cu).
#pragma once

#include "common.cuh"

namespace bb {

// 1/sqrt(t) for normal positive t: hardware seed (2^-23) + one third-order correction, ~1 ulp; no special-case path
static __device__ __forceinline__ double rsqrt_nr(double t) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(t));
    const double e = fma(-t * y, y, 1.0);
    return fma(y * e, fma(e, 0.375, 0.5), y);
}

// ------------------------------------------------------------------------------------------------------------------
// In-CTA symmetric eigendecomposition: EISPACK tred2 + tql2 as the reference uses them (cmaes.cpp:285-478), on
// Vt[c * ldv + r] = V[r][c] in shared memory (ldv odd: rows and columns are both conflict-free), d / e length n.
// On exit d holds the eigenvalues (unsorted), column c of V the eigenvector of d[c].  Returns false if the QL
// iteration did not converge (NaN input).
//
// ONE WARP runs the whole decomposition (the caller's other warps wait at a CTA barrier): lane L owns rows / columns L
// and L + 32 (n <= 64).  A batch fills the chip with runs, so what counts per run is the number of issued
// instructions, not the latency of one decomposition: spreading these O(n^2)-per-step loops over four warps cost a
// CTA barrier per phase, redundant reductions in every warp and -- in the QL iteration, a scalar recurrence -- three
// idle warps polling; measured 290 k warp instructions per decomposition at n = 40 against ~90 k here
// (profiles/r01_ncu_cmab.txt).  Only __syncwarp() orders the shared-memory traffic.
// tridiagonal = true: the input IS tridiagonal -- the caller presets V = I, d = diagonal, e[i] = T[i+1][i] (e[n-1] = 0) and only
// the QL iteration runs (the leaves of the divide & conquer in stedc.cu).
// ------------------------------------------------------------------------------------------------------------------
static __device__ __noinline__ bool warp_tred2_tql2(const int n, const int ldv, double *Vt, double *d, double *e, int *ql_iters,
        const bool tridiagonal) {
    const int lane = threadIdx.x & 31;
    const int r0 = lane, r1 = lane + 32;
    if (!tridiagonal) {
    if (r0 < n) d[r0] = Vt[r0 * ldv + n - 1];
    if (r1 < n) d[r1] = Vt[r1 * ldv + n - 1];
    __syncwarp();
    // ---- Householder reduction to tridiagonal form; the active block V[0:i,0:i] is kept fully symmetric
    for (int i = n - 1; i > 0; i--) {
        const bool a0 = r0 < i, a1 = r1 < i;
        const double v0 = a0 ? d[r0] : 0.0, v1 = a1 ? d[r1] : 0.0;
        const double sc = bb::warp_sum(fabs(v0) + fabs(v1));
        if (sc == 0.0) {
            const double dl = d[i - 1];
            __syncwarp();
            if (lane == 0) e[i] = dl;
            if (a0) { d[r0] = Vt[r0 * ldv + i - 1]; Vt[r0 * ldv + i] = 0.0; Vt[i * ldv + r0] = 0.0; }
            if (a1) { d[r1] = Vt[r1 * ldv + i - 1]; Vt[r1 * ldv + i] = 0.0; Vt[i * ldv + r1] = 0.0; }
            if (lane == 0) d[i] = 0.0;
            __syncwarp();
            continue;
        }
        const double t0 = v0 / sc, t1 = v1 / sc;
        const double hs = bb::warp_sum(t0 * t0 + t1 * t1);
        const double f = d[i - 1] / sc;
        double g = sqrt(hs);
        if (f > 0) g = -g;
        const double h = hs - f * g;
        __syncwarp();
        const double dr0 = (r0 == i - 1) ? f - g : t0, dr1 = (r1 == i - 1) ? f - g : t1; // the Householder vector
        if (a0) d[r0] = dr0;
        if (a1) d[r1] = dr1;
        if (lane == 0) e[i] = sc * g;
        __syncwarp();
        // p = A d / h (rows r0, r1), two interleaved partial sums per row; V[j][i] = d[j]
        double p0 = 0.0, p1 = 0.0, q0 = 0.0, q1 = 0.0;
        {
            int k = 0;
            for (; k + 1 < i; k += 2) {
                const double dk = d[k], dk1 = d[k + 1];
                if (a0) { p0 = fma(Vt[k * ldv + r0], dk, p0); q0 = fma(Vt[(k + 1) * ldv + r0], dk1, q0); }
                if (a1) { p1 = fma(Vt[k * ldv + r1], dk, p1); q1 = fma(Vt[(k + 1) * ldv + r1], dk1, q1); }
            }
            if (k < i) {
                const double dk = d[k];
                if (a0) p0 = fma(Vt[k * ldv + r0], dk, p0);
                if (a1) p1 = fma(Vt[k * ldv + r1], dk, p1);
            }
        }
        p0 = (p0 + q0) / h;
        p1 = (p1 + q1) / h;
        if (a0) Vt[i * ldv + r0] = dr0;
        if (a1) Vt[i * ldv + r1] = dr1;
        const double fs = bb::warp_sum((a0 ? p0 * dr0 : 0.0) + (a1 ? p1 * dr1 : 0.0));
        const double hh = fs / (h + h);
        const double er0 = p0 - hh * dr0, er1 = p1 - hh * dr1;
        if (a0) e[r0] = er0;
        if (a1) e[r1] = er1;
        __syncwarp();
        // rank-2 update of rows r0, r1 over all columns of the block
        for (int k = 0; k < i; k++) {
            const double ek = e[k], dk = d[k];
            if (a0) Vt[k * ldv + r0] -= dr0 * ek + er0 * dk;
            if (a1) Vt[k * ldv + r1] -= dr1 * ek + er1 * dk;
        }
        __syncwarp();
        if (a0) { d[r0] = Vt[r0 * ldv + i - 1]; Vt[r0 * ldv + i] = 0.0; }
        if (a1) { d[r1] = Vt[r1 * ldv + i - 1]; Vt[r1 * ldv + i] = 0.0; }
        if (lane == 0) d[i] = h;
        __syncwarp();
    }
    // ---- accumulate the transformations (lane L owns columns L, L + 32)
    for (int i = 0; i < n - 1; i++) {
        if (lane == 0) {
            Vt[i * ldv + n - 1] = Vt[i * ldv + i];
            Vt[i * ldv + i] = 1.0;
        }
        const double h = d[i + 1];
        __syncwarp();
        const bool a0 = r0 <= i, a1 = r1 <= i;
        double *u = Vt + (i + 1) * ldv;
        if (h != 0.0) {
            double g0 = 0.0, g1 = 0.0, w0 = 0.0, w1 = 0.0;
            const double *c0 = Vt + r0 * ldv, *c1 = Vt + r1 * ldv;
            int k = 0;
            for (; k + 1 <= i; k += 2) {
                const double uk = u[k], uk1 = u[k + 1];
                if (a0) { g0 = fma(uk, c0[k], g0); w0 = fma(uk1, c0[k + 1], w0); }
                if (a1) { g1 = fma(uk, c1[k], g1); w1 = fma(uk1, c1[k + 1], w1); }
            }
            if (k <= i) {
                const double uk = u[k];
                if (a0) g0 = fma(uk, c0[k], g0);
                if (a1) g1 = fma(uk, c1[k], g1);
            }
            g0 += w0;
            g1 += w1;
            if (a0) d[r0] = u[r0] / h;
            if (a1) d[r1] = u[r1] / h;
            __syncwarp();
            for (k = 0; k <= i; k++) {
                const double dk = d[k];
                if (a0) Vt[r0 * ldv + k] -= g0 * dk;
                if (a1) Vt[r1 * ldv + k] -= g1 * dk;
            }
            __syncwarp();
        }
        if (a0) u[r0] = 0.0;
        if (a1) u[r1] = 0.0;
        __syncwarp();
    }
    {
        const double x0 = r0 < n ? Vt[r0 * ldv + n - 1] : 0.0, x1 = r1 < n ? Vt[r1 * ldv + n - 1] : 0.0;
        const double s0 = r0 + 1 < n ? e[r0 + 1] : 0.0, s1 = r1 + 1 < n ? e[r1 + 1] : 0.0;
        __syncwarp();
        if (r0 < n) { d[r0] = x0; Vt[r0 * ldv + n - 1] = (r0 == n - 1) ? 1.0 : 0.0; e[r0] = s0; }
        if (r1 < n) { d[r1] = x1; Vt[r1 * ldv + n - 1] = (r1 == n - 1) ? 1.0 : 0.0; e[r1] = s1; }
        __syncwarp();
    }
    }
    // ---- implicit QL with accumulation (cmaes.cpp:383-463).  Every lane runs the scalar recurrence (identical values:
    // one warp instruction either way) and applies each rotation to its own two rows of V straight from registers.
    const double eps = 2.220446049250313e-16;
    const int rc0 = min(r0, n - 1), rc1 = r1 < n ? r1 : rc0;
    double fsh = 0.0, tst1 = 0.0;
    int iters = 0;
    for (int l = 0; l < n; l++) {
        tst1 = fmax(tst1, fabs(d[l]) + fabs(e[l]));
        int m = l;
        while (m < n && !(fabs(e[m]) <= eps * tst1)) m++;
        if (m >= n) return false; // cmaes.cpp:399-401 leaves the loop here (only reachable with non-finite input)
        if (m > l) {
            while (true) {
                double g = d[l];
                const double el = e[l];
                const double p0 = (d[l + 1] - g) / (2.0 * el);
                double rr = sqrt(fma(p0, p0, 1.0));
                if (p0 < 0) rr = -rr;
                const double dl = el / (p0 + rr), dl1 = el * (p0 + rr);
                const double hshift = g - dl, el1 = e[l + 1];
                __syncwarp();
                if (lane == 0) { d[l] = dl; d[l + 1] = dl1; }
                for (int i = l + 2 + lane; i < n; i += 32) d[i] -= hshift;
                __syncwarp();
                fsh += hshift;
                // one Givens rotation per step: r = hypot(p, e_i), s = e_i / r, c = p / r (cmaes.cpp:432-437) through
                // ONE reciprocal square root: the recurrence is a dependency chain, two divisions and a square root per
                // step were most of its length.  The loop is branch free: pointers walk the columns, every lane stores the
                // (identical) new d / e, and lanes without a row of their own redo row n-1 (same loads, same values).
                double p = d[m], c = 1.0, c2 = 1.0, c3 = 1.0, s = 0.0, s2 = 0.0;
                const double *pe = e + (m - 1), *pd = d + (m - 1);
                double *we = e + m, *wd = d + m;
                double ei = *pe, di = *pd;
                double *q0 = Vt + m * ldv + rc0, *q1 = Vt + m * ldv + rc1;
                double h0 = *q0, h1 = *q1;
                for (int cnt = m - l; cnt > 0; cnt--) {
                    const int back = cnt > 1 ? -1 : 0; // next step's operands (none after the last step)
                    const double en = pe[back], dn = pd[back];
                    const double vi0 = q0[-ldv], vi1 = q1[-ldv];
                    c3 = c2;
                    c2 = c;
                    s2 = s;
                    g = c * ei;
                    const double hq = c * p;
                    const double t = fma(p, p, ei * ei);
                    const double rinv = rsqrt_nr(t);
                    const double enew = s * (t * rinv);
                    s = ei * rinv;
                    c = p * rinv;
                    p = c * di - s * g;
                    *we = enew;
                    *wd = hq + s * (c * g + s * di);
                    *q0 = fma(s, vi0, c * h0);
                    *q1 = fma(s, vi1, c * h1);
                    h0 = fma(c, vi0, -s * h0);
                    h1 = fma(c, vi1, -s * h1);
                    ei = en;
                    di = dn;
                    pe--; pd--; we--; wd--;
                    q0 -= ldv;
                    q1 -= ldv;
                }
                *q0 = h0;
                *q1 = h1;
                __syncwarp();
                p = -s * s2 * c3 * el1 * e[l] / dl1;
                __syncwarp();
                if (lane == 0) { e[l] = s * p; d[l] = c * p; }
                __syncwarp();
                iters++;
                if (iters > 60 * n || !(p == p)) return false;
                if (!(fabs(s * p) > eps * tst1)) break;
            }
        }
        __syncwarp();
        if (lane == 0) { d[l] += fsh; e[l] = 0.0; }
        __syncwarp();
    }
    if (lane == 0) *ql_iters = iters;
    return true;
}


} // namespace bb
