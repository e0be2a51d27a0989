// Eigen-decomposition of a symmetric tridiagonal matrix by divide & conquer (Cuppen; deflation as in LAPACK dlaed2,
// eigenvectors through the Gu-Eisenstat / Loewner recomputation of z as in dlaed3), entirely on the device.
// Second half of the replacement for cuSOLVER's syevd at n = 1000 (first half: sytrd.cu); cuSOLVER exports no device-side
// tridiagonal solver.  Reference being replaced: the QL iteration of cma/cmaes.cpp:383-478 (any correct eigenbasis serves:
// DESIGN.md "The eigendecomposition at n = 1000").
//
// T is torn into leaves of <= 32 rows (boundaries at multiples of 8), solved by the one-warp QL solver of eig_small.cuh;
// every merge of two solved halves  diag(D1, D2) + rho z z^T  runs as a handful of launches over ALL merges of its level:
//   setup    z from the touching rows of the children's eigenvectors, merge-sort of D by counting ranks, the sequential
//            deflation scan (tiny z_j; close D with a Givens rotation) -> per position an action code and slot
//   secular  one warp per root: the root of 1 + rho sum z_j^2 / (d_j - lambda) in its interval by bisection on
//            mu = lambda - d_origin (origin = the nearer pole, so d_j - lambda keeps full relative accuracy)
//   zhat     z recomputed from the roots (Loewner), one warp per component: the computed vectors are orthogonal to
//            working precision whatever the clustering
//   vectors  eigenvectors of the rank-one problem, one warp per root, normalised
//   gather   one thread per row of the children's basis: permutation + the chain of Givens rotations, emitting deflated
//            vectors to their final place and the others as the right operand of
//   gemm     new vectors = (rank-one eigenvectors) x (gathered basis)
// Eigenvectors are stored TRANSPOSED (row = eigenvector): that is the layout the CMA engine keeps (Bt) and what
// cusolverDnDormtr sees as the column-major matrix to multiply by the Householder reflectors.
#include <algorithm>
#include <vector>

#include "eig_small.cuh"
#include "gemm_tma.cuh" // the engine's TMA-fed DMMA GEMM for the merges

namespace {

#ifndef DC_LEAF_ROWS
#define DC_LEAF_ROWS 32 // measured at n = 1000: 32-row leaves 1.93 ms, 64-row leaves 2.12 ms (the QL chain of a leaf is the longer pole)
#endif
constexpr int DC_LEAF = DC_LEAF_ROWS;
constexpr int DC_THREADS = 256;

struct DcNode { int o, k, m; }; // rows [o, o+k), children [o, o+m) and [o+m, o+k)

struct DcWork { // per-position scratch (length n each) and per-node scalars
    double *Ds, *zs, *dk, *zk, *mu, *zhat, *rc, *rs, *ddefl, *rho;
    int *perm, *act, *slot, *org, *K, *lastslot;
};

// ---- leaves ----------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32) dc_leaf_kernel(const DcNode *nodes, const double *dadj, const double *e, double *Zt, int ld,
        double *lam, int *info) {
    extern __shared__ __align__(16) double sm[];
    const DcNode nd = nodes[blockIdx.x];
    const int k = nd.k, o = nd.o, ldv = k | 1, lane = threadIdx.x;
    double *Vt = sm, *d = Vt + k * ldv, *ee = d + k, *srt = ee + k;
    for (int idx = lane; idx < k * ldv; idx += 32) Vt[idx] = 0.0;
    __syncwarp();
    for (int j = lane; j < k; j += 32) {
        Vt[j * ldv + j] = 1.0;
        d[j] = dadj[o + j];
        ee[j] = j < k - 1 ? e[o + j] : 0.0;
    }
    __syncwarp();
    int qi = 0;
    const bool ok = k == 1 ? true : bb::warp_tred2_tql2(k, ldv, Vt, d, ee, &qi, true);
    __syncwarp();
    if (!ok) { if (lane == 0) *info = 1; return; }
    for (int j = lane; j < k; j += 32) {
        const double val = d[j];
        int pos = 0;
        for (int q = 0; q < k; q++) pos += (d[q] < val) || (d[q] == val && q < j);
        srt[j] = (double) pos;
    }
    __syncwarp();
    for (int j = 0; j < k; j++) {
        const int pos = (int) srt[j];
        if (lane == 0) lam[o + pos] = d[j];
        for (int r = lane; r < k; r += 32) Zt[(size_t) (o + pos) * ld + o + r] = Vt[j * ldv + r];
    }
}

__global__ void dc_adjust_kernel(const double *d, const double *e, const int *bounds, int nb, double *dadj, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double v = d[i];
    for (int b = 0; b < nb; b++) { // Cuppen's tearing: both neighbours of a boundary lose |e|
        const int p = bounds[b];
        if (i == p - 1 || i == p) v -= fabs(e[p - 1]);
    }
    dadj[i] = v;
}

// ---- merge: setup + deflation scan (one CTA per merge) ---------------------------------------------------------------
__global__ void __launch_bounds__(DC_THREADS) dc_setup_kernel(const DcNode *nodes, const double *e, const double *Zt, int ld,
        const double *lam, DcWork w, double *lam_new) {
    const DcNode nd = nodes[blockIdx.x];
    const int k = nd.k, o = nd.o, m = nd.m, tid = threadIdx.x;
    __shared__ double red[32];
    __shared__ double s_tol;
    const double beta = e[o + m - 1];
    const double sgn = beta < 0 ? -1.0 : 1.0;
    // z = (last row of Q1, sign * first row of Q2); the rows of Zt are eigenvectors, so these are columns o+m-1 and o+m
    double ss = 0.0, dmax = 0.0;
    for (int j = tid; j < k; j += DC_THREADS) {
        const double z = j < m ? Zt[(size_t) (o + j) * ld + o + m - 1] : sgn * Zt[(size_t) (o + j) * ld + o + m];
        w.zk[o + j] = z; // scratch: unsorted
        ss += z * z;
        dmax = fmax(dmax, fabs(lam[o + j]));
    }
    ss = bb::block_sum(ss, red);
    dmax = bb::block_max(dmax, red);
    const double zn = sqrt(ss), rho = fabs(beta) * ss;
    __syncthreads();
    double zmax = 0.0;
    extern __shared__ __align__(16) double ssm[];
    for (int j = tid; j < k; j += DC_THREADS) ssm[j] = lam[o + j];
    __syncthreads();
    for (int j = tid; j < k; j += DC_THREADS) { // rank by counting (the halves hold roots, then deflated values: not sorted)
        const double dj = ssm[j];
        int pos = 0;
        for (int q = 0; q < k; q++) {
            const double dq = ssm[q]; // one address per warp: a broadcast
            pos += (dq < dj) || (dq == dj && q < j);
        }
        const double z = w.zk[o + j] / zn;
        w.perm[o + pos] = j;
        w.Ds[o + pos] = dj;
        w.zs[o + pos] = z;
        zmax = fmax(zmax, fabs(z));
    }
    zmax = bb::block_max(zmax, red);
    if (tid == 0) {
        s_tol = 8.0 * 2.220446049250313e-16 * fmax(dmax, zmax);
        w.rho[blockIdx.x] = rho;
    }
    __syncthreads();
    // The deflation scan is sequential in principle (dlaed2: every decision changes the next comparison); it runs over
    // SHARED memory (a walk over global memory cost ~2 us per pole and was most of the solver's time).
    double *D = ssm, *z = D + k, *src = z + k, *srs = src + k, *sdk = srs + k, *szk = sdk + k, *sdd = szk + k;
    int *sact = reinterpret_cast<int*>(sdd + k), *sslot = sact + k, *spec = sslot + k;
    for (int j = tid; j < k; j += DC_THREADS) { D[j] = w.Ds[o + j]; z[j] = w.zs[o + j]; src[j] = 0.0; srs[j] = 0.0; }
    __syncthreads();
    // Speculation, in parallel: the outcome of every pole's tests on the UNMODIFIED data (0 tiny, 1 stays, 2 would rotate
    // with its predecessor).  The sequential walk below only recomputes a test when the previous step really rotated.
    for (int j = tid; j < k; j += DC_THREADS) {
        int sp = 0;
        if (!(rho * fabs(z[j]) <= s_tol)) {
            int p = j - 1;
            while (p >= 0 && rho * fabs(z[p]) <= s_tol) p--;
            sp = 1;
            if (p >= 0) {
                const double zp = z[p], zj = z[j], t = D[j] - D[p];
                if (fabs(t * zj * zp) <= s_tol * (zj * zj + zp * zp)) sp = 2;
            }
        }
        spec[j] = sp;
    }
    __syncthreads();
    if (tid < 32) { // warp 0 walks the poles 32 at a time; only a chunk that holds a rotation is walked by one lane
        const double tol = s_tol;
        const int lane = tid;
        const unsigned below = (1u << lane) - 1u;
        int K = 0, nd_ = 0, pj = -1; // warp-uniform
        int dirty = 0; // the predecessor pole was changed by a rotation: its speculated test is void
        const bool alltiny = rho * zmax <= tol;
        for (int base = 0; base < k; base += 32) {
            const int j = base + lane;
            const bool valid = j < k;
            const int sp = valid ? (alltiny ? 0 : spec[j]) : -1;
            const unsigned rot = __ballot_sync(0xffffffffu, sp == 2);
            if (rot == 0u && !dirty) {
                // no decision in this chunk depends on another one: every pole finds its slot from two prefix counts
                const unsigned nt = __ballot_sync(0xffffffffu, sp > 0), tn = __ballot_sync(0xffffffffu, sp == 0);
                if (sp == 0) { // the component does not feel the rank-one term
                    const int r = nd_ + __popc(tn & below);
                    sact[j] = 0; sslot[j] = r; sdd[r] = D[j];
                } else if (sp > 0) {
                    const unsigned prev = nt & below;
                    const int pred = prev ? base + 31 - __clz(prev) : pj;
                    if (pred < 0) { sact[j] = 1; sslot[j] = 0; }
                    else { // the older pole stays in the secular problem
                        const int slot = K + __popc(prev) - (pj < 0 ? 1 : 0);
                        sact[j] = 3; sslot[j] = slot;
                        sdk[slot] = D[pred]; szk[slot] = z[pred];
                    }
                }
                if (nt) { K += __popc(nt) - (pj < 0 ? 1 : 0); pj = base + 31 - __clz(nt); }
                nd_ += __popc(tn);
            } else {
                if (lane == 0) {
                    const int jend = min(base + 32, k);
                    for (int q = base; q < jend; q++) {
                        const int sq = spec[q];
                        if (sq == 0) { sact[q] = 0; sslot[q] = nd_; sdd[nd_] = D[q]; nd_++; continue; }
                        if (pj < 0) { sact[q] = 1; sslot[q] = 0; pj = q; continue; }
                        if (sq == 1 && !dirty) { // speculated on the very data the walk sees
                            sact[q] = 3; sslot[q] = K;
                            sdk[K] = D[pj]; szk[K] = z[pj]; K++;
                            pj = q;
                            continue;
                        }
                        // dlaed2's test |t c s| <= tol with c = z_q / tau, s = -z_pj / tau, tau^2 = z_q^2 + z_pj^2, without
                        // forming them (no square root, no division unless a pole really deflates)
                        const double zp = z[pj], zj = z[q], t = D[q] - D[pj];
                        if (fabs(t * zj * zp) <= tol * (zj * zj + zp * zp)) { // two close poles: rotate the older one out
                            const double tau = sqrt(zj * zj + zp * zp);
                            const double c = zj / tau, sn = -zp / tau;
                            z[q] = tau;
                            z[pj] = 0.0;
                            const double tt = D[pj] * c * c + D[q] * sn * sn;
                            D[q] = D[pj] * sn * sn + D[q] * c * c;
                            D[pj] = tt;
                            sact[q] = 2; sslot[q] = nd_; src[q] = c; srs[q] = sn;
                            sdd[nd_] = tt; nd_++;
                            dirty = 1;
                        } else {
                            sact[q] = 3; sslot[q] = K;
                            sdk[K] = D[pj]; szk[K] = z[pj]; K++;
                            dirty = 0;
                        }
                        pj = q;
                    }
                }
                K = __shfl_sync(0xffffffffu, K, 0);
                nd_ = __shfl_sync(0xffffffffu, nd_, 0);
                pj = __shfl_sync(0xffffffffu, pj, 0);
                dirty = __shfl_sync(0xffffffffu, dirty, 0);
            }
            __syncwarp();
        }
        if (lane == 0) {
            if (pj >= 0) { w.lastslot[blockIdx.x] = K; sdk[K] = D[pj]; szk[K] = z[pj]; K++; }
            else w.lastslot[blockIdx.x] = -1;
            w.K[blockIdx.x] = K;
        }
    }
    __syncthreads();
    const int K = w.K[blockIdx.x];
    for (int j = tid; j < k; j += DC_THREADS) {
        w.act[o + j] = sact[j];
        w.slot[o + j] = sslot[j];
        w.rc[o + j] = src[j];
        w.rs[o + j] = srs[j];
        if (j < K) { w.dk[o + j] = sdk[j]; w.zk[o + j] = szk[j]; }
        else lam_new[o + j] = sdd[j - K]; // deflated eigenvalues follow the K roots
    }
}

// ---- merge: secular equation, one 128-thread CTA per root --------------------------------------------------------------
// (fp64 divisions are long dependency chains on this part: a warp per root left the chip at 9 % occupancy and 18 k cycles
// per evaluation; four warps per root and one root per CTA put ~6 warps on every scheduler.)
constexpr int DC_SEC_THREADS = 128;

// sums of three values over the CTA (one warp, or DC_SEC_THREADS threads); all threads get the same bits
__device__ __forceinline__ void dc_cta_sum3(double &a, double &b, double &c, double *red) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) { // the butterflies interleave
        a += __shfl_xor_sync(0xffffffffu, a, off);
        b += __shfl_xor_sync(0xffffffffu, b, off);
        c += __shfl_xor_sync(0xffffffffu, c, off);
    }
    if (blockDim.x == 32) return;
    __syncthreads(); // the previous sums have been read
    if ((threadIdx.x & 31) == 0) {
        const int q = threadIdx.x >> 5;
        red[q] = a; red[4 + q] = b; red[8 + q] = c;
    }
    __syncthreads();
    a = (red[0] + red[1]) + (red[2] + red[3]);
    b = (red[4] + red[5]) + (red[6] + red[7]);
    c = (red[8] + red[9]) + (red[10] + red[11]);
}

// 1 / x to ~1 ulp: the 20-bit hardware seed and two Newton steps (5 instructions against ~30 of the IEEE division; the
// root only has to be consistent with the Loewner z, not correctly rounded)
__device__ __forceinline__ double dc_rcp(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double err = fma(-x, r, 1.0);
    r = fma(r, err, r);
    err = fma(-x, r, 1.0);
    r = fma(r, err, r);
    return fabs(r) < 1.7976931348623157e308 ? r : 1.0 / x; // seed 0 / inf (subnormal or huge x): the exact path
}

// f, f' = rho sum z_j^2 / (d_j - lambda)^2 and the size of the terms f is the difference of (its rounding noise is
// ~eps * fabs_terms: the stopping test of dlaed4), one pass, one reciprocal per term
__device__ __forceinline__ double dc_secular_fd(const double *dk, const double *z2, int K, double dorg, double rho, double mu,
        double *red, double &fprime, double &fabs_terms) {
    double s = 0.0, s2 = 0.0, sa = 0.0, sb = 0.0, s2b = 0.0;
    const int T = blockDim.x;
    int j = threadIdx.x;
    for (; j + T < K; j += 2 * T) { // two reciprocals in flight
        const double r = dc_rcp((dk[j] - dorg) - mu), rb = dc_rcp((dk[j + T] - dorg) - mu);
        const double t = z2[j] * r, tb = z2[j + T] * rb;
        s += t;
        sb += tb;
        sa += fabs(t) + fabs(tb);
        s2 = fma(t, r, s2);
        s2b = fma(tb, rb, s2b);
    }
    for (; j < K; j += T) {
        const double r = dc_rcp((dk[j] - dorg) - mu);
        const double t = z2[j] * r;
        s += t;
        sa += fabs(t);
        s2 = fma(t, r, s2);
    }
    s += sb;
    s2 += s2b;
    dc_cta_sum3(s, s2, sa, red);
    fprime = rho * s2;
    fabs_terms = 1.0 + rho * sa;
    return 1.0 + rho * s;
}

__global__ void __launch_bounds__(DC_SEC_THREADS) dc_secular_kernel(const DcNode *nodes, DcWork w, double *lam_new) {
    extern __shared__ __align__(16) double sm[];
    __shared__ double red[12];
    const DcNode nd = nodes[blockIdx.x];
    const int o = nd.o, K = w.K[blockIdx.x], tid = threadIdx.x, i = blockIdx.y;
    if (i >= K) return;
    double *dk = sm, *z2 = sm + K;
    for (int j = tid; j < K; j += blockDim.x) {
        const double z = w.zk[o + j];
        dk[j] = w.dk[o + j];
        z2[j] = z * z;
    }
    __syncthreads();
    const double rho = w.rho[blockIdx.x];
    double fp, fa;
    int org;
    double lo, hi;
    bool left;
    double fstart, xstart; // f at the first evaluation point, and that point relative to the origin pole
    if (i < K - 1) {
        const double gap = dk[i + 1] - dk[i];
        fstart = dc_secular_fd(dk, z2, K, dk[i], rho, 0.5 * gap, red, fp, fa);
        left = fstart >= 0.0;
        org = left ? i : i + 1;
        lo = left ? 0.0 : -0.5 * gap;
        hi = left ? 0.5 * gap : 0.0;
        xstart = left ? 0.5 * gap : -0.5 * gap;
    } else {
        double zsum = 0.0, u1 = 0.0, u2 = 0.0;
        for (int j = tid; j < K; j += blockDim.x) zsum += z2[j];
        dc_cta_sum3(zsum, u1, u2, red);
        left = true;
        org = i;
        lo = 0.0;
        hi = rho * zsum; // the last root lies in (d_K-1, d_K-1 + rho sum z^2]
        xstart = hi;
        fstart = dc_secular_fd(dk, z2, K, dk[i], rho, hi, red, fp, fa);
    }
    const double dorg = dk[org];
    // f increases with mu: f(lo) < 0 <= f(hi); the zero end of the bracket is a pole.  Start: split f into the origin
    // pole's term -rho z_o^2 / mu and the rest g, which increases with mu; freezing g at the first evaluation point gives
    // mu0 = rho z_o^2 / g(xstart), which lies between the pole and the root -- the side from which Newton on f converges
    // monotonically (f is concave on the right of a pole, convex on its left).  Then safeguarded Newton: a step that
    // leaves the bracket, or three steps that do not halve, are replaced by a bisection.  When mu0 is unusable the bracket is
    // first shrunk geometrically until both ends are within a factor 16.  Every thread of the CTA takes the same decisions
    // from the same bits.
    const double pole = rho * z2[org];
    double mu0 = pole / (fstart + pole / xstart);
    const bool good = mu0 > lo && mu0 < hi;
    for (int it = 0; it < 400 && !good; it++) {
        double mid;
        const bool tight = left ? (lo > 0.0 && hi / lo < 16.0) : (hi < 0.0 && lo / hi < 16.0);
        if (tight) break;
        if (left) mid = lo > 0.0 ? sqrt(lo * hi) : hi * 0.0009765625;
        else mid = hi < 0.0 ? -sqrt(lo * hi) : lo * 0.0009765625;
        if (!(mid > lo && mid < hi)) break;
        if (dc_secular_fd(dk, z2, K, dorg, rho, mid, red, fp, fa) >= 0.0) hi = mid; else lo = mid;
    }
    double mu = good ? mu0 : 0.5 * (lo + hi), prev_step = 0.0;
    int slow = 0;
    for (int it = 0; it < 200; it++) {
        if (!(mu > lo && mu < hi)) { mu = 0.5 * (lo + hi); if (!(mu > lo && mu < hi)) break; }
        const double f = dc_secular_fd(dk, z2, K, dorg, rho, mu, red, fp, fa);
        if (f >= 0.0) hi = mu; else lo = mu;
        if (fabs(f) <= 8.0 * 2.220446049250313e-16 * fa) break; // f is zero to its rounding noise: no step can improve mu
        // Newton approaches the root from one side, so the bracket does not shrink while it converges: progress is
        // measured on the step (three oscillating steps that do not halve -> one bisection), not on the bracket
        const double step = -f / fp;
        double next = mu + step;
        slow = (fabs(step) > 0.5 * fabs(prev_step) && step * prev_step < 0.0) ? slow + 1 : 0; // oscillation, not a monotone approach
        prev_step = step;
        if (!(next > lo && next < hi) || slow >= 3) { next = 0.5 * (lo + hi); slow = 0; prev_step = 0.0; }
        if (next == mu || fabs(next - mu) <= 2.220446049250313e-16 * fabs(mu)) { mu = next; break; }
        mu = next;
    }
    if (tid == 0) {
        w.org[o + i] = org;
        w.mu[o + i] = mu;
        lam_new[o + i] = dorg + mu;
    }
}

// ---- merge: z from the roots (Loewner), one warp per component -----------------------------------------------------------
__global__ void __launch_bounds__(DC_THREADS) dc_zhat_kernel(const DcNode *nodes, DcWork w) {
    const DcNode nd = nodes[blockIdx.x];
    const int o = nd.o, K = w.K[blockIdx.x], tid = threadIdx.x, lane = tid & 31;
    const double rho = w.rho[blockIdx.x];
    const int warps = (DC_THREADS / 32) * gridDim.y, wg = blockIdx.y * (DC_THREADS / 32) + (tid >> 5);
    for (int j = wg; j < K; j += warps) {
        const double dj = w.dk[o + j];
        double p = 1.0;
        for (int i = lane; i < K; i += 32) {
            const double num = (w.dk[o + w.org[o + i]] - dj) + w.mu[o + i]; // lambda_i - d_j
            const double den = i == j ? rho : w.dk[o + i] - dj;
            p *= num / den;
        }
        for (int off = 16; off >= 1; off >>= 1) p *= __shfl_xor_sync(0xffffffffu, p, off);
        if (lane == 0) w.zhat[o + j] = copysign(sqrt(fabs(p)), w.zk[o + j]);
    }
}

// ---- merge: eigenvectors of the rank-one problem: Ut[i][j] = zhat_j / (d_j - lambda_i), rows normalised ---------------
__global__ void __launch_bounds__(DC_THREADS) dc_vectors_kernel(const DcNode *nodes, DcWork w, double *Ut, int ld) {
    const DcNode nd = nodes[blockIdx.x];
    const int o = nd.o, K = w.K[blockIdx.x], tid = threadIdx.x, lane = tid & 31;
    const int warps = (DC_THREADS / 32) * gridDim.y, wg = blockIdx.y * (DC_THREADS / 32) + (tid >> 5);
    for (int i = wg; i < K; i += warps) {
        const double dorg = w.dk[o + w.org[o + i]], mu = w.mu[o + i];
        double ss = 0.0;
        for (int j = lane; j < K; j += 32) {
            const double v = w.zhat[o + j] / ((w.dk[o + j] - dorg) - mu);
            ss += v * v;
        }
        ss = bb::warp_sum(ss);
        const double inv = 1.0 / sqrt(ss);
        for (int j = lane; j < K; j += 32)
            Ut[(size_t) (o + i) * ld + o + j] = w.zhat[o + j] / ((w.dk[o + j] - dorg) - mu) * inv;
    }
}

// ---- merge: permutation + Givens chains over the children's basis.  One thread per (row, segment of 64 sorted positions):
// the only state that crosses a segment boundary is the "carry" (the last pole that is neither deflated nor emitted yet),
// and it is recomputed at the segment start by walking back over the rotations that produced it (usually none).
constexpr int DC_SEG = 64;

__global__ void __launch_bounds__(DC_THREADS) dc_gather_kernel(const DcNode *nodes, DcWork w, const double *Zt, double *QselT,
        double *Zt_new, int ld) {
    const DcNode nd = nodes[blockIdx.x];
    const int o = nd.o, k = nd.k, K = w.K[blockIdx.x];
    const int r = blockIdx.y * DC_THREADS + threadIdx.x;
    const int j0 = blockIdx.z * DC_SEG, j1 = min(k, j0 + DC_SEG);
    if (r >= k || j0 >= k) return;
    const int *act = w.act + o, *slot = w.slot + o, *perm = w.perm + o;
    auto col = [&](int j) { return Zt[(size_t) (o + perm[j]) * ld + o + r]; };
    double carry = 0.0;
    {
        int p = j0 - 1;
        while (p >= 0 && act[p] == 0) p--; // the pole the carry belongs to
        if (p >= 0) {
            int q = p;
            while (q >= 0 && (act[q] == 2 || act[q] == 0)) q--; // the plain column the chain of rotations started from
            if (q >= 0) {
                carry = col(q);
                for (int t = q + 1; t <= p; t++)
                    if (act[t] == 2) carry = w.rc[o + t] * col(t) - w.rs[o + t] * carry;
            }
        }
    }
    for (int jb = j0; jb < j1; jb += 8) { // the loads of 8 positions are issued together: the walk is latency-bound otherwise
        double y[8];
        int a[8], sl[8];
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const int j = min(jb + u, k - 1);
            y[u] = col(j);
            a[u] = act[j];
            sl[u] = slot[j];
        }
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const int j = jb + u;
            if (j >= j1) break;
            if (a[u] == 0) Zt_new[(size_t) (o + K + sl[u]) * ld + o + r] = y[u];
            else if (a[u] == 1) carry = y[u];
            else if (a[u] == 2) {
                const double c = w.rc[o + j], s = w.rs[o + j];
                Zt_new[(size_t) (o + K + sl[u]) * ld + o + r] = c * carry + s * y[u];
                carry = c * y[u] - s * carry;
            } else {
                QselT[(size_t) (o + sl[u]) * ld + o + r] = carry;
                carry = y[u];
            }
        }
    }
    const int last = w.lastslot[blockIdx.x];
    if (j1 == k && last >= 0) QselT[(size_t) (o + last) * ld + o + r] = carry;
}

// ---- merge: Zt_new[o+i][o+r] = sum_j Ut[o+i][o+j] QselT[o+j][o+r], i, j < K (device-side), r < k -----------------------
// All merges of a level in ONE launch, with the live size K read on the device: deflation shrinks the product by (K / k)^2 and
// no operand has to be zero-padded.  32 x 32 output tiles (2 x 2 per thread), 32-deep steps, the next step's tiles are in
// flight (registers) while this one is multiplied, one barrier per step.  Up to k = 256 this beats one TMA/DMMA launch per
// merge (23 us against 4 x 12.6 us at k = 256 under ncu: the launches run one after the other and each multiplies the full
// k^3); at k = 504 it does not (72 us against 2 x 21.5), so the two upper levels keep the tensor-core GEMM.
constexpr int GS_T = 32, GS_LD = GS_T + 2;

__global__ void __launch_bounds__(256) dc_gemm_kernel(const DcNode *nodes, const int *Kdev, const double *Ut, const double *QselT,
        double *Zt_new, int ld) {
    const DcNode nd = nodes[blockIdx.z];
    const int o = nd.o, k = nd.k, K = Kdev[blockIdx.z];
    const int i0 = blockIdx.y * GS_T, r0 = blockIdx.x * GS_T;
    if (i0 >= K || r0 >= k) return;
    __shared__ __align__(16) double As[2][GS_T][GS_LD], Bs[2][GS_T][GS_LD]; // [step parity][j][i or r]
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int lj = tid & 31, li = tid >> 5; // loader: column lj of rows li, li + 8, li + 16, li + 24 of a tile
    const double *A = Ut + (size_t) o * ld + o, *B = QselT + (size_t) o * ld + o;
    double pa[4], pb[4];
    auto fetch = [&](int j0) {
#pragma unroll
        for (int t = 0; t < 4; t++) {
            const int ii = li + 8 * t;
            pa[t] = (i0 + ii < K && j0 + lj < K) ? A[(size_t) (i0 + ii) * ld + j0 + lj] : 0.0;  // A[i][j]
            pb[t] = (j0 + ii < K && r0 + lj < k) ? B[(size_t) (j0 + ii) * ld + r0 + lj] : 0.0;  // B[j][r]
        }
    };
    auto stash = [&](int buf) {
#pragma unroll
        for (int t = 0; t < 4; t++) {
            As[buf][lj][li + 8 * t] = pa[t];
            Bs[buf][li + 8 * t][lj] = pb[t];
        }
    };
    double c00 = 0.0, c01 = 0.0, c10 = 0.0, c11 = 0.0;
    const int nsteps = (K + GS_T - 1) / GS_T;
    fetch(0);
    stash(0);
    __syncthreads();
    for (int st = 0; st < nsteps; st++) {
        const int buf = st & 1;
        if (st + 1 < nsteps) fetch((st + 1) * GS_T);
#pragma unroll
        for (int jj = 0; jj < GS_T; jj++) {
            const double2 a = *reinterpret_cast<const double2*>(&As[buf][jj][ty * 2]);
            const double2 b = *reinterpret_cast<const double2*>(&Bs[buf][jj][tx * 2]);
            c00 = fma(a.x, b.x, c00); c01 = fma(a.x, b.y, c01);
            c10 = fma(a.y, b.x, c10); c11 = fma(a.y, b.y, c11);
        }
        if (st + 1 < nsteps) stash(buf ^ 1); // last read two steps ago: every thread is past that step's barrier
        __syncthreads();
    }
    const int i = i0 + ty * 2, r = r0 + tx * 2;
    double *out = Zt_new + (size_t) (o + i) * ld + o + r;
    if (i < K) {
        if (r < k) out[0] = c00;
        if (r + 1 < k) out[1] = c01;
    }
    if (i + 1 < K) {
        if (r < k) out[ld] = c10;
        if (r + 1 < k) out[ld + 1] = c11;
    }
}

// rows < K (device-side) of a node's product block into the new basis
__global__ void __launch_bounds__(DC_THREADS) dc_copyrows_kernel(const DcNode *nodes, const int *Kdev, const double *T, double *Zt_new, int ld) {
    const DcNode nd = nodes[blockIdx.z];
    const int i = blockIdx.y, r = blockIdx.x * DC_THREADS + threadIdx.x;
    if (i >= Kdev[blockIdx.z] || r >= nd.k) return;
    Zt_new[(size_t) (nd.o + i) * ld + nd.o + r] = T[(size_t) (nd.o + i) * ld + nd.o + r];
}

// ---- final: ascending order ----------------------------------------------------------------------------------------------
// one warp per eigenvalue (a thread per eigenvalue walked all n values alone: 25 us at n = 1000)
__global__ void __launch_bounds__(DC_THREADS) dc_rank_kernel(const double *lam, int n, int *pos, double *W) {
    const int j = blockIdx.x * (DC_THREADS / 32) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (j >= n) return;
    const double v = lam[j];
    int p = 0;
    for (int q = lane; q < n; q += 32) {
        const double u = lam[q];
        p += (u < v) || (u == v && q < j);
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) p += __shfl_xor_sync(0xffffffffu, p, off);
    if (lane == 0) {
        pos[j] = p;
        W[p] = v;
    }
}

__global__ void __launch_bounds__(DC_THREADS) dc_permute_kernel(const double *Zt, const int *pos, int n, int ld, double *out, int ldo) {
    const int j = blockIdx.y;
    const int r = blockIdx.x * DC_THREADS + threadIdx.x;
    if (r < n) out[(size_t) pos[j] * ldo + r] = Zt[(size_t) j * ld + r];
}

// rows of a leaf: DC_LEAF unless BBGPU_DC_LEAF (8, 16, 24: a comparison switch) asks for smaller ones
int dc_leaf_rows() {
    static const int rows = [] {
        const char *ev = getenv("BBGPU_DC_LEAF");
        const int v = ev ? atoi(ev) : DC_LEAF;
        return (v >= 8 && v <= DC_LEAF) ? v : DC_LEAF;
    }();
    return rows;
}

void dc_build_tree(int o, int k, int depth, std::vector<std::vector<DcNode>> &levels, std::vector<DcNode> &leaves, std::vector<int> &bounds,
        int &maxdepth) {
    if (k <= dc_leaf_rows()) {
        leaves.push_back({ o, k, 0 });
        return;
    }
    int m = ((k / 2 + 7) / 8) * 8;
    if (m >= k) m = k / 2;
    if ((int) levels.size() <= depth) levels.resize(depth + 1);
    levels[depth].push_back({ o, k, m });
    bounds.push_back(o + m);
    maxdepth = std::max(maxdepth, depth);
    dc_build_tree(o, m, depth + 1, levels, leaves, bounds, maxdepth);
    dc_build_tree(o + m, k - m, depth + 1, levels, leaves, bounds, maxdepth);
}

} // namespace

namespace bb {

// T = tridiag(d, e) (device, n and n-1 entries) -> W[n] ascending and Zt (n x ldz, row j = eigenvector of W[j]).
// Everything is enqueued on the ctx stream; no host synchronisation.  *d_info (device int, zeroed here) becomes 1 when a
// leaf's QL iteration fails (non-finite input).
int stedc_device(bbgpu_ctx *ctx, int n, const double *d, const double *e, double *W, double *Zt_out, int ldz, int *d_info) {
    cudaStream_t st = ctx->stream;
    std::vector<std::vector<DcNode>> levels;
    std::vector<DcNode> leaves;
    std::vector<int> bounds;
    int maxdepth = -1;
    dc_build_tree(0, n, 0, levels, leaves, bounds, maxdepth);
    const int ld = (n + 7) & ~7;
    // one scratch allocation: 4 matrices + per-position arrays + node tables
    size_t nodes_total = leaves.size();
    for (auto &lv : levels) nodes_total += lv.size();
    const size_t mat = (size_t) n * ld;
    const size_t bytes = (5 * mat + 12 * (size_t) n + 64) * sizeof(double) + (6 * (size_t) n + 64) * sizeof(int) +
            4096;
    void *sp = nullptr;
    BB_TRY(ctx->scratch(bytes, &sp));
    char *base = (char*) sp;
    auto take = [&](size_t b) { char *p = base; base += (b + 255) & ~(size_t) 255; return p; };
    double *ZA = (double*) take(mat * 8), *ZB = (double*) take(mat * 8), *Qs = (double*) take(mat * 8), *Ut = (double*) take(mat * 8),
            *Tb = (double*) take(mat * 8);
    DcWork w;
    w.Ds = (double*) take(n * 8); w.zs = (double*) take(n * 8); w.dk = (double*) take(n * 8); w.zk = (double*) take(n * 8);
    w.mu = (double*) take(n * 8); w.zhat = (double*) take(n * 8); w.rc = (double*) take(n * 8); w.rs = (double*) take(n * 8);
    w.ddefl = (double*) take(n * 8); w.rho = (double*) take(n * 8);
    double *lamA = (double*) take(n * 8), *lamB = (double*) take(n * 8), *dadj = (double*) take(n * 8);
    w.perm = (int*) take(n * 4); w.act = (int*) take(n * 4); w.slot = (int*) take(n * 4); w.org = (int*) take(n * 4);
    w.K = (int*) take(n * 4); w.lastslot = (int*) take(n * 4);
    int *pos = (int*) take(n * 4);
    // The node tables are a function of n alone: they live in their own allocation and are uploaded when n changes, so a
    // decomposition enqueues without a host synchronisation -- its ~35 launches are issued while the tridiagonalisation
    // is still running.
    const size_t nodes_bytes = (nodes_total * sizeof(DcNode) + 255) & ~(size_t) 255, tables_bytes = nodes_bytes + (bounds.size() + 8) * sizeof(int);
    bool upload = ctx->dc_tables == nullptr || ctx->dc_tables_n != n;
    if (ctx->dc_tables_bytes < tables_bytes) {
        upload = true;
        BB_CUDA(cudaStreamSynchronize(st));
        if (ctx->dc_tables) cudaFree(ctx->dc_tables);
        ctx->dc_tables = nullptr;
        ctx->dc_tables_bytes = 0;
        BB_CUDA(cudaMalloc(&ctx->dc_tables, tables_bytes));
        ctx->dc_tables_bytes = tables_bytes;
    }
    DcNode *d_nodes = (DcNode*) ctx->dc_tables;
    int *d_bounds = (int*) ((char*) ctx->dc_tables + nodes_bytes);
    // node tables: leaves first, then the merge levels from the deepest up
    std::vector<DcNode> table(leaves);
    std::vector<size_t> level_off(levels.size());
    for (int lv = (int) levels.size() - 1; lv >= 0; lv--) {
        level_off[lv] = table.size();
        table.insert(table.end(), levels[lv].begin(), levels[lv].end());
    }
    if (upload) {
        void *pin = nullptr;
        BB_CUDA(cudaStreamSynchronize(st)); // the pinned staging buffer; earlier decompositions still reading the old tables
        BB_TRY(ctx->pinned(table.size() * sizeof(DcNode) + (bounds.size() + 8) * sizeof(int), &pin));
        memcpy(pin, table.data(), table.size() * sizeof(DcNode));
        int *hb = (int*) ((char*) pin + table.size() * sizeof(DcNode));
        for (size_t q = 0; q < bounds.size(); q++) hb[q] = bounds[q];
        BB_CUDA(cudaMemcpyAsync(d_nodes, pin, table.size() * sizeof(DcNode), cudaMemcpyHostToDevice, st));
        if (!bounds.empty()) BB_CUDA(cudaMemcpyAsync(d_bounds, hb, bounds.size() * sizeof(int), cudaMemcpyHostToDevice, st));
        BB_CUDA(cudaStreamSynchronize(st)); // (the staging buffer may be reused by the caller right away)
        ctx->dc_tables_n = n;
    }
    BB_CUDA(cudaMemsetAsync(d_info, 0, sizeof(int), st));
    BB_CUDA(cudaMemsetAsync(ZA, 0, mat * 8, st));
    BB_CUDA(cudaMemsetAsync(ZB, 0, mat * 8, st)); // entries outside a node's block are never written: zero in both buffers
    dc_adjust_kernel<<<ceil_div(n, 256), 256, 0, st>>>(d, e, d_bounds, (int) bounds.size(), dadj, n);
    BB_TRY(launched(ctx, "dc_adjust_kernel"));
    const size_t leaf_smem = ((size_t) DC_LEAF * (DC_LEAF | 1) + 3 * DC_LEAF) * sizeof(double);
    BB_CUDA(cudaFuncSetAttribute(dc_leaf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) leaf_smem));
    dc_leaf_kernel<<<(int) leaves.size(), 32, leaf_smem, st>>>(d_nodes, dadj, e, ZA, ld, lamA, d_info);
    BB_TRY(launched(ctx, "dc_leaf_kernel"));
    double *Zc = ZA, *Zn = ZB, *lc = lamA, *ln = lamB;
    static const bool use_dmma = [] { const char *ev = getenv("BBGPU_DC_GEMM"); return !(ev && ev[0] == 'f'); }(); // f: plain FMA kernel
    // merges larger than this take the tensor-core GEMM (BBGPU_DC_GEMM=d: from 256 rows, the round-1 setting)
    static const int dmma_from = [] { const char *ev = getenv("BBGPU_DC_GEMM"); return (ev && ev[0] == 'd') ? 255 : 256; }();
    for (int lv = (int) levels.size() - 1; lv >= 0; lv--) {
        const int cnt = (int) levels[lv].size();
        const DcNode *nodes = d_nodes + level_off[lv];
        int kmax = 0;
        for (auto &nd : levels[lv]) kmax = std::max(kmax, nd.k);
        // rows outside this level's merges (none: every row belongs to exactly one node per level when the tree is
        // balanced; unbalanced leaves that are already final are copied through)
        // (a level whose nodes cover every row rewrites every block entry of the new basis: nothing to carry over)
        int covered = 0;
        for (auto &nd : levels[lv]) covered += nd.k;
        if (covered != n) BB_CUDA(cudaMemcpyAsync(Zn, Zc, mat * 8, cudaMemcpyDeviceToDevice, st));
        const bool tensor = kmax > dmma_from && use_dmma;
        if (tensor) {
            BB_CUDA(cudaMemsetAsync(Ut, 0, mat * 8, st));
            BB_CUDA(cudaMemsetAsync(Qs, 0, mat * 8, st));
        }
        BB_CUDA(cudaMemcpyAsync(ln, lc, (size_t) n * 8, cudaMemcpyDeviceToDevice, st));
        const size_t setup_smem = (size_t) kmax * (7 * sizeof(double) + 3 * sizeof(int));
        BB_CUDA(cudaFuncSetAttribute(dc_setup_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) std::max(setup_smem, (size_t) 1024)));
        dc_setup_kernel<<<cnt, DC_THREADS, setup_smem, st>>>(nodes, e, Zc, ld, lc, w, ln);
        BB_TRY(launched(ctx, "dc_setup_kernel"));
        const int chunks = std::max(1, std::min(512, ceil_div(kmax, DC_THREADS / 32))); // one root (component) per warp
        dc_secular_kernel<<<dim3(cnt, kmax), kmax <= 128 ? 32 : DC_SEC_THREADS, 2 * (size_t) kmax * 8, st>>>(nodes, w, ln); // <= 4 terms per lane: one warp, no CTA barriers
        BB_TRY(launched(ctx, "dc_secular_kernel"));
        dc_zhat_kernel<<<dim3(cnt, chunks), DC_THREADS, 0, st>>>(nodes, w);
        BB_TRY(launched(ctx, "dc_zhat_kernel"));
        dc_vectors_kernel<<<dim3(cnt, chunks), DC_THREADS, 0, st>>>(nodes, w, Ut, ld);
        BB_TRY(launched(ctx, "dc_vectors_kernel"));
        dc_gather_kernel<<<dim3(cnt, ceil_div(kmax, DC_THREADS), ceil_div(kmax, DC_SEG)), DC_THREADS, 0, st>>>(nodes, w, Zc, Qs, Zn, ld);
        BB_TRY(launched(ctx, "dc_gather_kernel"));
        if (tensor) {
            // the engine's DMMA GEMM wants its sizes on the host: run it over the whole k x k x k block of each merge (the
            // operands are zero beyond the K live rows / columns) and keep the first K rows of the product
            for (const DcNode &nd : levels[lv]) {
                const size_t off = (size_t) nd.o * ld + nd.o;
                GemmArgs g {};
                g.A = Ut + off; g.lda = ld; g.B = Qs + off; g.ldb = ld; g.M = nd.k; g.N = nd.k; g.K = nd.k; g.C = Tb + off; g.ldc = ld;
                BB_TRY((launch_gemm<1, EPI_STORE>(ctx, g)));
            }
            dc_copyrows_kernel<<<dim3(ceil_div(kmax, DC_THREADS), kmax, cnt), DC_THREADS, 0, st>>>(nodes, w.K, Tb, Zn, ld);
            BB_TRY(launched(ctx, "dc_copyrows_kernel"));
        } else {
            dc_gemm_kernel<<<dim3(ceil_div(kmax, GS_T), ceil_div(kmax, GS_T), cnt), 256, 0, st>>>(nodes, w.K, Ut, Qs, Zn, ld);
            BB_TRY(launched(ctx, "dc_gemm_kernel"));
        }
        std::swap(Zc, Zn);
        std::swap(lc, ln);
    }
    dc_rank_kernel<<<ceil_div(n, DC_THREADS / 32), DC_THREADS, 0, st>>>(lc, n, pos, W);
    BB_TRY(launched(ctx, "dc_rank_kernel"));
    dc_permute_kernel<<<dim3(ceil_div(n, DC_THREADS), n), DC_THREADS, 0, st>>>(Zc, pos, n, ld, Zt_out, ldz);
    return launched(ctx, "dc_permute_kernel");
}

} // namespace bb

extern "C" {

// the tridiagonal divide & conquer alone, host buffers (unit tests): d[n], e[n-1] -> w[n] ascending, Zt[n*n] row j = eigenvector j
int bbgpu_stedc_host(bbgpu_ctx *ctx, int n, const double *h_d, const double *h_e, double *h_w, double *h_Zt, float *ms) {
    BB_ARG(ctx && n >= 1 && h_d && h_w && h_Zt && (n == 1 || h_e));
    cudaStream_t st = ctx->stream;
    double *dd = nullptr, *de = nullptr, *dw = nullptr, *dz = nullptr;
    int *dinfo = nullptr;
    BB_TRY(bb::dev_alloc(&dd, (size_t) n));
    BB_TRY(bb::dev_alloc(&de, (size_t) n));
    BB_TRY(bb::dev_alloc(&dw, (size_t) n));
    BB_TRY(bb::dev_alloc(&dz, (size_t) n * n));
    BB_TRY(bb::dev_alloc(&dinfo, (size_t) 1));
    BB_CUDA(cudaMemcpyAsync(dd, h_d, (size_t) n * 8, cudaMemcpyHostToDevice, st));
    BB_CUDA(cudaMemsetAsync(de, 0, (size_t) n * 8, st));
    if (n > 1) BB_CUDA(cudaMemcpyAsync(de, h_e, (size_t) (n - 1) * 8, cudaMemcpyHostToDevice, st));
    BB_TRY(bb::stedc_device(ctx, n, dd, de, dw, dz, n, dinfo)); // warm-up: scratch allocation
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0, st);
    BB_TRY(bb::stedc_device(ctx, n, dd, de, dw, dz, n, dinfo));
    cudaEventRecord(e1, st);
    BB_CUDA(cudaEventSynchronize(e1));
    float t = 0.f;
    cudaEventElapsedTime(&t, e0, e1);
    if (ms) *ms = t;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    int info = 0;
    BB_CUDA(cudaMemcpyAsync(&info, dinfo, sizeof(int), cudaMemcpyDeviceToHost, st));
    BB_CUDA(cudaMemcpyAsync(h_w, dw, (size_t) n * 8, cudaMemcpyDeviceToHost, st));
    BB_CUDA(cudaMemcpyAsync(h_Zt, dz, (size_t) n * n * 8, cudaMemcpyDeviceToHost, st));
    BB_CUDA(cudaStreamSynchronize(st));
    cudaFree(dd); cudaFree(de); cudaFree(dw); cudaFree(dz); cudaFree(dinfo);
    if (info != 0) return bb::fail(BBGPU_ERR_SOLVER, "bbgpu_stedc_host: the QL iteration of a leaf did not converge");
    return BBGPU_OK;
}

} // extern "C"
