// Batched small-n CMA-ES engine: many INDEPENDENT runs (the restarts of IPOP/BIPOP batches, BASELINE config 5:
// 4096 runs at n = 40) advance together, ONE CTA PER RUN, whole generations inside one kernel.
//
// Reference being replaced, per run (paths under /root/reference/src/multivariate/cma):
//   optimize loop   base_cmaes.cpp:162-174      iterate           base_cmaes.cpp:150-156
//   sample          cmaes.cpp:65-80             evaluate + sort   base_cmaes.cpp:211-230
//   update          cmaes.cpp:82-142            updateSigma       base_cmaes.cpp:176-189
//   eig gate/clamps cmaes.cpp:229-283           tred2 / tql2      cmaes.cpp:285-381, 383-478
//   updateHistory   base_cmaes.cpp:191-209      converged         cmaes.cpp:151-227
//
// Why a second engine: with one handle per run (cma.cu) a generation at n = 40, lambda = 15 is ~10 kernel launches
// and a host round trip -- pure launch latency.  Here a run's whole state (B in shared memory; m, p_c, p_sigma, D in
// shared memory; C, C^-1/2 in L2) stays with its CTA for as many generations as the launch allows, the O(n) control
// logic (sigma, history rings, the nine stop tests, the lazy-decomposition gate) runs on the device, and the host only
// sees a status word per run when a base run stops (the restart schedule of IPOP/BIPOP stays on the host,
// host/batch.cpp).  The chip is filled by runs, not by candidates: ~7-12 CTAs per SM x 148 SMs run concurrently and hide
// each other's dependency chains (the QL iteration of the eigensolver is a scalar recurrence).
//
// The decomposition is the reference's own algorithm (Householder tridiagonalisation + implicit QL, same index
// conventions, ascending order) with the O(n^2) inner loops spread over the CTA: for well separated spectra the basis
// agrees with the reference's to rounding, sign included.  (Nearly isotropic C: the eigenbasis is not a function of C,
// SURVEY 7.3 -- trajectory parity tests replay the oracle's basis through bbgpu_cmab_adopt_basis.)
#include <algorithm>
#include <map>
#include <vector>

#include "objectives.cuh"
#include "philox.cuh"
#include "rank.cuh"
#include "eig_small.cuh"

namespace {

constexpr int NMAX = 64;            // B (n x n) lives in shared memory; lanes own coordinates (i, i + 32)
constexpr int HCAP = 512;           // capacity of the two fitness-history rings per run
#ifndef CMAB_THREADS
#define CMAB_THREADS 128
#endif
constexpr int THREADS = CMAB_THREADS;
#ifndef CMAB_MIN_CTAS
#define CMAB_MIN_CTAS 8
#endif
constexpr int MIN_CTAS = CMAB_MIN_CTAS; // 8: 64 registers per thread, 8 CTAs (runs) per SM
constexpr int GROUPS = THREADS / NMAX > 0 ? THREADS / NMAX : 1; // thread groups splitting the ranks of the weighted mean
static_assert(THREADS >= NMAX, "a thread per coordinate");
constexpr int NWARPS = THREADS / 32;
constexpr int KCMAX = 32;           // selected steps staged per chunk of the rank-mu update (as many as fit the ranking keys' space)
constexpr int SMALL_LAMBDA = 256;   // <= : counting rank; larger: bitonic network over (key, index)

enum { ST_RUNNING = 0, ST_CONVERGED = 1, ST_BUDGET = 2, ST_FAILED = 3 };

// one run's scalars (device resident; the kernel works on a shared-memory copy)
struct RunDev {
    // base-run configuration: BaseCmaes::init / Cmaes::init (base_cmaes.cpp:54-134, cmaes.cpp:44-63)
    int lambda, mu, mit, mfev, hlen, ik;
    double sigma0, tol, cc, cs, c1, cmu, mueff, damps, chi, eigenfreq;
    unsigned long long seed;
    const double *w;
    // evolving
    double sigma, fbest, fworst, fkth, pslen, fx_final;
    double ybw[4];
    int ibw[4];
    int it, fev, eigenlastev, flag, status, hcur, hfill, hsig, eig_due, eig_count, ql_iters, ran;
    unsigned long long gen;
};

struct StepArgs {
    RunDev *runs;
    const int *list;
    const long long *xoff; // per run: first candidate of its slice of the pool in this launch
    bbgpu_cmab_run_status *status;
    int n, ldv, bound, obj_id, quantum, max_gens, eig_mode, sort_cap;
    int phase;      // 0: whole generations, objective inside; 1: sample only; 2: from the fitness on (external objectives)
    int accumulate; // phases 1/2: keep counting generations_this_step across the launches of one step call
    double *xmean, *xold, *pc, *ps, *D, *C, *B, *isC, *xbest, *hbest, *hkth;
    const double *lower, *upper, *aux, *Z;
    double *X, *F;
    int *I;
};

__device__ __forceinline__ void tri_decode(int e, int &i, int &j) {
    // e = i (i + 1) / 2 + j, j <= i
    i = (int) ((sqrtf(8.0f * (float) e + 1.0f) - 1.0f) * 0.5f);
    while ((i + 1) * (i + 2) / 2 <= e) i++;
    while (i * (i + 1) / 2 > e) i--;
    j = e - i * (i + 1) / 2;
}

// C^-1/2 = B D^-1 B^T (cmaes.cpp:273-282) from Bt (shared, Bt[c * ldv + r] = B[r][c]) and D (shared)
__device__ void cta_invsqrt(const int n, const int ldv, const double *Vt, const double *vD, double *rD, double *isC) {
    if (threadIdx.x < n) rD[threadIdx.x] = 1.0 / vD[threadIdx.x];
    __syncthreads();
    const int ne = n * (n + 1) / 2;
    for (int e = threadIdx.x; e < ne; e += THREADS) {
        int i, j;
        tri_decode(e, i, j);
        double s0 = 0.0, s1 = 0.0;
        int k = 0;
        for (; k + 1 < n; k += 2) {
            s0 = fma(Vt[k * ldv + i] * rD[k], Vt[k * ldv + j], s0);
            s1 = fma(Vt[(k + 1) * ldv + i] * rD[k + 1], Vt[(k + 1) * ldv + j], s1);
        }
        if (k < n) s0 = fma(Vt[k * ldv + i] * rD[k], Vt[k * ldv + j], s0);
        const double sum = s0 + s1;
        isC[i * n + j] = sum;
        isC[j * n + i] = sum;
    }
}

// ------------------------------------------------------------------------------------------------------------------
// The step kernel: CTA b advances run list[b] by whole generations until it stops, has spent `quantum` evaluations in
// this launch, or has done max_gens generations.
// ------------------------------------------------------------------------------------------------------------------
template<bool LARGE>
__global__ void __launch_bounds__(THREADS, MIN_CTAS) cmab_step_kernel(const StepArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ RunDev S;
    __shared__ int ctl[4];
    const int n = a.n, ldv = a.ldv, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int run = a.list[blockIdx.x];
    RunDev *R = a.runs + run;
    double *Vt = (double*) smem_raw;
    double *xm = Vt + n * ldv, *xo = xm + n, *vpc = xo + n, *vps = vpc + n, *vD = vps + n, *ve = vD + n;
    double *xrow = ve + n;                 // NWARPS * n
    double *cw = xrow + NWARPS * n;        // KCMAX
    double *red = cw + KCMAX;              // 32
    unsigned long long *skey = (unsigned long long*) (red + 32);
    int *sidx = (int*) (skey + a.sort_cap);
    // the ranking keys are dead once the permutation exists: their space stages the recombination partials and the
    // selected steps of the covariance update (sort_cap >= 256 doubles)
    double *ysel = (double*) skey;
    const int KC = min(KCMAX, a.sort_cap / n);

    for (int i = tid; i < (int) (sizeof(RunDev) / 4); i += THREADS) ((int*) &S)[i] = ((const int*) R)[i];
    const size_t vo = (size_t) run * n, mo = (size_t) run * n * n;
    double *gC = a.C + mo, *gB = a.B + mo, *gisC = a.isC + mo;
    double *hb = a.hbest + (size_t) run * HCAP, *hk = a.hkth + (size_t) run * HCAP;
    for (int idx = tid; idx < n * n; idx += THREADS) {
        const int r = idx / n, c = idx - r * n;
        Vt[c * ldv + r] = gB[idx];
    }
    if (tid < n) {
        xm[tid] = a.xmean[vo + tid];
        xo[tid] = a.xold[vo + tid];
        vpc[tid] = a.pc[vo + tid];
        vps[tid] = a.ps[vo + tid];
        vD[tid] = a.D[vo + tid];
    }
    __syncthreads();
    if (tid == 0) {
        if (a.phase == 0 || (a.phase == 1 && !a.accumulate)) S.ran = 0;
        S.eig_due = 0;
    }
    __syncthreads();
    const int lam = S.lambda, mu = S.mu;
    const long long xoff = a.xoff[run];
    double *Xr = a.X + xoff * n, *Fr = a.F + xoff;
    int *Ir = a.I + xoff;
    const double *w = S.w;
    int spent = 0;

    for (int g = 0; g < a.max_gens; g++) {
        if (S.status != ST_RUNNING) break;
        if (S.fev >= S.mfev) { // the optimize loop's own exit (base_cmaes.cpp:166)
            __syncthreads();
            if (tid == 0) S.status = ST_BUDGET;
            __syncthreads();
            break;
        }
        if (g > 0 && spent >= a.quantum) break;
        const double sigma = S.sigma;
        const unsigned long long gen = S.gen, seed = S.seed;
        // ---- sample + evaluate: one warp per candidate; lane p draws the coordinate pair (2p, 2p+1), lanes own
        // the output coordinates (lane, lane + 32); x_i = m_i + sigma sum_j B_ij (D_j z_j), j ascending (cmaes.cpp:65-80)
        for (int k = wid; k < lam && a.phase != 2; k += NWARPS) {
            double d0 = 0.0, d1 = 0.0;
            const int c0 = 2 * lane;
            if (c0 < n) {
                double z0, z1;
                if (a.Z != nullptr) {
                    const double *zr = a.Z + (xoff + k) * n;
                    z0 = zr[c0];
                    z1 = c0 + 1 < n ? zr[c0 + 1] : 0.0;
                } else {
                    const bb::Philox4 r = bb::philox4x32_10((uint32_t) lane, (uint32_t) k, (uint32_t) gen,
                            (uint32_t) (gen >> 32), (uint32_t) seed, (uint32_t) (seed >> 32));
                    bb::normal_pair(r, z0, z1);
                }
                d0 = vD[c0] * z0;
                d1 = c0 + 1 < n ? vD[c0 + 1] * z1 : 0.0;
            }
            // D∘z of this candidate goes through the warp's row buffer: every j is then ONE broadcast load instead of
            // two shuffles of a double; lanes without a coordinate of their own redo coordinate n-1 (branch-free loop)
            double *xr = xrow + wid * n;
            if (c0 < n) xr[c0] = d0;
            if (c0 + 1 < n) xr[c0 + 1] = d1;
            __syncwarp();
            double acc0 = 0.0, acc1 = 0.0;
            const int i0 = lane, i1 = lane + 32;
            const bool on0 = i0 < n, on1 = i1 < n;
            {
                const double *b0 = Vt + min(i0, n - 1), *b1 = Vt + (on1 ? i1 : min(i0, n - 1));
                int j = 0;
                for (; j + 1 < n; j += 2) {
                    const double za = xr[j], zb = xr[j + 1];
                    acc0 = fma(b0[0], za, acc0);
                    acc1 = fma(b1[0], za, acc1);
                    acc0 = fma(b0[ldv], zb, acc0);
                    acc1 = fma(b1[ldv], zb, acc1);
                    b0 += 2 * ldv;
                    b1 += 2 * ldv;
                }
                if (j < n) {
                    const double za = xr[j];
                    acc0 = fma(b0[0], za, acc0);
                    acc1 = fma(b1[0], za, acc1);
                }
            }
            __syncwarp(); // the buffer now takes x
            if (on0) {
                double x = xm[i0] + sigma * acc0;
                if (a.bound) x = fmax(a.lower[i0], fmin(x, a.upper[i0]));
                xr[i0] = x;
                Xr[(size_t) k * n + i0] = x;
            }
            if (on1) {
                double x = xm[i1] + sigma * acc1;
                if (a.bound) x = fmax(a.lower[i1], fmin(x, a.upper[i1]));
                xr[i1] = x;
                Xr[(size_t) k * n + i1] = x;
            }
            __syncwarp();
            if (a.phase == 0) {
                const double f = bb::warp_objective(a.obj_id, n, xr, a.aux);
                if (lane == 0) {
                    Fr[k] = f;
                    skey[k] = bb::sortable_key(f);
                }
            }
            __syncwarp();
        }
        if (a.phase == 1) break; // the candidates are in the pool: the caller evaluates them, phase 2 takes over
        if (a.phase == 2)
            for (int k = tid; k < lam; k += THREADS) skey[k] = bb::sortable_key(Fr[k]);
        __syncthreads();
        // ---- rank ascending by (fitness, index) (base_cmaes.cpp:221)
        if (!LARGE) {
            for (int t = tid; t < lam; t += THREADS) {
                const unsigned long long key = skey[t];
                int cnt = 0;
                for (int j = 0; j < lam; j++) {
                    const unsigned long long kj = skey[j];
                    cnt += (kj < key) || (kj == key && j < t);
                }
                sidx[cnt] = t;
            }
            __syncthreads();
        } else {
            int P = 2;
            while (P < lam) P <<= 1;
            for (int t = tid; t < P; t += THREADS) {
                if (t >= lam) skey[t] = ~0ull;
                sidx[t] = t < lam ? t : 0x7fffffff;
            }
            __syncthreads();
            for (int k = 2; k <= P; k <<= 1)
                for (int j = k >> 1; j > 0; j >>= 1) {
                    for (int p = tid; p < (P >> 1); p += THREADS) {
                        const int lo = 2 * j * (p / j) + (p % j), hi = lo + j;
                        const bool asc = (lo & k) == 0;
                        const unsigned long long ka = skey[lo], kb = skey[hi];
                        const int ia = sidx[lo], ib = sidx[hi];
                        if (bb::tuple_greater(ka, (unsigned) ia, kb, (unsigned) ib) == asc) {
                            skey[lo] = kb; skey[hi] = ka;
                            sidx[lo] = ib; sidx[hi] = ia;
                        }
                    }
                    __syncthreads();
                }
        }
        for (int t = tid; t < lam; t += THREADS) Ir[t] = sidx[t];
        if (tid == 0) {
            S.fev += lam;
            const int q[4] = { 0, lam > 1 ? 1 : 0, lam > 1 ? lam - 2 : 0, lam - 1 };
            for (int u = 0; u < 4; u++) {
                S.ibw[u] = sidx[q[u]];
                S.ybw[u] = Fr[sidx[q[u]]];
            }
            S.fkth = Fr[sidx[min(S.ik, lam - 1)]];
        }
        // ---- updateDistribution (cmaes.cpp:82-149).  Mean: 4 groups of threads split the ranks, partials added in
        // group order.
        {
            const int i = tid % NMAX, grp = tid / NMAX, per = (mu + GROUPS - 1) / GROUPS;
            if (i < n) {
                double s = 0.0;
                const int k1 = min(mu, (grp + 1) * per);
                for (int k = grp * per; k < k1; k++) s = fma(w[k], Xr[(size_t) sidx[k] * n + i], s);
                ysel[grp * NMAX + i] = s;
            }
        }
        __syncthreads();
        if (tid < n) {
            double m = ysel[tid];
            for (int q = 1; q < GROUPS; q++) m += ysel[q * NMAX + tid];
            if (a.bound) m = fmax(a.lower[tid], fmin(m, a.upper[tid]));
            xo[tid] = xm[tid];
            xm[tid] = m;
        }
        __syncthreads();
        const double csc = sqrt(S.cs * (2.0 - S.cs) * S.mueff);
        double psq = 0.0;
        if (tid < n) { // p_sigma (cmaes.cpp:99-105); C^-1/2 is symmetric, column tid is read (coalesced)
            double p = vps[tid] * (1.0 - S.cs);
            for (int j = 0; j < n; j++) p += csc * gisC[(size_t) j * n + tid] * (xm[j] - xo[j]) / sigma;
            vps[tid] = p;
            psq = p * p;
        }
        const double pslen = sqrt(bb::block_sum(psq, red));
        const double denom = 1.0 - pow(1.0 - S.cs, 2.0 * S.fev / lam);
        const int hsig = (pslen / sqrt(denom) / S.chi < 1.4 + 2.0 / (n + 1.0)) ? 1 : 0;
        const double ccc = sqrt(S.cc * (2.0 - S.cc) * S.mueff);
        if (tid < n) vpc[tid] = (1.0 - S.cc) * vpc[tid] + hsig * ccc * (xm[tid] - xo[tid]) / sigma;
        __syncthreads();
        // covariance (cmaes.cpp:125-142): lower-triangle elements are thread-owned, the selected steps are staged KC at
        // a time; an element's sum runs over the ranks in ascending order, continued through C between the chunks
        {
            const int ne = n * (n + 1) / 2;
            const double c2 = (1.0 - hsig) * S.cc * (2.0 - S.cc), c1 = S.c1, cmu = S.cmu;
            for (int k0 = 0; k0 < mu; k0 += KC) {
                const int kc = min(KC, mu - k0);
                for (int idx = tid; idx < kc * n; idx += THREADS) {
                    const int kk = idx / n, i = idx - kk * n;
                    ysel[idx] = (Xr[(size_t) sidx[k0 + kk] * n + i] - xo[i]) / sigma;
                }
                if (tid < kc) cw[tid] = cmu * w[k0 + tid];
                __syncthreads();
                for (int e = tid; e < ne; e += THREADS) {
                    int i, j;
                    tri_decode(e, i, j);
                    double s = gC[i * n + j];
                    if (k0 == 0) s = (1.0 - c1 - cmu) * s + c1 * (vpc[i] * vpc[j] + c2 * s);
                    for (int kk = 0; kk < kc; kk++) s += cw[kk] * ysel[kk * n + i] * ysel[kk * n + j];
                    gC[i * n + j] = s;
                }
                __syncthreads();
            }
        }
        // updateSigma (base_cmaes.cpp:176-189) + the decomposition gate (cmaes.cpp:233)
        if (tid == 0) {
            S.pslen = pslen;
            S.hsig = hsig;
            double sg = sigma * exp(fmin(1.0, (S.cs / S.damps) * (pslen / S.chi - 1.0)));
            if (S.ybw[0] == S.fkth) sg *= exp(0.2 + S.cs / S.damps);
            if (S.it >= S.hlen && S.fworst - S.fbest == 0.0) sg *= exp(0.2 + S.cs / S.damps);
            S.sigma = sg;
            ctl[2] = 0;
            if ((double) (S.fev - S.eigenlastev) > S.eigenfreq) {
                S.eigenlastev = S.fev;
                S.eig_count++;
                if (a.eig_mode == 1) S.eig_due = 1; // replay: the host adopts the oracle's basis
                else ctl[2] = 1;
            }
        }
        __syncthreads();
        if (ctl[2]) {
            // eigenDecomposition (cmaes.cpp:236-282): symmetrise the lower triangle, tred2 + tql2, ascending order,
            // condition clamps, D = sqrt, C^-1/2
            for (int idx = tid; idx < n * n; idx += THREADS) {
                const int r = idx / n, c = idx - r * n;
                Vt[c * ldv + r] = gC[max(r, c) * n + min(r, c)];
            }
            __syncthreads();
            if (wid == 0) {
                int qi = 0;
                const bool okw = bb::warp_tred2_tql2(n, ldv, Vt, vD, ve, &qi, false);
                if (lane == 0) {
                    ctl[3] = okw ? 1 : 0;
                    if (okw) S.ql_iters += qi;
                }
            }
            __syncthreads();
            const bool ok = ctl[3] != 0;
            if (!ok) {
                __syncthreads();
                if (tid == 0) { S.status = ST_FAILED; S.flag = -1; }
                __syncthreads();
                break;
            }
            int pos = 0;
            double val = 0.0;
            if (tid < n) {
                val = vD[tid];
                for (int j = 0; j < n; j++) {
                    const double dj = vD[j];
                    pos += (dj < val) || (dj == val && j < tid);
                }
            }
            __syncthreads();
            if (tid < n) {
                ve[pos] = val;
                for (int i = 0; i < n; i++) gB[(size_t) i * n + pos] = Vt[tid * ldv + i];
            }
            __syncthreads();
            if (tid < n) { // condition clamps (cmaes.cpp:249-266) on the ascending spectrum, then D = sqrt
                double lo = ve[0], hi = ve[n - 1], v = ve[tid], shift = 0.0;
                if (lo <= 0.0) {
                    lo = fmax(lo, 0.0);
                    hi = fmax(hi, 0.0);
                    v = fmax(v, 0.0);
                    const double s1 = hi / 1e14;
                    lo += s1;
                    hi += s1;
                    v += s1;
                    shift += s1;
                }
                if (hi > 1e14 * lo) {
                    const double s2 = hi / 1e14 - lo;
                    v += s2;
                    shift += s2;
                }
                if (shift != 0.0) gC[tid * n + tid] += shift;
                vD[tid] = sqrt(v);
            }
            for (int idx = tid; idx < n * n; idx += THREADS) {
                const int r = idx / n, c = idx - r * n;
                Vt[c * ldv + r] = gB[idx];
            }
            __syncthreads();
            cta_invsqrt(n, ldv, Vt, vD, ve, gisC);
            __syncthreads();
        }
        // updateHistory (base_cmaes.cpp:191-209), _it++, converged() (cmaes.cpp:151-227): warp 0
        if (wid == 0) {
            const int hlen = S.hlen;
            int hcur = S.hcur, hfill = S.hfill;
            double fbest = S.fbest, fworst = S.fworst;
            if (S.it < S.mit) {
                hcur = (hcur + 1) % hlen;
                if (lane == 0) {
                    hb[hcur] = S.ybw[0];
                    hk[hcur] = S.fkth;
                }
                if (hfill < hlen) hfill++;
                __syncwarp();
                if (hfill == hlen) {
                    double lo = INFINITY, hi = -INFINITY;
                    for (int q = lane; q < hlen; q += 32) {
                        const double v = hb[q];
                        lo = fmin(lo, v);
                        hi = fmax(hi, v);
                    }
                    fbest = -bb::warp_max(-lo);
                    fworst = bb::warp_max(hi);
                }
            }
            const int it = S.it + 1;
            int flag = 0;
            if (it >= S.mit) flag = 1;
            else if (it >= hlen && fworst - fbest < S.tol) flag = 2;
            if (flag == 0 && hfill >= n) {
                int eq = 0;
                for (int q = lane; q < n; q += 32) {
                    const int at = (hlen + hcur - q) % hlen;
                    eq += hb[at] == hk[at];
                }
                for (int off = 16; off >= 1; off >>= 1) eq += __shfl_xor_sync(0xffffffffu, eq, off);
                if (3 * eq >= n && eq > 0) flag = 3;
            }
            if (flag == 0) {
                const double sg = S.sigma, rel = sg / S.sigma0;
                int iaxis = n - 1 - ((it - 1) % n);
                const double da = vD[iaxis];
                int tolx_all = 1, axis_all = 1, coor_any = 0;
                for (int q = lane; q < n; q += 32) {
                    const double sd = sqrt(gC[q * n + q]), m = xm[q];
                    if (fmax(vpc[q], sd) * sg / S.sigma0 >= S.tol) tolx_all = 0;
                    // row iaxis of the reference's _b (cmaes.cpp:211 indexes _b[iaxis][i])
                    if (m != m + 0.1 * sg * da * Vt[q * ldv + iaxis]) axis_all = 0;
                    if (m == m + 0.2 * sg * sd) coor_any = 1;
                }
                tolx_all = __all_sync(0xffffffffu, tolx_all);
                axis_all = __all_sync(0xffffffffu, axis_all);
                coor_any = __any_sync(0xffffffffu, coor_any);
                if (tolx_all) flag = 4;
                else if (rel > 1.0e20 * vD[n - 1]) flag = 5;
                else if (vD[n - 1] > 1.0e7 * vD[0]) flag = 7;
                else if (axis_all) flag = 8;
                else if (coor_any) flag = 9;
            }
            if (lane == 0) {
                S.hcur = hcur;
                S.hfill = hfill;
                S.fbest = fbest;
                S.fworst = fworst;
                S.it = it;
                S.gen = gen + 1;
                S.ran++;
                if (flag != 0) {
                    S.flag = flag;
                    S.status = ST_CONVERGED;
                }
            }
        }
        spent += lam;
        __syncthreads();
    }
    // ---- write back
    __syncthreads();
    if (a.phase == 1) { // nothing but the candidates changed (and the budget exit of the optimize loop)
        if (tid == 0) {
            R->status = S.status;
            R->ran = S.ran;
        }
        return;
    }
    if (tid < n) {
        a.xmean[vo + tid] = xm[tid];
        a.xold[vo + tid] = xo[tid];
        a.pc[vo + tid] = vpc[tid];
        a.ps[vo + tid] = vps[tid];
        a.D[vo + tid] = vD[tid];
        // bestSolution() (base_cmaes.cpp:232-238): the best candidate of the last generation, or the mean
        if (S.it <= 0) a.xbest[vo + tid] = xm[tid];
        else if (spent > 0) a.xbest[vo + tid] = Xr[(size_t) S.ibw[0] * n + tid]; // a generation ran in THIS launch
    }
    if (wid == 0 && S.it <= 0) { // the restart drivers evaluate f(solution) once more (bipop_cmaes.cpp:84)
        if (a.obj_id >= 0) {     // (an external objective: the host evaluates the mean, bbgpu_cmab_step)
            const double f = bb::warp_objective(a.obj_id, n, xm, a.aux);
            if (lane == 0) S.fx_final = f;
        }
    } else if (tid == 0 && spent > 0) {
        S.fx_final = S.ybw[0];
    }
    __syncthreads();
    for (int i = tid; i < (int) (sizeof(RunDev) / 4); i += THREADS) ((int*) R)[i] = ((const int*) &S)[i];
    if (tid == 0) {
        bbgpu_cmab_run_status st;
        st.status = S.status; st.flag = S.flag; st.it = S.it; st.fev = S.fev; st.lambda = S.lambda;
        st.eig_count = S.eig_count; st.eig_due = S.eig_due; st.hsig = S.hsig;
        st.sigma = S.sigma; st.fx_final = S.fx_final; st.pslen = S.pslen; st.fkth = S.fkth;
        st.hist_best = S.fbest; st.hist_worst = S.fworst;
        st.ql_iters = S.ql_iters; st.generations_this_step = S.ran;
        for (int u = 0; u < 4; u++) { st.ibw[u] = S.ibw[u]; st.ybw[u] = S.ybw[u]; }
        a.status[run] = st;
    }
}

// (re)start runs: BaseCmaes::init + Cmaes::init on the device state.  keep_offdiag reproduces the reference's re-init
// quirk on restarts of the SAME object (cmaes.cpp:53-60: only the diagonals of B and C are reset).
__global__ void __launch_bounds__(THREADS) cmab_start_kernel(RunDev *runs, const RunDev *fresh, const int *list, int n,
        const double *guess, int keep_offdiag, double *xmean, double *xold, double *pc, double *ps, double *D, double *C,
        double *B, double *isC, double *xbest, bbgpu_cmab_run_status *status) {
    const int run = list[blockIdx.x], tid = threadIdx.x;
    const size_t vo = (size_t) run * n, mo = (size_t) run * n * n;
    if (tid == 0) runs[run] = fresh[blockIdx.x];
    if (tid < n) {
        const double gx = guess[(size_t) blockIdx.x * n + tid];
        xmean[vo + tid] = gx;
        xold[vo + tid] = gx;
        xbest[vo + tid] = gx;
        pc[vo + tid] = 0.0;
        ps[vo + tid] = 0.0;
        D[vo + tid] = 1.0;
    }
    for (int idx = tid; idx < n * n; idx += THREADS) {
        const int r = idx / n, c = idx - r * n;
        if (r == c) {
            C[mo + idx] = 1.0;
            B[mo + idx] = 1.0;
        } else if (!keep_offdiag) {
            C[mo + idx] = 0.0;
            B[mo + idx] = 0.0;
        }
        isC[mo + idx] = r == c ? 1.0 : 0.0;
    }
    if (tid == 0) {
        bbgpu_cmab_run_status st;
        memset(&st, 0, sizeof(st));
        st.lambda = fresh[blockIdx.x].lambda;
        st.sigma = fresh[blockIdx.x].sigma;
        status[run] = st;
    }
}

// eigenbasis replay: rebuild C^-1/2 from an adopted (B, D)
__global__ void __launch_bounds__(THREADS) cmab_invsqrt_kernel(int n, int ldv, const double *B, const double *D, double *isC) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *Vt = (double*) smem_raw, *vD = Vt + n * ldv;
    for (int idx = threadIdx.x; idx < n * n; idx += THREADS) {
        const int r = idx / n, c = idx - r * n;
        Vt[c * ldv + r] = B[idx];
    }
    if (threadIdx.x < n) vD[threadIdx.x] = D[threadIdx.x];
    __syncthreads();
    cta_invsqrt(n, ldv, Vt, vD, vD + n, isC);
}

// standalone decomposition of count symmetric matrices (unit tests of the in-CTA solver)
__global__ void __launch_bounds__(THREADS) cmab_eig_kernel(int n, int ldv, const double *A, double *Bout, double *Dout, int *info) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int ctl[4];
    double *Vt = (double*) smem_raw, *d = Vt + n * ldv, *e = d + n, *tmp = e + n, *rc = tmp + n, *rs = rc + n, *srt = rs + n;
    const int tid = threadIdx.x;
    const double *gA = A + (size_t) blockIdx.x * n * n;
    for (int idx = tid; idx < n * n; idx += THREADS) {
        const int r = idx / n, c = idx - r * n;
        Vt[c * ldv + r] = gA[max(r, c) * n + min(r, c)];
    }
    __syncthreads();
    if (tid < 32) {
        int qi = 0;
        const bool okw = bb::warp_tred2_tql2(n, ldv, Vt, d, e, &qi, false);
        if (tid == 0) {
            ctl[3] = okw ? 1 : 0;
            info[blockIdx.x] = okw ? qi : -1;
        }
    }
    __syncthreads();
    if (ctl[3] == 0) return;
    int pos = 0;
    double val = 0.0;
    if (tid < n) {
        val = d[tid];
        for (int j = 0; j < n; j++) pos += (d[j] < val) || (d[j] == val && j < tid);
        srt[pos] = val;
        for (int i = 0; i < n; i++) Bout[(size_t) blockIdx.x * n * n + (size_t) i * n + pos] = Vt[tid * ldv + i];
    }
    __syncthreads();
    if (tid < n) Dout[(size_t) blockIdx.x * n + tid] = srt[tid];
}

size_t step_smem_bytes(int n, int ldv, int sort_cap) {
    size_t doubles = (size_t) n * ldv + 6 * (size_t) n + (size_t) NWARPS * n + KCMAX + 32;
    return doubles * 8 + (size_t) sort_cap * 12;
}

} // namespace

struct bbgpu_cmab {
    bbgpu_ctx *ctx = nullptr;
    int n = 0, ldv = 0, runs = 0, bound = 0;
    RunDev *d_runs = nullptr;
    std::vector<RunDev> h_runs; // host mirror of the configuration part (lambda, xoff); evolving fields are device-owned
    bbgpu_cmab_run_status *d_status = nullptr;
    double *xmean = nullptr, *xold = nullptr, *pc = nullptr, *ps = nullptr, *D = nullptr, *C = nullptr, *B = nullptr,
            *isC = nullptr, *xbest = nullptr, *hbest = nullptr, *hkth = nullptr, *lower = nullptr, *upper = nullptr;
    // candidate pool of the current launch
    double *X = nullptr, *F = nullptr, *Z = nullptr;
    int *I = nullptr;
    size_t pool_cap = 0, z_cap = 0;
    int *d_list = nullptr;      // 2 * runs ints (small class, large class)
    long long *d_xoff = nullptr; // runs
    std::vector<long long> h_xoff;
    RunDev *d_fresh = nullptr;  // staging for starts
    double *d_guess = nullptr;
    std::map<int, double*> weights; // lambda -> device table of mu weights (base_cmaes.cpp:94-100), carved from slabs:
    std::vector<double*> slabs;     // no cudaMalloc while a batch runs (it would stall on the other streams' kernels)
    size_t slab_left = 0;
    double *slab_next = nullptr;
    bool attr_small = false, attr_large = false;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr; // the large-lambda class runs on the ctx's side stream
    int last_ns = 0, last_nl = 0, last_maxlam = 0;    // lists of the last launch (phase 2 reuses its phase 1's)
    size_t last_cands = 0;
};

namespace {

int ensure_pool(bbgpu_cmab *h, size_t cands) {
    if (cands <= h->pool_cap) return BBGPU_OK;
    BB_CUDA(cudaStreamSynchronize(h->ctx->stream));
    if (h->X) cudaFree(h->X);
    if (h->F) cudaFree(h->F);
    if (h->I) cudaFree(h->I);
    h->X = nullptr; h->F = nullptr; h->I = nullptr; h->pool_cap = 0;
    // generous: growing frees and allocates, and both stall on whatever the other streams are running
    const size_t want = std::max(2 * cands, (size_t) h->runs * 128) + 1024;
    BB_TRY(bb::dev_alloc(&h->X, want * h->n));
    BB_TRY(bb::dev_alloc(&h->F, want));
    BB_TRY(bb::dev_alloc(&h->I, want));
    h->pool_cap = want;
    return BBGPU_OK;
}

int field_ptr(bbgpu_cmab *h, int run, int field, double **p, int64_t *len) {
    const size_t n = h->n, vo = (size_t) run * n, mo = (size_t) run * n * n;
    switch (field) {
    case BBGPU_CMA_XMEAN: *p = h->xmean + vo; *len = n; break;
    case BBGPU_CMA_XOLD: *p = h->xold + vo; *len = n; break;
    case BBGPU_CMA_PC: *p = h->pc + vo; *len = n; break;
    case BBGPU_CMA_PS: *p = h->ps + vo; *len = n; break;
    case BBGPU_CMA_D: *p = h->D + vo; *len = n; break;
    case BBGPU_CMA_C: *p = h->C + mo; *len = n * n; break;
    case BBGPU_CMA_B: *p = h->B + mo; *len = n * n; break;
    case BBGPU_CMA_INVSQRTC: *p = h->isC + mo; *len = n * n; break;
    default: return bb::fail(BBGPU_ERR_ARG, "bbgpu_cmab: field %d is not a per-run state field", field);
    }
    return BBGPU_OK;
}

} // namespace

extern "C" {

int bbgpu_cmab_limits(int n, int *max_n, int *max_lambda, int *max_hlen) {
    BB_ARG(n >= 2);
    if (max_n) *max_n = NMAX;
    if (max_hlen) *max_hlen = HCAP;
    if (max_lambda) {
        int lam = 0;
        if (n <= NMAX) // the bitonic ranking buffer (power of two >= lambda) must fit beside the run's state
            for (int P = 2; P <= (1 << 20) && step_smem_bytes(n, n | 1, P) <= (size_t) 220 * 1024; P <<= 1) lam = P;
        *max_lambda = lam;
    }
    return BBGPU_OK;
}

int bbgpu_cmab_create(bbgpu_ctx *ctx, int n, int runs, int bound, const double *h_lower, const double *h_upper,
                      bbgpu_cmab **out) {
    BB_ARG(ctx != nullptr && out != nullptr);
    *out = nullptr;
    BB_ARG(n >= 2 && n <= NMAX);
    BB_ARG(runs >= 1);
    BB_ARG(!bound || (h_lower != nullptr && h_upper != nullptr));
    BB_CUDA(cudaSetDevice(ctx->device));
    bbgpu_cmab *h = new bbgpu_cmab();
    h->ctx = ctx;
    h->n = n;
    h->ldv = n | 1;
    h->runs = runs;
    h->bound = bound ? 1 : 0;
    h->h_runs.resize(runs);
    h->h_xoff.assign(runs, 0);
    memset(h->h_runs.data(), 0, sizeof(RunDev) * runs);
    const size_t R = runs, vn = R * n, mn = R * n * n;
    int rc = BBGPU_OK;
    auto A = [&](double **p, size_t c) { if (rc == BBGPU_OK) rc = bb::dev_alloc(p, c); };
    if (rc == BBGPU_OK) rc = bb::dev_alloc(&h->d_runs, R);
    if (rc == BBGPU_OK) rc = bb::dev_alloc(&h->d_status, R);
    if (rc == BBGPU_OK) rc = bb::dev_alloc(&h->d_list, 2 * R);
    if (rc == BBGPU_OK) rc = bb::dev_alloc(&h->d_fresh, R);
    if (rc == BBGPU_OK) rc = bb::dev_alloc(&h->d_xoff, R);
    A(&h->xmean, vn); A(&h->xold, vn); A(&h->pc, vn); A(&h->ps, vn); A(&h->D, vn); A(&h->xbest, vn);
    A(&h->C, mn); A(&h->B, mn); A(&h->isC, mn);
    A(&h->hbest, R * HCAP); A(&h->hkth, R * HCAP); A(&h->d_guess, vn);
    A(&h->lower, n); A(&h->upper, n);
    if (rc != BBGPU_OK) { bbgpu_cmab_destroy(h); return rc; }
    cudaStream_t st = ctx->stream;
    cudaMemsetAsync(h->d_runs, 0, sizeof(RunDev) * R, st);
    cudaMemsetAsync(h->d_status, 0, sizeof(bbgpu_cmab_run_status) * R, st);
    cudaMemsetAsync(h->C, 0, mn * 8, st);
    cudaMemsetAsync(h->B, 0, mn * 8, st);
    cudaMemsetAsync(h->hbest, 0, R * HCAP * 8, st);
    cudaMemsetAsync(h->hkth, 0, R * HCAP * 8, st);
    if (h_lower && h_upper) {
        cudaMemcpyAsync(h->lower, h_lower, n * 8, cudaMemcpyHostToDevice, st);
        cudaMemcpyAsync(h->upper, h_upper, n * 8, cudaMemcpyHostToDevice, st);
    }
    BB_CUDA(cudaStreamSynchronize(st));
    *out = h;
    return BBGPU_OK;
}

int bbgpu_cmab_destroy(bbgpu_cmab *h) {
    if (!h) return BBGPU_OK;
    cudaStreamSynchronize(h->ctx->stream);
    void *ptrs[] = { h->d_runs, h->d_status, h->d_list, h->d_fresh, h->d_xoff, h->xmean, h->xold, h->pc, h->ps, h->D, h->xbest, h->C,
            h->B, h->isC, h->hbest, h->hkth, h->d_guess, h->lower, h->upper, h->X, h->F, h->I, h->Z };
    for (void *p : ptrs) if (p) cudaFree(p);
    for (double *sl : h->slabs) cudaFree(sl);
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    if (h->ev_join) cudaEventDestroy(h->ev_join);
    delete h;
    return BBGPU_OK;
}

int bbgpu_cmab_start(bbgpu_cmab *h, int count, const int32_t *run_ids, const bbgpu_cmab_run_params *params,
                     const double *h_guess, int keep_offdiag) {
    BB_ARG(h != nullptr && count >= 0 && count <= h->runs);
    if (count == 0) return BBGPU_OK;
    BB_ARG(run_ids != nullptr && params != nullptr && h_guess != nullptr);
    bbgpu_ctx *ctx = h->ctx;
    cudaStream_t st = ctx->stream;
    const int n = h->n;
    // stage in pinned memory: [count RunDev][count ints][count * n doubles]
    const size_t b_runs = sizeof(RunDev) * count, b_list = bb::round_up(sizeof(int) * count, 8), b_guess = (size_t) count * n * 8;
    void *pin = nullptr;
    BB_CUDA(cudaStreamSynchronize(st)); // the pinned staging buffer may still be in flight
    size_t wnew = 0; // weights of the lambdas seen for the first time in this call
    {
        std::map<int, int> fresh;
        for (int q = 0; q < count; q++)
            if (params[q].lambda >= 2 && params[q].mu >= 1 && !h->weights.count(params[q].lambda)) fresh[params[q].lambda] = params[q].mu;
        for (auto &kv : fresh) wnew += (size_t) kv.second;
    }
    BB_TRY(ctx->pinned(b_runs + b_list + b_guess + wnew * sizeof(double), &pin));
    RunDev *fr = (RunDev*) pin;
    int *ls = (int*) ((char*) pin + b_runs);
    double *gs = (double*) ((char*) pin + b_runs + b_list);
    double *ws = (double*) ((char*) pin + b_runs + b_list + b_guess);
    size_t wused = 0;
    for (int q = 0; q < count; q++) {
        const bbgpu_cmab_run_params &p = params[q];
        const int run = run_ids[q];
        BB_ARG(run >= 0 && run < h->runs);
        BB_ARG(p.lambda >= 2 && p.mu >= 1 && p.mu <= p.lambda && p.weights != nullptr);
        const int hlen = p.hlen;
        if (hlen < 1 || hlen > HCAP)
            return bb::fail(BBGPU_ERR_ARG, "bbgpu_cmab_start: history length %d outside [1, %d]", hlen, HCAP);
        auto it = h->weights.find(p.lambda);
        if (it == h->weights.end()) { // first run at this lambda: a table from the slab, uploaded with the rest of this call
            const size_t need = ((size_t) p.mu + 1) & ~(size_t) 1;
            if (need > h->slab_left) {
                const size_t slab = std::max(need, (size_t) 1 << 21); // 16 MB
                double *sl = nullptr;
                BB_TRY(bb::dev_alloc(&sl, slab));
                h->slabs.push_back(sl);
                h->slab_next = sl;
                h->slab_left = slab;
            }
            double *dw = h->slab_next;
            h->slab_next += need;
            h->slab_left -= need;
            memcpy(ws + wused, p.weights, sizeof(double) * p.mu);
            BB_CUDA(cudaMemcpyAsync(dw, ws + wused, sizeof(double) * p.mu, cudaMemcpyHostToDevice, st));
            wused += p.mu;
            it = h->weights.emplace(p.lambda, dw).first;
        }
        RunDev r;
        memset(&r, 0, sizeof(r));
        r.lambda = p.lambda; r.mu = p.mu; r.mit = p.mit; r.mfev = p.mfev; r.hlen = hlen; r.ik = p.ik;
        r.sigma0 = p.sigma0; r.tol = p.tol; r.cc = p.cc; r.cs = p.cs; r.c1 = p.c1; r.cmu = p.cmu; r.mueff = p.mueff;
        r.damps = p.damps; r.chi = p.chi; r.eigenfreq = p.eigenfreq; r.seed = p.seed; r.w = it->second;
        r.sigma = p.sigma0;
        r.fbest = -INFINITY; // base_cmaes.cpp:131-132
        r.fworst = INFINITY;
        r.hcur = -1;
        r.status = ST_RUNNING;
        fr[q] = r;
        ls[q] = run;
        h->h_runs[run] = r;
        memcpy(gs + (size_t) q * n, h_guess + (size_t) q * n, sizeof(double) * n);
    }
    BB_CUDA(cudaMemcpyAsync(h->d_fresh, fr, b_runs, cudaMemcpyHostToDevice, st));
    BB_CUDA(cudaMemcpyAsync(h->d_list, ls, sizeof(int) * count, cudaMemcpyHostToDevice, st));
    BB_CUDA(cudaMemcpyAsync(h->d_guess, gs, b_guess, cudaMemcpyHostToDevice, st));
    cmab_start_kernel<<<count, THREADS, 0, st>>>(h->d_runs, h->d_fresh, h->d_list, n, h->d_guess, keep_offdiag, h->xmean,
            h->xold, h->pc, h->ps, h->D, h->C, h->B, h->isC, h->xbest, h->d_status);
    BB_TRY(bb::launched(ctx, "cmab_start_kernel"));
    BB_CUDA(cudaStreamSynchronize(st)); // staging buffers are reused by the next call
    return BBGPU_OK;
}

static bool g_status_needs_mean(const std::vector<bbgpu_cmab_run_status> &st, int run) {
    return st[run].status != ST_RUNNING && st[run].it <= 0;
}

// one launch (pair) over the listed runs; phase as in StepArgs.  The candidate pool slices are (re)assigned here.
static int cmab_launch(bbgpu_cmab *h, int obj_id, const double *aux, int count, const int32_t *run_ids, int quantum,
        int max_gens, const double *h_z, int eig_mode, int phase, int accumulate, bool assign, size_t *cands_out) {
    bbgpu_ctx *ctx = h->ctx;
    cudaStream_t st = ctx->stream;
    const int n = h->n;
    int ns = 0, nl = 0, maxlam = 0;
    size_t cands = 0;
    if (assign) {
        BB_CUDA(cudaStreamSynchronize(st));
        void *pin = nullptr;
        BB_TRY(ctx->pinned(sizeof(int) * 2 * (size_t) count + sizeof(long long) * (size_t) h->runs, &pin));
        int *small = (int*) pin, *large = small + count;
        long long *px = (long long*) (large + count);
        for (int q = 0; q < count; q++) {
            const int run = run_ids[q];
            BB_ARG(run >= 0 && run < h->runs);
            const int lam = h->h_runs[run].lambda;
            if (lam < 2) return bb::fail(BBGPU_ERR_ARG, "bbgpu_cmab_step: run %d was never started", run);
            h->h_xoff[run] = (long long) cands;
            cands += lam;
            maxlam = std::max(maxlam, lam);
            if (lam <= SMALL_LAMBDA) small[ns++] = run; else large[nl++] = run;
        }
        memcpy(px, h->h_xoff.data(), sizeof(long long) * h->runs);
        BB_TRY(ensure_pool(h, cands));
        if (h_z != nullptr) {
            if (cands * n > h->z_cap) {
                if (h->Z) cudaFree(h->Z);
                h->Z = nullptr;
                BB_TRY(bb::dev_alloc(&h->Z, cands * n));
                h->z_cap = cands * n;
            }
            BB_CUDA(cudaMemcpyAsync(h->Z, h_z, cands * n * 8, cudaMemcpyHostToDevice, st));
        }
        BB_CUDA(cudaMemcpyAsync(h->d_xoff, px, sizeof(long long) * h->runs, cudaMemcpyHostToDevice, st));
        BB_CUDA(cudaMemcpyAsync(h->d_list, small, sizeof(int) * ns, cudaMemcpyHostToDevice, st));
        BB_CUDA(cudaMemcpyAsync(h->d_list + h->runs, large, sizeof(int) * nl, cudaMemcpyHostToDevice, st));
        h->last_ns = ns; h->last_nl = nl; h->last_maxlam = maxlam; h->last_cands = cands;
    } else { // phase 2 reuses the lists and pool slices of its phase 1
        ns = h->last_ns; nl = h->last_nl; maxlam = h->last_maxlam; cands = h->last_cands;
    }
    if (cands_out) *cands_out = cands;
    StepArgs a;
    a.runs = h->d_runs; a.status = h->d_status; a.xoff = h->d_xoff;
    a.n = n; a.ldv = h->ldv; a.bound = h->bound; a.obj_id = obj_id; a.quantum = quantum; a.max_gens = max_gens;
    a.eig_mode = eig_mode; a.phase = phase; a.accumulate = accumulate;
    a.xmean = h->xmean; a.xold = h->xold; a.pc = h->pc; a.ps = h->ps; a.D = h->D; a.C = h->C; a.B = h->B; a.isC = h->isC;
    a.xbest = h->xbest; a.hbest = h->hbest; a.hkth = h->hkth;
    a.lower = h->lower; a.upper = h->upper; a.aux = aux; a.Z = h_z ? h->Z : nullptr;
    a.X = h->X; a.F = h->F; a.I = h->I;
    bb::ProfScope prof(ctx, BBGPU_PROF_CMAB_STEP);
    if (nl > 0) {
        int P = 2;
        while (P < maxlam) P <<= 1;
        a.list = h->d_list + h->runs;
        a.sort_cap = P;
        const size_t smem = step_smem_bytes(n, h->ldv, a.sort_cap);
        if (smem > 220 * 1024)
            return bb::fail(BBGPU_ERR_ARG, "bbgpu_cmab_step: lambda = %d needs %zu bytes of shared memory for the ranking", maxlam, smem);
        BB_CUDA(cudaFuncSetAttribute(cmab_step_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
        // independent runs: the large class goes to the side stream and overlaps the small class (fork after the
        // uploads above, join before anything later on the main stream)
        if (!ctx->aux) BB_CUDA(cudaStreamCreateWithFlags(&ctx->aux, cudaStreamNonBlocking));
        if (!h->ev_fork) {
            BB_CUDA(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
            BB_CUDA(cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
        }
        BB_CUDA(cudaEventRecord(h->ev_fork, st));
        BB_CUDA(cudaStreamWaitEvent(ctx->aux, h->ev_fork, 0));
        cmab_step_kernel<true><<<nl, THREADS, smem, ctx->aux>>>(a);
        BB_TRY(bb::launched(ctx, "cmab_step_kernel<large>"));
        BB_CUDA(cudaEventRecord(h->ev_join, ctx->aux));
    }
    if (ns > 0) {
        a.list = h->d_list;
        a.sort_cap = SMALL_LAMBDA;
        const size_t smem = step_smem_bytes(n, h->ldv, a.sort_cap);
        if (!h->attr_small) {
            BB_CUDA(cudaFuncSetAttribute(cmab_step_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
            h->attr_small = true;
        }
        cmab_step_kernel<false><<<ns, THREADS, smem, st>>>(a);
        BB_TRY(bb::launched(ctx, "cmab_step_kernel<small>"));
    }
    if (nl > 0) BB_CUDA(cudaStreamWaitEvent(st, h->ev_join, 0));
    return BBGPU_OK;
}

int bbgpu_cmab_step(bbgpu_cmab *h, bbgpu_objective *obj, int count, const int32_t *run_ids, int quantum, int max_gens,
                    const double *h_z, int eig_mode) {
    BB_ARG(h != nullptr && obj != nullptr && count >= 0 && count <= h->runs);
    if (count == 0) return BBGPU_OK;
    BB_ARG(run_ids != nullptr && max_gens >= 1 && quantum >= 0);
    BB_ARG(obj->n == h->n);
    BB_ARG(h_z == nullptr || max_gens == 1);
    if (obj->mode == 0) // registry objective: evaluated inside the kernel, whole generations per launch
        return cmab_launch(h, obj->id, obj->d_aux, count, run_ids, quantum, max_gens, h_z, eig_mode, 0, 0, true, nullptr);
    // User CUDA objective / batched host callback: a generation is sample (phase 1) -> ONE evaluate call over the
    // candidates of all listed runs -> rank ... converged() (phase 2).  The host has to see which runs are still
    // running before it sizes the next evaluate call, so there is one status read per generation.
    bbgpu_ctx *ctx = h->ctx;
    cudaStream_t st = ctx->stream;
    std::vector<int32_t> active(run_ids, run_ids + count), next;
    std::vector<int> spent(h->runs, 0);
    std::vector<bbgpu_cmab_run_status> status(h->runs);
    for (int g = 0; g < max_gens && !active.empty(); g++) {
        size_t cands = 0;
        BB_TRY(cmab_launch(h, -1, nullptr, (int) active.size(), active.data(), quantum, 1, h_z, eig_mode, 1, g > 0 ? 1 : 0, true, &cands));
        BB_TRY(bb::evaluate(ctx, obj, h->X, h->n, (int) cands, h->F));
        BB_TRY(cmab_launch(h, -1, nullptr, (int) active.size(), active.data(), quantum, 1, h_z, eig_mode, 2, 1, false, nullptr));
        BB_CUDA(cudaMemcpyAsync(status.data(), h->d_status, sizeof(bbgpu_cmab_run_status) * h->runs, cudaMemcpyDeviceToHost, st));
        BB_CUDA(cudaStreamSynchronize(st));
        next.clear();
        for (int32_t run : active) {
            spent[run] += h->h_runs[run].lambda;
            if (status[run].status == ST_RUNNING && spent[run] < quantum) next.push_back(run);
        }
        active.swap(next);
    }
    // runs that never iterated (budget <= 0) report f(mean): the kernel cannot evaluate an external objective
    for (int q = 0; q < count; q++) {
        const int run = run_ids[q];
        if (g_status_needs_mean(status, run)) {
            BB_TRY(bb::evaluate(ctx, obj, h->xmean + (size_t) run * h->n, h->n, 1, h->F));
            BB_CUDA(cudaMemcpyAsync((char*) (h->d_status + run) + offsetof(bbgpu_cmab_run_status, fx_final), h->F, sizeof(double),
                    cudaMemcpyDeviceToDevice, st));
        }
    }
    return BBGPU_OK;
}

int bbgpu_cmab_fetch(bbgpu_cmab *h, bbgpu_cmab_run_status *h_out) {
    BB_ARG(h != nullptr && h_out != nullptr);
    BB_CUDA(cudaMemcpyAsync(h_out, h->d_status, sizeof(bbgpu_cmab_run_status) * h->runs, cudaMemcpyDeviceToHost, h->ctx->stream));
    BB_CUDA(cudaStreamSynchronize(h->ctx->stream));
    return BBGPU_OK;
}

int bbgpu_cmab_get_best(bbgpu_cmab *h, int run0, int count, double *h_x) {
    BB_ARG(h != nullptr && h_x != nullptr && run0 >= 0 && count >= 0 && run0 + count <= h->runs);
    BB_CUDA(cudaMemcpyAsync(h_x, h->xbest + (size_t) run0 * h->n, sizeof(double) * count * h->n, cudaMemcpyDeviceToHost,
            h->ctx->stream));
    BB_CUDA(cudaStreamSynchronize(h->ctx->stream));
    return BBGPU_OK;
}

int bbgpu_cmab_get(bbgpu_cmab *h, int run, int field, double *h_dst, int64_t count) {
    BB_ARG(h != nullptr && h_dst != nullptr && run >= 0 && run < h->runs);
    double *p = nullptr;
    int64_t len = 0;
    BB_TRY(field_ptr(h, run, field, &p, &len));
    BB_ARG(count == len);
    BB_CUDA(cudaMemcpyAsync(h_dst, p, sizeof(double) * len, cudaMemcpyDeviceToHost, h->ctx->stream));
    BB_CUDA(cudaStreamSynchronize(h->ctx->stream));
    return BBGPU_OK;
}

int bbgpu_cmab_set(bbgpu_cmab *h, int run, int field, const double *h_src, int64_t count) {
    BB_ARG(h != nullptr && h_src != nullptr && run >= 0 && run < h->runs);
    double *p = nullptr;
    int64_t len = 0;
    BB_TRY(field_ptr(h, run, field, &p, &len));
    BB_ARG(count == len);
    BB_CUDA(cudaMemcpyAsync(p, h_src, sizeof(double) * len, cudaMemcpyHostToDevice, h->ctx->stream));
    BB_CUDA(cudaStreamSynchronize(h->ctx->stream));
    return BBGPU_OK;
}

int bbgpu_cmab_set_sigma(bbgpu_cmab *h, int run, double sigma) {
    BB_ARG(h != nullptr && run >= 0 && run < h->runs);
    BB_CUDA(cudaMemcpyAsync((char*) (h->d_runs + run) + offsetof(RunDev, sigma), &sigma, sizeof(double), cudaMemcpyHostToDevice,
            h->ctx->stream));
    BB_CUDA(cudaStreamSynchronize(h->ctx->stream));
    return BBGPU_OK;
}

int bbgpu_cmab_get_population(bbgpu_cmab *h, int run, double *h_X, double *h_f, int32_t *h_rank) {
    BB_ARG(h != nullptr && run >= 0 && run < h->runs);
    const RunDev &r = h->h_runs[run];
    BB_ARG(r.lambda >= 2 && h->X != nullptr);
    cudaStream_t st = h->ctx->stream;
    if (h_X) BB_CUDA(cudaMemcpyAsync(h_X, h->X + h->h_xoff[run] * h->n, sizeof(double) * r.lambda * h->n, cudaMemcpyDeviceToHost, st));
    if (h_f) BB_CUDA(cudaMemcpyAsync(h_f, h->F + h->h_xoff[run], sizeof(double) * r.lambda, cudaMemcpyDeviceToHost, st));
    if (h_rank) BB_CUDA(cudaMemcpyAsync(h_rank, h->I + h->h_xoff[run], sizeof(int) * r.lambda, cudaMemcpyDeviceToHost, st));
    BB_CUDA(cudaStreamSynchronize(st));
    return BBGPU_OK;
}

int bbgpu_cmab_adopt_basis(bbgpu_cmab *h, int run, const double *h_B, const double *h_D) {
    BB_ARG(h != nullptr && h_B != nullptr && h_D != nullptr && run >= 0 && run < h->runs);
    const int n = h->n;
    cudaStream_t st = h->ctx->stream;
    const size_t vo = (size_t) run * n, mo = (size_t) run * n * n;
    BB_CUDA(cudaMemcpyAsync(h->B + mo, h_B, sizeof(double) * n * n, cudaMemcpyHostToDevice, st));
    BB_CUDA(cudaMemcpyAsync(h->D + vo, h_D, sizeof(double) * n, cudaMemcpyHostToDevice, st));
    cmab_invsqrt_kernel<<<1, THREADS, ((size_t) n * h->ldv + 2 * n) * 8, st>>>(n, h->ldv, h->B + mo, h->D + vo, h->isC + mo);
    BB_TRY(bb::launched(h->ctx, "cmab_invsqrt_kernel"));
    BB_CUDA(cudaStreamSynchronize(st));
    return BBGPU_OK;
}

int bbgpu_eig_small_host(bbgpu_ctx *ctx, int n, int count, const double *h_A, double *h_B, double *h_D, int32_t *h_info) {
    BB_ARG(ctx != nullptr && n >= 2 && n <= NMAX && count >= 1 && h_A && h_B && h_D && h_info);
    cudaStream_t st = ctx->stream;
    const size_t mn = (size_t) count * n * n;
    double *dA = nullptr, *dB = nullptr, *dD = nullptr;
    int *dinfo = nullptr;
    BB_TRY(bb::dev_alloc(&dA, mn));
    BB_TRY(bb::dev_alloc(&dB, mn));
    BB_TRY(bb::dev_alloc(&dD, (size_t) count * n));
    BB_TRY(bb::dev_alloc(&dinfo, (size_t) count));
    BB_CUDA(cudaMemcpyAsync(dA, h_A, mn * 8, cudaMemcpyHostToDevice, st));
    const int ldv = n | 1;
    cmab_eig_kernel<<<count, THREADS, ((size_t) n * ldv + 6 * n) * 8, st>>>(n, ldv, dA, dB, dD, dinfo);
    BB_TRY(bb::launched(ctx, "cmab_eig_kernel"));
    BB_CUDA(cudaMemcpyAsync(h_B, dB, mn * 8, cudaMemcpyDeviceToHost, st));
    BB_CUDA(cudaMemcpyAsync(h_D, dD, (size_t) count * n * 8, cudaMemcpyDeviceToHost, st));
    BB_CUDA(cudaMemcpyAsync(h_info, dinfo, (size_t) count * 4, cudaMemcpyDeviceToHost, st));
    BB_CUDA(cudaStreamSynchronize(st));
    cudaFree(dA); cudaFree(dB); cudaFree(dD); cudaFree(dinfo);
    return BBGPU_OK;
}

} // extern "C"
