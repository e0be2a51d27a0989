// Differential-evolution device engine: JADE (de/jade.cpp:98-206) and L-SHADE
// (de/shade.cpp:96-236) generations behind the C ABI.
//
// The reference updates the swarm IN PLACE inside the candidate loop (asynchronous DE):
// candidate i reads x_best, x_r1, x_r2 that candidates j < i may already have replaced, the
// archive grows / is overwritten inside the loop, and its size feeds the range of later r2
// draws.  A one-shot synchronous kernel is not equivalent.  The engine keeps the exact
// sequential semantics and resolves them in parallel as a fixed point:
//
//   round r: every candidate (one warp each) derives the *signature* of what the sequential
//            loop would have shown it -- for each parent j < i either "old row" or "trial of j,
//            version v", the archive writer of its slot, its r2 -- from the previous round's
//            success flags.  If the signature changed it rebuilds its trial (mutation, binomial
//            crossover, midpoint bound repair), evaluates it with the registry objective in the
//            same warp, and publishes (success, version).  Candidate 0 is final after round 1,
//            candidate i once its predecessors are; a round with no rebuild is the fixed point,
//            which equals the sequential result by induction on i.
//
// Trials are double buffered per candidate so a reader never sees a row being rewritten.
// After the fixed point: archive writes (last writer per slot wins, old rows), parameter
// adaptation, greedy replacement, and for L-SHADE the end-of-generation sort, population
// reduction and archive trimming.  Memory-bound: 5 row passes of n doubles per evaluation.
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdlib>

#include <cooperative_groups.h>

#include "objectives.cuh"
#include "philox.cuh"
#include "rank.cuh"

namespace bb {

constexpr int DE_MAX_ROUNDS = 4096;
enum { TAG_CR = 1, TAG_F = 2, TAG_IBEST = 3, TAG_R1 = 4, TAG_R2 = 5, TAG_JRAND = 6, TAG_SLOT = 7, TAG_CROSS = 8,
       TAG_RI = 9, TAG_PI = 10, TAG_TRIM = 11, TAG_INIT = 12 };

struct DeDev {
    int kind, n, np, npA /* archive capacity == current np */, h, archive, repaircr, replay, obj_id;
    int64_t ld;
    uint64_t seed, gen;
    double pelite;
    const double *aux, *lower, *upper;
    double *SW, *F, *ARCH, *TR[2], *FT, *CR, *Fm, *CR1, *MCR, *MF, *mu /* [mucr, muf] */;
    const double *zcr, *cauchy, *U;
    int *ibest, *r1, *r2, *jrand, *slot, *ri, *arch_row, *larch /* device scalar */;
    int *succ[2], *ver[2], *trbuf[2], *sig, *nrecomp, *kslot;
};

__device__ __forceinline__ uint32_t de_u32(const DeDev &d, int cand, int slot, int tag) {
    const Philox4 r = philox4x32_10((uint32_t) slot, (uint32_t) cand, (uint32_t) d.gen, ((uint32_t) (d.gen >> 32)) ^ ((uint32_t) tag << 24),
            (uint32_t) d.seed, (uint32_t) (d.seed >> 32));
    return r.x;
}
__device__ __forceinline__ double de_u53(const DeDev &d, int cand, int slot, int tag) {
    const Philox4 r = philox4x32_10((uint32_t) slot, (uint32_t) cand, (uint32_t) d.gen, ((uint32_t) (d.gen >> 32)) ^ ((uint32_t) tag << 24),
            (uint32_t) d.seed, (uint32_t) (d.seed >> 32));
    return u53(r.x, r.y);
}
__device__ __forceinline__ int de_int(const DeDev &d, int cand, int slot, int tag, int range) {
    return (int) __umulhi(de_u32(d, cand, slot, tag), (uint32_t) range);
}

// per-candidate control parameters for the generation (de/jade.cpp:107-126, de/shade.cpp:104-126)
__device__ __forceinline__ void de_params_one(const DeDev &d, int i) {
    int ri = 0;
    if (d.kind == BBGPU_DE_SHADE) ri = d.replay ? d.ri[i] : de_int(d, i, 0, TAG_RI, d.h);
    const double mcr = d.kind == BBGPU_DE_JADE ? d.mu[0] : d.MCR[ri];
    const double mf = d.kind == BBGPU_DE_JADE ? d.mu[1] : d.MF[ri];
    double z, c, F;
    if (d.replay) {
        z = d.zcr[i];
        c = d.cauchy[i];
        F = fmin(__dadd_rn(__dmul_rn(c, 0.1), mf), 1.0);
    } else {
        const Philox4 r = philox4x32_10(0u, (uint32_t) i, (uint32_t) d.gen, ((uint32_t) (d.gen >> 32)) ^ ((uint32_t) TAG_CR << 24),
                (uint32_t) d.seed, (uint32_t) (d.seed >> 32));
        double z1;
        normal_pair(r, z, z1);
        int t = 0;
        do {
            c = tan(3.141592653589793 * (de_u53(d, i, t++, TAG_F) - 0.5));
            F = fmin(__dadd_rn(__dmul_rn(c, 0.1), mf), 1.0);
        } while (d.kind == BBGPU_DE_JADE ? (F < 0.0) : (F <= 0.0));
        int nelite;
        if (d.kind == BBGPU_DE_JADE) {
            nelite = max(1, (int) (d.pelite * d.np));
        } else {
            const double plo = fmin(2.0 / d.n, 0.2);
            const double pi = de_u53(d, i, 0, TAG_PI) * (0.2 - plo) + plo;
            nelite = max(1, (int) (pi * d.np));
        }
        d.ibest[i] = de_int(d, i, 0, TAG_IBEST, nelite);
        int r1 = i, tt = 0;
        while (r1 == i && d.np > 1) r1 = de_int(d, i, tt++, TAG_R1, d.np);
        d.r1[i] = r1;
        d.jrand[i] = de_int(d, i, 0, TAG_JRAND, d.n);
        d.slot[i] = de_int(d, i, 0, TAG_SLOT, d.np);
    }
    d.CR[i] = fmax(0.0, fmin(__dadd_rn(__dmul_rn(z, 0.1), mcr), 1.0));
    d.Fm[i] = F;
    d.sig[4 * i + 0] = INT_MIN; // force a build in round 0
    d.succ[0][i] = 0;
    d.ver[0][i] = 0;
    d.trbuf[0][i] = 0;
}

__global__ void __launch_bounds__(256) de_params_kernel(DeDev d) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < d.np) de_params_one(d, i);
}

// CTA-wide (256 threads): push flags -> exclusive prefix counts `pre` and archive targets `tgt` (both in
// smem, np ints each); wsum[0..8] scratch, wsum[8] = total pushes
__device__ __forceinline__ void de_scan_targets(const DeDev &d, const int *__restrict__ succ, int larch0, int *pre, int *tgt,
        int *wsum) {
    const int np = d.np, T = blockDim.x;
    const int per = (np + T - 1) / T;
    const int b = threadIdx.x * per, e = min(np, b + per);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int local = 0;
    for (int i = b; i < e; i++) {
        const int s = succ[i];
        const int push = d.archive && (d.kind == BBGPU_DE_JADE ? (s >= 1) : (s == 2));
        tgt[i] = push;
        local += push;
    }
    int incl = local;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= off) incl += v;
    }
    __syncthreads(); // previous users of wsum are done
    if (lane == 31) wsum[wid] = incl;
    __syncthreads();
    if (threadIdx.x == 0) {
        int run = 0;
        const int nw = T >> 5;
        for (int w = 0; w < nw; w++) {
            const int v = wsum[w];
            wsum[w] = run;
            run += v;
        }
        wsum[8] = run;
    }
    __syncthreads();
    int run = wsum[wid] + incl - local;
    for (int i = b; i < e; i++) {
        const int push = tgt[i];
        pre[i] = run;
        const int la = larch0 + run;
        tgt[i] = push ? (la < d.npA ? la : d.slot[i]) : -1;
        run += push;
    }
    __syncthreads();
}

// one candidate (one warp) in one resolution round
__device__ __forceinline__ void de_candidate(const DeDev &d, int i, int round, int larch0, const int *pre, const int *tgt) {
    const int cur = round & 1, nxt = cur ^ 1;
    const int lane = threadIdx.x & 31;
    const int np = d.np, n = d.n;
    const int *succ = d.succ[cur], *ver = d.ver[cur], *trb = d.trbuf[cur];
    const int larch_i = d.archive ? min(d.npA, larch0 + pre[i]) : 0;
    const int jb = d.ibest[i], j1 = d.r1[i];
    int j2;
    if (d.replay) {
        j2 = d.r2[i];
    } else {
        j2 = i;
        int t = 0;
        while ((j2 == i || j2 == j1) && (np + larch_i > 2)) j2 = de_int(d, i, t++, TAG_R2, np + larch_i);
    }
    int s0 = (jb < i && succ[jb]) ? ver[jb] : 0;
    int s1 = (j1 < i && succ[j1]) ? ver[j1] : 0;
    int s2, writer = -1;
    if (j2 >= np) {
        const int a = j2 - np;
        for (int j = lane; j < i; j += 32)
            if (tgt[j] == a) writer = j;
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) writer = max(writer, __shfl_xor_sync(0xffffffffu, writer, off));
        s2 = -2 - writer;
    } else {
        s2 = (j2 < i && succ[j2]) ? ver[j2] : 0;
    }
    int *sg = d.sig + 4 * i;
    const bool same = sg[0] == s0 && sg[1] == s1 && sg[2] == s2 && sg[3] == j2;
    __syncwarp();
    if (same) {
        if (lane == 0) {
            d.succ[nxt][i] = succ[i];
            d.ver[nxt][i] = ver[i];
            d.trbuf[nxt][i] = trb[i];
        }
        return;
    }
    const int64_t ld = d.ld;
    const double *xi = d.SW + (int64_t) i * ld;
    const double *xb = (jb < i && succ[jb]) ? d.TR[trb[jb]] + (int64_t) jb * ld : d.SW + (int64_t) jb * ld;
    const double *x1 = (j1 < i && succ[j1]) ? d.TR[trb[j1]] + (int64_t) j1 * ld : d.SW + (int64_t) j1 * ld;
    const double *x2;
    if (j2 >= np) x2 = writer >= 0 ? d.SW + (int64_t) writer * ld : d.ARCH + (int64_t) d.arch_row[j2 - np] * ld;
    else x2 = (j2 < i && succ[j2]) ? d.TR[trb[j2]] + (int64_t) j2 * ld : d.SW + (int64_t) j2 * ld;
    const int mybuf = 1 - trb[i];
    double *out = d.TR[mybuf] + (int64_t) i * ld;
    const double F = d.Fm[i], CR = d.CR[i];
    const int jr = d.jrand[i];
    int cnt = 0;
    for (int j = lane; j < n; j += 32) {
        bool cross = (j == jr);
        if (!cross) {
            double u;
            if (d.replay) {
                u = d.U[(int64_t) i * n + j];
            } else {
                const Philox4 r = philox4x32_10((uint32_t) (j >> 2), (uint32_t) i, (uint32_t) d.gen,
                        ((uint32_t) (d.gen >> 32)) ^ ((uint32_t) TAG_CROSS << 24), (uint32_t) d.seed, (uint32_t) (d.seed >> 32));
                const uint32_t w = (j & 3) == 0 ? r.x : (j & 3) == 1 ? r.y : (j & 3) == 2 ? r.z : r.w;
                u = (double) w * (1.0 / 4294967296.0);
            }
            cross = u < CR;
        }
        const double x = xi[j];
        double v = x;
        if (cross) {
            // x + F*(best - x) + F*(xr1 - xr2), left to right, no contraction (de/jade.cpp:253)
            v = __dadd_rn(__dadd_rn(x, __dmul_rn(F, __dsub_rn(xb[j], x))), __dmul_rn(F, __dsub_rn(x1[j], x2[j])));
            cnt++;
        }
        const double lo = d.lower[j], hi = d.upper[j];
        if (v < lo) v = __dmul_rn(__dadd_rn(lo, x), 0.5);       // midpoint repair (de/jade.cpp:150-156)
        else if (v > hi) v = __dmul_rn(__dadd_rn(hi, x), 0.5);
        out[j] = v;
    }
    __syncwarp();
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, off);
    const double ft = warp_objective(d.obj_id, n, (const double*) out, d.aux);
    if (lane == 0) {
        const double fi = d.F[i];
        const int s = ft <= fi ? (ft < fi ? 2 : 1) : 0;
        d.FT[i] = ft;
        d.CR1[i] = d.repaircr ? __ddiv_rn((double) cnt, (double) n) : CR;
        d.succ[nxt][i] = s;
        d.trbuf[nxt][i] = mybuf;
        d.ver[nxt][i] = ver[i] + ((s != succ[i] || s > 0) ? 1 : 0);
        sg[0] = s0; sg[1] = s1; sg[2] = s2; sg[3] = j2;
        d.r2[i] = j2;
        atomicAdd(d.nrecomp + round, 1);
    }
}

// one resolution round; 8 candidates (warps) per CTA
__global__ void __launch_bounds__(256) de_round_kernel(DeDev d, int round) {
    extern __shared__ int sm[];
    int *pre = sm, *tgt = sm + d.np, *wsum = sm + 2 * d.np;
    const int larch0 = *d.larch;
    de_scan_targets(d, d.succ[round & 1], larch0, pre, tgt, wsum);
    const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (i < d.np) de_candidate(d, i, round, larch0, pre, tgt);
}

// after the fixed point: archive rows (last writer per slot wins; the value is the OLD parent row)
__device__ __forceinline__ void de_archive_one(const DeDev &d, int j, const int *tgt) {
    const int lane = threadIdx.x & 31;
    const int a = tgt[j];
    if (a < 0) return;
    int later = 0;
    for (int q = j + 1 + lane; q < d.np; q += 32) later |= (tgt[q] == a);
    later = __any_sync(0xffffffffu, later);
    if (later) return;
    const double *src = d.SW + (int64_t) j * d.ld;
    double *dst = d.ARCH + (int64_t) d.arch_row[a] * d.ld;
    for (int c = lane; c < d.n; c += 32) dst[c] = src[c];
}

__global__ void __launch_bounds__(256) de_archive_kernel(DeDev d, int fin, int *larch_new) {
    extern __shared__ int sm[];
    int *pre = sm, *tgt = sm + d.np, *wsum = sm + 2 * d.np;
    const int larch0 = *d.larch;
    de_scan_targets(d, d.succ[fin], larch0, pre, tgt, wsum);
    if (blockIdx.x == 0 && threadIdx.x == 0) *larch_new = d.archive ? min(d.npA, larch0 + wsum[8]) : 0;
    const int j = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (j < d.np) de_archive_one(d, j, tgt);
}

// parameter adaptation (de/jade.cpp:185-205, de/shade.cpp:186-212); reads the OLD parent fitness
__device__ __forceinline__ void de_adapt_block(const DeDev &d, int fin, double cdamp, double sigma_cr, int *nsucc_out,
        double *red) {
    const int *succ = d.succ[fin];
    const bool jade = d.kind == BBGPU_DE_JADE;
    double cnt = 0, s_cr = 0, s_cr2 = 0, s_f = 0, s_f2 = 0, s_w = 0, s_wcr = 0, s_wf = 0, s_wf2 = 0, acc = 0;
    for (int i = threadIdx.x; i < d.np; i += blockDim.x) {
        const int s = succ[i];
        if (s >= 1) acc += 1.0;
        if (jade ? (s >= 1) : (s == 2)) {
            const double cr = d.CR1[i], f = d.Fm[i];
            cnt += 1.0;
            s_cr += cr; s_cr2 += cr * cr; s_f += f; s_f2 += f * f;
            if (!jade) {
                const double w = d.F[i] - d.FT[i];
                s_w += w; s_wcr += w * cr; s_wf += w * f; s_wf2 += w * f * f;
            }
        }
    }
    cnt = block_sum(cnt, red); acc = block_sum(acc, red);
    s_cr = block_sum(s_cr, red); s_cr2 = block_sum(s_cr2, red); s_f = block_sum(s_f, red); s_f2 = block_sum(s_f2, red);
    if (!jade) {
        s_w = block_sum(s_w, red); s_wcr = block_sum(s_wcr, red); s_wf = block_sum(s_wf, red); s_wf2 = block_sum(s_wf2, red);
    }
    double dev2 = 0.0;
    if (jade && cnt > 0) {
        const double mean = s_cr / cnt;
        for (int i = threadIdx.x; i < d.np; i += blockDim.x)
            if (succ[i] >= 1) { const double t = d.CR1[i] - mean; dev2 += t * t; }
        dev2 = block_sum(dev2, red);
    }
    if (threadIdx.x == 0) {
        *nsucc_out = (int) acc;
        if (jade) {
            double meancr = 0.0, meanf = 0.0;
            if (cnt > 0) {
                const double sd = sqrt(dev2 / cnt);
                meancr = sd > sigma_cr ? sqrt(s_cr2 / cnt) : s_cr / cnt;
                meanf = (s_f2 / cnt) / (s_f / cnt);
            }
            d.mu[0] = (1.0 - cdamp) * d.mu[0] + cdamp * meancr;
            d.mu[1] = (1.0 - cdamp) * d.mu[1] + cdamp * meanf;
        } else if (cnt > 0) {
            int k = *d.kslot;
            d.MCR[k - 1] = s_wcr / s_w;
            d.MF[k - 1] = s_wf2 / s_wf;
            k++;
            if (k > d.h) k = 1;
            *d.kslot = k;
        }
    }
}

__global__ void __launch_bounds__(1024) de_adapt_kernel(DeDev d, int fin, double cdamp, double sigma_cr, int *nsucc_out) {
    __shared__ double red[32];
    de_adapt_block(d, fin, cdamp, sigma_cr, nsucc_out, red);
}

// greedy replacement (de/jade.cpp:174-176)
__device__ __forceinline__ void de_replace_one(const DeDev &d, int i, int fin) {
    if (d.succ[fin][i] == 0) return;
    const int lane = threadIdx.x & 31;
    const double *src = d.TR[d.trbuf[fin][i]] + (int64_t) i * d.ld;
    double *dst = d.SW + (int64_t) i * d.ld;
    for (int c = lane; c < d.n; c += 32) dst[c] = src[c];
    if (lane == 0) d.F[i] = d.FT[i];
}

__global__ void __launch_bounds__(256) de_replace_kernel(DeDev d, int fin) {
    const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (i < d.np) de_replace_one(d, i, fin);
}

__global__ void __launch_bounds__(256) de_gather_rows_kernel(const double *__restrict__ src, double *__restrict__ dst,
        const int *__restrict__ perm, int rows, int n, int64_t ld) {
    const int r = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (r >= rows) return;
    const double *s = src + (int64_t) perm[r] * ld;
    double *t = dst + (int64_t) r * ld;
    for (int c = threadIdx.x & 31; c < n; c += 32) t[c] = s[c];
}

// archive trimming by random erase (de/shade.cpp:228-235) on the row-index indirection
__device__ __forceinline__ void de_trim_thread(const DeDev &d, int *larch, int npnew, const int *trim, int ntrim) {
    int la = *larch, t = 0;
    while (la > npnew) {
        int r;
        if (d.replay) r = t < ntrim ? trim[t] : 0;
        else r = de_int(d, 0, t, TAG_TRIM, la);
        t++;
        const int saved = d.arch_row[r];
        for (int q = r; q + 1 < la; q++) d.arch_row[q] = d.arch_row[q + 1];
        d.arch_row[la - 1] = saved;
        la--;
    }
    *larch = la;
}

__global__ void de_trim_kernel(DeDev d, int *larch, int npnew, const int *trim, int ntrim) {
    if (threadIdx.x == 0 && blockIdx.x == 0) de_trim_thread(d, larch, npnew, trim, ntrim);
}

// ||x_i||_2 per member, then Welford-equivalent M2 of the radii and the best fitness
__global__ void __launch_bounds__(256) de_norms_kernel(const double *__restrict__ SW, int np, int n, int64_t ld,
        double *__restrict__ radius) {
    const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (i >= np) return;
    const double *x = SW + (int64_t) i * ld;
    double s = 0.0;
    for (int c = threadIdx.x & 31; c < n; c += 32) s = fma(x[c], x[c], s);
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) radius[i] = sqrt(s);
}

__device__ __forceinline__ void de_report_block(const DeDev &d, const double *__restrict__ radius, int np, const int *larch,
        const int *nsucc, bbgpu_de_report *rep, double *red) {
    double s = 0.0, fb = INFINITY;
    for (int i = threadIdx.x; i < np; i += blockDim.x) {
        s += radius[i];
        fb = fmin(fb, d.F[i]);
    }
    s = block_sum(s, red);
    fb = -block_max(-fb, red);
    const double mean = s / np;
    double m2 = 0.0;
    for (int i = threadIdx.x; i < np; i += blockDim.x) {
        const double t = radius[i] - mean;
        m2 += t * t;
    }
    m2 = block_sum(m2, red);
    if (threadIdx.x == 0) {
        rep->np = np;
        rep->larch = *larch;
        rep->nsucc = *nsucc;
        rep->kslot = *d.kslot;
        rep->mucr = d.mu[0];
        rep->muf = d.mu[1];
        rep->fbest = fb;
        rep->radius_m2 = m2;
    }
}

__global__ void __launch_bounds__(1024) de_report_kernel(DeDev d, const double *__restrict__ radius, int np, const int *larch,
        const int *nsucc, bbgpu_de_report *rep) {
    __shared__ double red[32];
    de_report_block(d, radius, np, larch, nsucc, rep, red);
}


// ------------------------------------------------------------------------------------------
// The whole generation as ONE cooperative kernel: sort -> parameters -> resolution rounds (grid.sync
// between rounds, loop until a round rebuilds nothing) -> archive -> adaptation -> replacement ->
// (L-SHADE: sort, population reduction, archive trimming) -> radii -> report.  No host round trip
// inside a generation, and none between generations unless the caller asks for the report.
// ------------------------------------------------------------------------------------------
struct DeCoop {
    double *SWalt, *Falt, *radius;
    int *nsucc, *rounds_out;
    bbgpu_de_report *rep;
    double cdamp, sigma_cr;
    const int *trim;
    int ntrim, npnew;
};

// counting rank + row scatter: (SW, F) -> (SWalt, Falt) sorted ascending by (fitness, index)
__device__ __forceinline__ void de_sort_rows(const double *__restrict__ SW, const double *__restrict__ F,
        double *__restrict__ SWalt, double *__restrict__ Falt, int np, int n, int64_t ld, int gw, int W) {
    const int lane = threadIdx.x & 31;
    for (int i = gw; i < np; i += W) {
        const double fi = F[i];
        const unsigned long long ki = sortable_key(fi);
        int cnt = 0;
        for (int j = lane; j < np; j += 32) {
            const unsigned long long kj = sortable_key(F[j]);
            cnt += (kj < ki) || (kj == ki && j < i);
        }
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, off);
        const double *src = SW + (int64_t) i * ld;
        double *dst = SWalt + (int64_t) cnt * ld;
        for (int c = lane; c < n; c += 32) dst[c] = src[c];
        if (lane == 0) Falt[cnt] = fi;
    }
}

__global__ void __launch_bounds__(256) de_generation_kernel(DeDev d, DeCoop c) {
    namespace cg = cooperative_groups;
    cg::grid_group grid = cg::this_grid();
    extern __shared__ int sm[];
    int *pre = sm, *tgt = sm + d.np, *wsum = sm + 2 * d.np;
    __shared__ double red[32];
    const int lane = threadIdx.x & 31;
    const int W = gridDim.x * 8, gw = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int GT = gridDim.x * 256, gtid = blockIdx.x * 256 + threadIdx.x;
    const int np = d.np, n = d.n;

    if (d.kind == BBGPU_DE_JADE) { // de/jade.cpp:101
        de_sort_rows(d.SW, d.F, c.SWalt, c.Falt, np, n, d.ld, gw, W);
        grid.sync();
        double *t = d.SW; d.SW = c.SWalt; c.SWalt = t;
        t = d.F; d.F = c.Falt; c.Falt = t;
    }
    for (int i = gtid; i < np; i += GT) de_params_one(d, i);
    grid.sync();

    const int larch0 = *d.larch;
    int round = 0;
    while (true) {
        de_scan_targets(d, d.succ[round & 1], larch0, pre, tgt, wsum);
        for (int i = gw; i < np; i += W) de_candidate(d, i, round, larch0, pre, tgt);
        grid.sync();
        const int rebuilt = *((volatile int*) (d.nrecomp + round));
        round++;
        if (rebuilt == 0 || round >= DE_MAX_ROUNDS) break;
    }
    const int fin = round & 1;

    de_scan_targets(d, d.succ[fin], larch0, pre, tgt, wsum);
    for (int j = gw; j < np; j += W) de_archive_one(d, j, tgt);
    if (blockIdx.x == 0) de_adapt_block(d, fin, c.cdamp, c.sigma_cr, c.nsucc, red);
    grid.sync(); // archive rows and adaptation read the OLD parents; replacement comes after

    for (int i = gw; i < np; i += W) de_replace_one(d, i, fin);
    if (gtid == 0) {
        *d.larch = d.archive ? min(d.npA, larch0 + wsum[8]) : 0;
        for (int r = 0; r < round; r++) d.nrecomp[r] = 0;
        *c.rounds_out = round;
    }
    grid.sync();

    int np_after = np;
    if (d.kind == BBGPU_DE_SHADE) { // de/shade.cpp:215-235
        de_sort_rows(d.SW, d.F, c.SWalt, c.Falt, np, n, d.ld, gw, W);
        grid.sync();
        double *t = d.SW; d.SW = c.SWalt; c.SWalt = t;
        t = d.F; d.F = c.Falt; c.Falt = t;
        np_after = min(np, c.npnew);
        if (gtid == 0 && d.archive) de_trim_thread(d, d.larch, c.npnew, c.trim, c.ntrim);
    }
    for (int i = gw; i < np_after; i += W) {
        const double *x = d.SW + (int64_t) i * d.ld;
        double s = 0.0;
        for (int q = lane; q < n; q += 32) s = fma(x[q], x[q], s);
        s = warp_sum(s);
        if (lane == 0) c.radius[i] = sqrt(s);
    }
    grid.sync();
    if (blockIdx.x == 0) {
        de_report_block(d, c.radius, np_after, d.larch, c.nsucc, c.rep, red);
        if (threadIdx.x == 0) c.rep->rounds = round;
    }
}

__global__ void __launch_bounds__(256) de_uniform_init_kernel(DeDev d) {
    const int64_t gid = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (int64_t) d.np * d.n) return;
    const int i = (int) (gid / d.n), j = (int) (gid % d.n);
    const Philox4 r = philox4x32_10((uint32_t) j, (uint32_t) i, 0u, (uint32_t) TAG_INIT << 24, (uint32_t) d.seed, (uint32_t) (d.seed >> 32));
    const double u = u53(r.x, r.y);
    d.SW[(int64_t) i * d.ld + j] = u * (d.upper[j] - d.lower[j]) + d.lower[j];
}

__global__ void de_iota_kernel(int *v, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[i] = i;
}

} // namespace bb

using namespace bb;

struct bbgpu_de {
    bbgpu_ctx *ctx = nullptr;
    bbgpu_de_config cfg;
    int n = 0, npcap = 0, np = 0, fev = 0;
    int64_t ld = 0;
    uint64_t gen = 0;
    int last_rounds = 6;
    DeDev d;
    double *SWalt = nullptr, *Falt = nullptr, *radius = nullptr, *Ubuf = nullptr, *zbuf = nullptr, *cbuf = nullptr;
    int *perm = nullptr, *ribuf = nullptr, *trimbuf = nullptr, *larch_new = nullptr, *nsucc = nullptr, *rounds_dev = nullptr;
    int coop_blocks = 0;
    bool use_coop = true;
    int trim_cap = 0;
    int *h_nrecomp = nullptr; // pinned
    bbgpu_de_report *d_rep = nullptr, *h_rep = nullptr;
    std::vector<void*> allocs;
    int fin = 0;
};

namespace {

template<typename T> int de_alloc(bbgpu_de *h, T **p, size_t count, bool zero = true) {
    BB_TRY(dev_alloc(p, count));
    h->allocs.push_back(*p);
    if (zero) BB_CUDA(cudaMemsetAsync(*p, 0, count * sizeof(T), h->ctx->stream));
    return BBGPU_OK;
}

int de_sort(bbgpu_de *h) {
    bbgpu_ctx *ctx = h->ctx;
    BB_TRY(rank_fitness(ctx, h->d.F, h->np, h->perm, h->Falt));
    de_gather_rows_kernel<<<ceil_div(h->np, 8), 256, 0, ctx->stream>>>(h->d.SW, h->SWalt, h->perm, h->np, h->n, h->ld);
    BB_TRY(launched(ctx, "de_gather_rows_kernel"));
    std::swap(h->d.SW, h->SWalt);
    std::swap(h->d.F, h->Falt);
    return BBGPU_OK;
}

} // namespace

extern "C" {

int bbgpu_de_create(bbgpu_ctx *ctx, const bbgpu_de_config *cfg, const double *h_lower, const double *h_upper, bbgpu_de **out) {
    BB_ARG(ctx && cfg && h_lower && h_upper && out);
    BB_ARG(cfg->n >= 1 && cfg->np >= 4 && cfg->np <= 16384);
    BB_ARG(cfg->kind == BBGPU_DE_JADE || cfg->kind == BBGPU_DE_SHADE);
    BB_ARG(cfg->kind == BBGPU_DE_JADE || (cfg->h >= 1 && cfg->npmin >= 4 && cfg->mfev > 0));
    BB_CUDA(cudaSetDevice(ctx->device));
    bbgpu_de *h = new bbgpu_de();
    h->ctx = ctx;
    h->cfg = *cfg;
    h->n = cfg->n;
    h->npcap = h->np = cfg->np;
    h->ld = round_up(cfg->n, 8);
    const size_t rows = (size_t) h->npcap, ld = (size_t) h->ld;
    const int hh = std::max(1, cfg->h);
    DeDev &d = h->d;
    memset(&d, 0, sizeof(d));
    d.kind = cfg->kind; d.n = cfg->n; d.np = h->np; d.npA = h->np; d.h = hh; d.archive = cfg->archive; d.repaircr = cfg->repaircr;
    d.ld = h->ld; d.seed = cfg->seed; d.pelite = cfg->pelite;
    int rc = BBGPU_OK;
    double *lo = nullptr, *hi = nullptr;
#define DA(ptr, count) if (rc == BBGPU_OK) rc = de_alloc(h, &(ptr), (size_t) (count))
    DA(lo, ld); DA(hi, ld);
    DA(d.SW, rows * ld); DA(h->SWalt, rows * ld); DA(d.F, rows); DA(h->Falt, rows); DA(d.ARCH, rows * ld);
    DA(d.TR[0], rows * ld); DA(d.TR[1], rows * ld); DA(d.FT, rows); DA(d.CR, rows); DA(d.Fm, rows); DA(d.CR1, rows);
    DA(d.MCR, hh); DA(d.MF, hh); DA(d.mu, 2); DA(h->radius, rows);
    DA(d.ibest, rows); DA(d.r1, rows); DA(d.r2, rows); DA(d.jrand, rows); DA(d.slot, rows); DA(d.ri, rows); DA(d.arch_row, rows);
    DA(d.larch, 1); DA(h->larch_new, 1); DA(h->nsucc, 1); DA(d.kslot, 1); DA(h->rounds_dev, 1);
    DA(d.succ[0], rows); DA(d.succ[1], rows); DA(d.ver[0], rows); DA(d.ver[1], rows); DA(d.trbuf[0], rows); DA(d.trbuf[1], rows);
    DA(d.sig, rows * 4); DA(d.nrecomp, DE_MAX_ROUNDS); DA(h->perm, rows);
    DA(h->zbuf, rows); DA(h->cbuf, rows); DA(h->ribuf, rows);
    DA(h->d_rep, 1);
#undef DA
    if (rc == BBGPU_OK && cudaMallocHost((void**) &h->h_rep, sizeof(bbgpu_de_report)) != cudaSuccess) rc = fail(BBGPU_ERR_CUDA, "pinned alloc failed");
    if (rc == BBGPU_OK && cudaMallocHost((void**) &h->h_nrecomp, DE_MAX_ROUNDS * sizeof(int)) != cudaSuccess) rc = fail(BBGPU_ERR_CUDA, "pinned alloc failed");
    if (rc != BBGPU_OK) { bbgpu_de_destroy(h); return rc; }
    d.lower = lo; d.upper = hi;
    cudaStream_t st = ctx->stream;
    cudaMemcpyAsync(lo, h_lower, (size_t) h->n * 8, cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(hi, h_upper, (size_t) h->n * 8, cudaMemcpyHostToDevice, st);
    std::vector<double> half(std::max(hh, 2), 0.5);
    cudaMemcpyAsync(d.MCR, half.data(), (size_t) hh * 8, cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(d.MF, half.data(), (size_t) hh * 8, cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(d.mu, half.data(), 16, cudaMemcpyHostToDevice, st);
    const int one = 1;
    cudaMemcpyAsync(d.kslot, &one, 4, cudaMemcpyHostToDevice, st);
    de_iota_kernel<<<ceil_div(h->npcap, 256), 256, 0, st>>>(d.arch_row, h->npcap);
    rc = launched(ctx, "de_iota_kernel");
    {
        // cooperative launch geometry: all CTAs must be co-resident
        const char *e = std::getenv("BBGPU_DE_COOP");
        h->use_coop = e ? std::atoi(e) != 0 : true;
        const size_t smem = ((size_t) 2 * h->npcap + 16) * sizeof(int);
        int supported = 0, per_sm = 0;
        cudaDeviceGetAttribute(&supported, cudaDevAttrCooperativeLaunch, ctx->device);
        if (smem > 48 * 1024) cudaFuncSetAttribute(de_generation_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, de_generation_kernel, 256, smem);
        h->coop_blocks = std::min(ceil_div(h->npcap, 8), per_sm * ctx->sm_count);
        if (!supported || h->coop_blocks < 1) h->use_coop = false;
    }
    if (cudaStreamSynchronize(st) != cudaSuccess && rc == BBGPU_OK) rc = fail(BBGPU_ERR_CUDA, "de create sync failed");
    if (rc != BBGPU_OK) { bbgpu_de_destroy(h); return rc; }
    *out = h;
    return BBGPU_OK;
}

int bbgpu_de_destroy(bbgpu_de *h) {
    if (!h) return BBGPU_OK;
    cudaStreamSynchronize(h->ctx->stream);
    for (void *p : h->allocs) cudaFree(p);
    if (h->Ubuf) cudaFree(h->Ubuf);
    if (h->trimbuf) cudaFree(h->trimbuf);
    if (h->h_rep) cudaFreeHost(h->h_rep);
    if (h->h_nrecomp) cudaFreeHost(h->h_nrecomp);
    delete h;
    return BBGPU_OK;
}

int bbgpu_de_init(bbgpu_de *h, bbgpu_objective *obj, const double *h_x) {
    BB_ARG(h && obj && obj->n == h->n);
    bbgpu_ctx *ctx = h->ctx;
    cudaStream_t st = ctx->stream;
    DeDev &d = h->d;
    h->np = h->npcap;
    d.np = d.npA = h->np;
    h->fev = 0;
    h->gen = 0;
    const int hh = d.h;
    std::vector<double> half(std::max(hh, 2), 0.5);
    BB_CUDA(cudaMemcpyAsync(d.MCR, half.data(), (size_t) hh * 8, cudaMemcpyHostToDevice, st));
    BB_CUDA(cudaMemcpyAsync(d.MF, half.data(), (size_t) hh * 8, cudaMemcpyHostToDevice, st));
    BB_CUDA(cudaMemcpyAsync(d.mu, half.data(), 16, cudaMemcpyHostToDevice, st));
    const int one = 1;
    BB_CUDA(cudaMemcpyAsync(d.kslot, &one, 4, cudaMemcpyHostToDevice, st));
    BB_CUDA(cudaMemsetAsync(d.larch, 0, 4, st));
    BB_CUDA(cudaStreamSynchronize(st)); // `half` / `one` are stack temporaries
    if (h_x) {
        BB_CUDA(cudaMemcpy2DAsync(d.SW, (size_t) h->ld * 8, h_x, (size_t) h->n * 8, (size_t) h->n * 8, h->np, cudaMemcpyHostToDevice, st));
    } else {
        de_uniform_init_kernel<<<ceil_div((int64_t) h->np * h->n, 256), 256, 0, st>>>(d);
        BB_TRY(launched(ctx, "de_uniform_init_kernel"));
    }
    BB_TRY(evaluate(ctx, obj, d.SW, h->ld, h->np, d.F));
    h->fev += h->np;
    if (h->cfg.kind == BBGPU_DE_SHADE) BB_TRY(de_sort(h)); // de/shade.cpp:93
    BB_CUDA(cudaStreamSynchronize(st));
    return BBGPU_OK;
}

int bbgpu_de_generation(bbgpu_de *h, bbgpu_objective *obj, const bbgpu_de_draws *draws, bbgpu_de_report *out) {
    BB_ARG(h && obj && obj->n == h->n);
    if (obj->mode != 0) return fail(BBGPU_ERR_ARG, "the fused DE generation needs a registry objective (user objectives: next round)");
    bbgpu_ctx *ctx = h->ctx;
    cudaStream_t st = ctx->stream;
    DeDev &d = h->d;
    const int np = h->np, n = h->n;
    d.np = np; d.npA = np; d.gen = h->gen; d.obj_id = obj->id; d.aux = obj->d_aux;
    d.replay = draws != nullptr;
    if (h->cfg.kind == BBGPU_DE_JADE && !h->use_coop) BB_TRY(de_sort(h)); // de/jade.cpp:101 (the cooperative kernel sorts itself)
    if (draws) {
        BB_ARG(draws->zcr && draws->cauchy && draws->ibest && draws->r1 && draws->r2 && draws->jrand && draws->slot && draws->u);
        BB_ARG(h->cfg.kind == BBGPU_DE_JADE || draws->ri);
        if (!h->Ubuf) BB_TRY(dev_alloc(&h->Ubuf, (size_t) h->npcap * n));
        BB_CUDA(cudaMemcpyAsync(h->zbuf, draws->zcr, (size_t) np * 8, cudaMemcpyHostToDevice, st));
        BB_CUDA(cudaMemcpyAsync(h->cbuf, draws->cauchy, (size_t) np * 8, cudaMemcpyHostToDevice, st));
        BB_CUDA(cudaMemcpyAsync(d.ibest, draws->ibest, (size_t) np * 4, cudaMemcpyHostToDevice, st));
        BB_CUDA(cudaMemcpyAsync(d.r1, draws->r1, (size_t) np * 4, cudaMemcpyHostToDevice, st));
        BB_CUDA(cudaMemcpyAsync(d.r2, draws->r2, (size_t) np * 4, cudaMemcpyHostToDevice, st));
        BB_CUDA(cudaMemcpyAsync(d.jrand, draws->jrand, (size_t) np * 4, cudaMemcpyHostToDevice, st));
        BB_CUDA(cudaMemcpyAsync(d.slot, draws->slot, (size_t) np * 4, cudaMemcpyHostToDevice, st));
        BB_CUDA(cudaMemcpyAsync(h->Ubuf, draws->u, (size_t) np * n * 8, cudaMemcpyHostToDevice, st));
        if (draws->ri) BB_CUDA(cudaMemcpyAsync(d.ri, draws->ri, (size_t) np * 4, cudaMemcpyHostToDevice, st));
        if (draws->ntrim > 0) {
            BB_ARG(draws->trim);
            if (draws->ntrim > h->trim_cap) {
                if (h->trimbuf) cudaFree(h->trimbuf);
                BB_TRY(dev_alloc(&h->trimbuf, (size_t) draws->ntrim));
                h->trim_cap = draws->ntrim;
            }
            BB_CUDA(cudaMemcpyAsync(h->trimbuf, draws->trim, (size_t) draws->ntrim * 4, cudaMemcpyHostToDevice, st));
        }
        d.zcr = h->zbuf; d.cauchy = h->cbuf; d.U = h->Ubuf;
    }
    if (h->use_coop) {
        // ---- single cooperative launch per generation
        DeCoop c;
        c.SWalt = h->SWalt; c.Falt = h->Falt; c.radius = h->radius; c.nsucc = h->nsucc; c.rounds_out = h->rounds_dev;
        c.rep = h->d_rep; c.cdamp = h->cfg.cdamp; c.sigma_cr = h->cfg.sigma;
        c.trim = h->trimbuf; c.ntrim = draws ? draws->ntrim : 0;
        const int fev_after = h->fev + np;
        c.npnew = h->cfg.kind == BBGPU_DE_SHADE
                ? (int) std::round((h->cfg.npmin - h->npcap) * ((1.0 * fev_after) / h->cfg.mfev) + h->npcap) : np;
        const size_t smem = ((size_t) 2 * np + 16) * sizeof(int);
        void *args[] = { (void*) &d, (void*) &c };
        {
            ProfScope prof(ctx, BBGPU_PROF_DE_ROUND);
            BB_CUDA(cudaLaunchCooperativeKernel((const void*) de_generation_kernel, dim3(h->coop_blocks), dim3(256), args, smem, st));
            BB_TRY(launched(ctx, "de_generation_kernel"));
        }
        // the kernel sorted once (JADE: at the start, L-SHADE: at the end): the live buffers swapped
        std::swap(h->d.SW, h->SWalt);
        std::swap(h->d.F, h->Falt);
        h->fev = fev_after;
        h->gen++;
        if (h->cfg.kind == BBGPU_DE_SHADE && c.npnew < h->np) h->np = c.npnew;
        h->fin = -1; // resolved from rounds_dev on demand
        if (out) {
            BB_CUDA(cudaMemcpyAsync(h->h_rep, h->d_rep, sizeof(bbgpu_de_report), cudaMemcpyDeviceToHost, st));
            BB_CUDA(cudaStreamSynchronize(st));
            *out = *h->h_rep;
            out->fev = h->fev;
        }
        return BBGPU_OK;
    }
    BB_CUDA(cudaMemsetAsync(d.nrecomp, 0, DE_MAX_ROUNDS * sizeof(int), st));
    de_params_kernel<<<ceil_div(np, 256), 256, 0, st>>>(d);
    BB_TRY(launched(ctx, "de_params_kernel"));
    const size_t smem = ((size_t) 2 * np + 257 + 8) * sizeof(int);
    if (smem > 48 * 1024) {
        BB_CUDA(cudaFuncSetAttribute(de_round_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
        BB_CUDA(cudaFuncSetAttribute(de_archive_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
    }
    // resolution rounds, launched in batches; a round that rebuilt nothing is the fixed point
    int round = 0, done_round = -1;
    int batch = std::max(2, h->last_rounds + 1);
    while (done_round < 0) {
        const int first = round;
        for (int b = 0; b < batch && round < DE_MAX_ROUNDS; b++, round++) {
            ProfScope prof(ctx, BBGPU_PROF_DE_ROUND);
            de_round_kernel<<<ceil_div(np, 8), 256, smem, st>>>(d, round);
            BB_TRY(launched(ctx, "de_round_kernel"));
        }
        BB_CUDA(cudaMemcpyAsync(h->h_nrecomp + first, d.nrecomp + first, (size_t) (round - first) * sizeof(int), cudaMemcpyDeviceToHost, st));
        BB_CUDA(cudaStreamSynchronize(st));
        for (int r = first; r < round; r++)
            if (h->h_nrecomp[r] == 0) { done_round = r; break; }
        if (done_round < 0 && round >= DE_MAX_ROUNDS) return fail(BBGPU_ERR_ARG, "DE dependency resolution did not converge");
        batch = 2;
    }
    h->last_rounds = done_round;
    const int fin = round & 1; // buffers written by the last launched round (identical past the fixed point)
    h->fin = fin;
    de_archive_kernel<<<ceil_div(np, 8), 256, smem, st>>>(d, fin, h->larch_new);
    BB_TRY(launched(ctx, "de_archive_kernel"));
    de_adapt_kernel<<<1, 1024, 0, st>>>(d, fin, h->cfg.cdamp, h->cfg.sigma, h->nsucc);
    BB_TRY(launched(ctx, "de_adapt_kernel"));
    de_replace_kernel<<<ceil_div(np, 8), 256, 0, st>>>(d, fin);
    BB_TRY(launched(ctx, "de_replace_kernel"));
    BB_CUDA(cudaMemcpyAsync(d.larch, h->larch_new, 4, cudaMemcpyDeviceToDevice, st));
    h->fev += np;
    h->gen++;
    if (h->cfg.kind == BBGPU_DE_SHADE) {
        BB_TRY(de_sort(h)); // de/shade.cpp:215
        const int npnew = (int) std::round((h->cfg.npmin - h->npcap) * ((1.0 * h->fev) / h->cfg.mfev) + h->npcap);
        if (npnew < h->np) h->np = npnew; // pop_back on the sorted swarm (de/shade.cpp:218-225)
        if (h->cfg.archive) {
            de_trim_kernel<<<1, 32, 0, st>>>(d, d.larch, npnew, h->trimbuf, draws ? draws->ntrim : 0);
            BB_TRY(launched(ctx, "de_trim_kernel"));
        }
    }
    de_norms_kernel<<<ceil_div(h->np, 8), 256, 0, st>>>(d.SW, h->np, n, h->ld, h->radius);
    BB_TRY(launched(ctx, "de_norms_kernel"));
    de_report_kernel<<<1, 1024, 0, st>>>(d, h->radius, h->np, d.larch, h->nsucc, h->d_rep);
    BB_TRY(launched(ctx, "de_report_kernel"));
    if (out) {
        BB_CUDA(cudaMemcpyAsync(h->h_rep, h->d_rep, sizeof(bbgpu_de_report), cudaMemcpyDeviceToHost, st));
        BB_CUDA(cudaStreamSynchronize(st));
        *out = *h->h_rep;
        out->fev = h->fev;
        out->rounds = done_round + 1;
    }
    return BBGPU_OK;
}

int bbgpu_de_get_swarm(bbgpu_de *h, double *h_x, double *h_f) {
    BB_ARG(h);
    cudaStream_t st = h->ctx->stream;
    if (h_x) BB_CUDA(cudaMemcpy2DAsync(h_x, (size_t) h->n * 8, h->d.SW, (size_t) h->ld * 8, (size_t) h->n * 8, h->np, cudaMemcpyDeviceToHost, st));
    if (h_f) BB_CUDA(cudaMemcpyAsync(h_f, h->d.F, (size_t) h->np * 8, cudaMemcpyDeviceToHost, st));
    BB_CUDA(cudaStreamSynchronize(st));
    return BBGPU_OK;
}

int bbgpu_de_get_archive(bbgpu_de *h, double *h_arch, int *rows) {
    BB_ARG(h && rows);
    cudaStream_t st = h->ctx->stream;
    int larch = 0;
    BB_CUDA(cudaMemcpyAsync(&larch, h->d.larch, 4, cudaMemcpyDeviceToHost, st));
    std::vector<int> rowmap((size_t) h->npcap);
    BB_CUDA(cudaMemcpyAsync(rowmap.data(), h->d.arch_row, (size_t) h->npcap * 4, cudaMemcpyDeviceToHost, st));
    BB_CUDA(cudaStreamSynchronize(st));
    *rows = larch;
    if (h_arch) {
        for (int a = 0; a < larch; a++)
            BB_CUDA(cudaMemcpyAsync(h_arch + (size_t) a * h->n, h->d.ARCH + (size_t) rowmap[a] * h->ld, (size_t) h->n * 8, cudaMemcpyDeviceToHost, st));
        BB_CUDA(cudaStreamSynchronize(st));
    }
    return BBGPU_OK;
}

int bbgpu_de_get_memory(bbgpu_de *h, double *h_mcr, double *h_mf) {
    BB_ARG(h && h_mcr && h_mf);
    cudaStream_t st = h->ctx->stream;
    BB_CUDA(cudaMemcpyAsync(h_mcr, h->d.MCR, (size_t) h->d.h * 8, cudaMemcpyDeviceToHost, st));
    BB_CUDA(cudaMemcpyAsync(h_mf, h->d.MF, (size_t) h->d.h * 8, cudaMemcpyDeviceToHost, st));
    BB_CUDA(cudaStreamSynchronize(st));
    return BBGPU_OK;
}

// per-candidate outcome of the generation just run, in the pre-sort candidate order
int bbgpu_de_get_trace(bbgpu_de *h, int32_t *h_succ, double *h_ftrial, double *h_F, double *h_CR) {
    BB_ARG(h);
    cudaStream_t st = h->ctx->stream;
    const int np = h->d.np;
    if (h->fin < 0) {
        int rounds = 0;
        BB_CUDA(cudaMemcpyAsync(&rounds, h->rounds_dev, 4, cudaMemcpyDeviceToHost, st));
        BB_CUDA(cudaStreamSynchronize(st));
        h->fin = rounds & 1;
    }
    if (h_succ) BB_CUDA(cudaMemcpyAsync(h_succ, h->d.succ[h->fin], (size_t) np * 4, cudaMemcpyDeviceToHost, st));
    if (h_ftrial) BB_CUDA(cudaMemcpyAsync(h_ftrial, h->d.FT, (size_t) np * 8, cudaMemcpyDeviceToHost, st));
    if (h_F) BB_CUDA(cudaMemcpyAsync(h_F, h->d.Fm, (size_t) np * 8, cudaMemcpyDeviceToHost, st));
    if (h_CR) BB_CUDA(cudaMemcpyAsync(h_CR, h->d.CR, (size_t) np * 8, cudaMemcpyDeviceToHost, st));
    BB_CUDA(cudaStreamSynchronize(st));
    return BBGPU_OK;
}

} // extern "C"
