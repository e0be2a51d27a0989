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
__global__ void __launch_bounds__(256) de_params_kernel(DeDev d) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d.np) return;
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

// CTA-wide: push flags -> exclusive prefix counts `pre` and archive targets `tgt` (both in smem)
__device__ __forceinline__ void de_scan_targets(const DeDev &d, const int *__restrict__ succ, int larch0, int *pre, int *tgt,
        int *chunk_tot) {
    const int np = d.np, T = blockDim.x;
    const int per = (np + T - 1) / T;
    const int b = threadIdx.x * per, e = min(np, b + per);
    int local = 0;
    for (int i = b; i < e; i++) {
        const int s = succ[i];
        const int push = d.archive && (d.kind == BBGPU_DE_JADE ? (s >= 1) : (s == 2));
        tgt[i] = push;
        local += push;
    }
    chunk_tot[threadIdx.x] = local;
    __syncthreads();
    if (threadIdx.x == 0) {
        int run = 0;
        for (int t = 0; t < T; t++) {
            const int v = chunk_tot[t];
            chunk_tot[t] = run;
            run += v;
        }
        chunk_tot[T] = run;
    }
    __syncthreads();
    int run = chunk_tot[threadIdx.x];
    for (int i = b; i < e; i++) {
        const int push = tgt[i];
        pre[i] = run;
        const int la = larch0 + run;
        tgt[i] = push ? (la < d.npA ? la : d.slot[i]) : -1;
        run += push;
    }
    __syncthreads();
}

// one resolution round; 8 candidates (warps) per CTA
__global__ void __launch_bounds__(256) de_round_kernel(DeDev d, int round) {
    extern __shared__ int sm[];
    int *pre = sm, *tgt = sm + d.np, *chunk_tot = sm + 2 * d.np;
    const int cur = round & 1, nxt = cur ^ 1;
    const int larch0 = *d.larch;
    de_scan_targets(d, d.succ[cur], larch0, pre, tgt, chunk_tot);
    const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (i >= d.np) return;
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

// after the fixed point: archive rows (last writer per slot), new archive size
__global__ void __launch_bounds__(256) de_archive_kernel(DeDev d, int fin, int *larch_new) {
    extern __shared__ int sm[];
    int *pre = sm, *tgt = sm + d.np, *chunk_tot = sm + 2 * d.np;
    const int larch0 = *d.larch;
    de_scan_targets(d, d.succ[fin], larch0, pre, tgt, chunk_tot);
    if (blockIdx.x == 0 && threadIdx.x == 0) *larch_new = d.archive ? min(d.npA, larch0 + chunk_tot[blockDim.x]) : 0;
    const int j = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (j >= d.np) return;
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

// parameter adaptation (de/jade.cpp:185-205, de/shade.cpp:186-212); reads the OLD parent fitness
__global__ void __launch_bounds__(1024) de_adapt_kernel(DeDev d, int fin, double cdamp, double sigma_cr, int *nsucc_out) {
    __shared__ double red[32];
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

// greedy replacement (de/jade.cpp:174-176)
__global__ void __launch_bounds__(256) de_replace_kernel(DeDev d, int fin) {
    const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (i >= d.np) return;
    if (d.succ[fin][i] == 0) return;
    const int lane = threadIdx.x & 31;
    const double *src = d.TR[d.trbuf[fin][i]] + (int64_t) i * d.ld;
    double *dst = d.SW + (int64_t) i * d.ld;
    for (int c = lane; c < d.n; c += 32) dst[c] = src[c];
    if (lane == 0) d.F[i] = d.FT[i];
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
__global__ void de_trim_kernel(DeDev d, int *larch, int npnew, const int *trim, int ntrim) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
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

__global__ void __launch_bounds__(1024) de_report_kernel(DeDev d, const double *__restrict__ radius, int np, const int *larch,
        const int *nsucc, bbgpu_de_report *rep) {
    __shared__ double red[32];
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
    int *perm = nullptr, *ribuf = nullptr, *trimbuf = nullptr, *larch_new = nullptr, *nsucc = nullptr;
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
    DA(d.larch, 1); DA(h->larch_new, 1); DA(h->nsucc, 1); DA(d.kslot, 1);
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
    if (h->cfg.kind == BBGPU_DE_JADE) BB_TRY(de_sort(h)); // de/jade.cpp:101
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
    if (h_succ) BB_CUDA(cudaMemcpyAsync(h_succ, h->d.succ[h->fin], (size_t) np * 4, cudaMemcpyDeviceToHost, st));
    if (h_ftrial) BB_CUDA(cudaMemcpyAsync(h_ftrial, h->d.FT, (size_t) np * 8, cudaMemcpyDeviceToHost, st));
    if (h_F) BB_CUDA(cudaMemcpyAsync(h_F, h->d.Fm, (size_t) np * 8, cudaMemcpyDeviceToHost, st));
    if (h_CR) BB_CUDA(cudaMemcpyAsync(h_CR, h->d.CR, (size_t) np * 8, cudaMemcpyDeviceToHost, st));
    BB_CUDA(cudaStreamSynchronize(st));
    return BBGPU_OK;
}

} // extern "C"
