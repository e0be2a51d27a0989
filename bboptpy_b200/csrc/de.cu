// Differential-evolution device engine: JADE (de/jade.cpp:98-206) and L-SHADE
// (de/shade.cpp:96-236) generations behind the C ABI.
//
// The reference updates the swarm IN PLACE inside the candidate loop (asynchronous DE):
// candidate i reads x_best, x_r1, x_r2 that candidates j < i may already have replaced, the
// archive grows / is overwritten inside the loop, and its size feeds the range of later r2
// draws.  A one-shot synchronous kernel is not equivalent.  The engine keeps the exact
// sequential semantics and resolves them in parallel as a fixed point:
//
//   round r: every candidate derives the *signature* of what the sequential loop would have
//            shown it -- for each parent j < i either "old row" or "trial of j, version v", the
//            archive writer of its slot, its r2 -- from the previous round's success flags.  If
//            the signature changed it rebuilds its trial (mutation, binomial crossover, midpoint
//            bound repair), evaluates it with the registry objective in the same pass, and
//            publishes (success, version).  Candidate 0 is final after round 1, candidate i once
//            its predecessors are; a round with no rebuild is the fixed point, which equals the
//            sequential result by induction on i.
//
// Layout: every row (swarm member, archive entry, trial) lives in ONE pool of 4*np rows and is
// addressed through row ids (srow / arch_row / trow).  Greedy replacement, the archive push, the
// sort and L-SHADE's population reduction therefore move ids, never rows: the only row traffic of
// a generation is the 4 parent reads + 1 trial write of every (re)built trial -- the
// 5 * 8n bytes per candidate-evaluation the roofline counts.
//
// One cooperative launch runs a whole generation for one run or for a batch of independent runs
// (same configuration, different seeds): sort + parameters -> rounds (grid.sync between rounds)
// -> finalize (id swaps, adaptation) -> [L-SHADE sort] -> report.  A candidate is built by a
// group of 128 threads with 32-byte row accesses; 4 groups per CTA, 2 CTAs per SM.
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdlib>

#include <cooperative_groups.h>

#include "gemm_tma.cuh" // mbarrier + 1-D bulk copy helpers
#define BB_COS2PI_TABLE 1 // cos(2 pi x) coefficients from constant memory in this translation unit (objectives.cuh)
#include "objectives.cuh"
#include "philox.cuh"
#include "rank.cuh"

namespace bb {

constexpr int DE_MAX_ROUNDS = 4096;
constexpr int DE_GROUP = 128;                 // threads that build one candidate
constexpr int DE_GROUPS = 4;                  // candidates in flight per CTA (8 x 64 registers measured no faster: the
                                              // build is bound by HBM traffic and issue slots, not by resident warps)
constexpr int DE_CTA = DE_GROUP * DE_GROUPS;  // 512
constexpr int DE_CHUNK = 512;                 // candidates whose signatures are checked per pass (work-list capacity)
constexpr int DE_TILE = 512;                  // coordinates per bulk-copy stage (4 parent rows x 4 KB)
constexpr int DE_STAGES = 2;                  // bulk-copy stages per group: the next tile / candidate loads while this one computes
enum { TAG_CR = 1, TAG_F = 2, TAG_IBEST = 3, TAG_R1 = 4, TAG_R2 = 5, TAG_JRAND = 6, TAG_SLOT = 7, TAG_CROSS = 8,
       TAG_RI = 9, TAG_PI = 10, TAG_TRIM = 11, TAG_INIT = 12 };

// One run's device-resident state.  The struct itself lives in device memory; the generation
// kernel advances the evolving scalars (np, fev, gen, cur), so a free-running run needs no host
// traffic between generations.
struct DeDev {
    // configuration
    int kind, n, npcap, h, archive, repaircr, npmin, mfev;
    int64_t ld;
    uint64_t seed;
    double pelite, cdamp, sigma_cr;
    const double *lower, *upper;
    // evolving scalars
    int np, fev, cur /* live side of the double-buffered (srow, F, R) */, last_fin, last_rounds, last_np;
    uint64_t gen;
    // rows
    double *POOL;
    int *srow[2], *trow[2], *arch_row;
    // per member / per candidate
    double *F[2], *R[2] /* ||x_i|| */, *FT, *RT, *DF /* f_parent - f_trial */, *CR, *Fm, *CR1, *MCR, *MF, *mu /* [mucr, muf] */;
    const double *zcr, *cauchy, *U;
    int *ibest, *r1, *r2, *jrand, *slot, *ri;
    const int *trim;
    int *larch, *larch_new, *nsucc, *kslot;
    int *st[2] /* per candidate: version << 3 | trial buffer << 2 | success (0, 1: <=, 2: <) */, *sig;
    bbgpu_de_report *rep;
};

struct DeLaunch {
    int nruns, cpr /* CTAs per run */, kind, obj_id, replay, ntrim, npnew;
    const double *aux;
    int *nrecomp;                 // [DE_MAX_ROUNDS] rebuilds per round, over all runs of the launch
    // objectives evaluated OUTSIDE the kernel (user device functions, the host-callback bridge): the generation is
    // split into launches -- mode 1 begin, mode 2 one round that only builds trials and lists them in `pend`, mode 3
    // finish -- with gather / evaluate / select between the rounds.  mode 0: the fused generation.
    int mode, round;
    int *pend, *pend_count;
    unsigned long long *phase_ns; // [32] %globaltimer at phase boundaries (diagnostics)
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

__device__ __forceinline__ void group_sync(int g) {
    asm volatile("bar.sync %0, %1;" :: "r"(g + 1), "r"(DE_GROUP) : "memory");
}

// per-candidate control parameters for the generation (de/jade.cpp:107-126, de/shade.cpp:104-126)
__device__ __noinline__ void de_params_one(const DeDev &d, int replay, int i) {
    int ri = 0;
    if (d.kind == BBGPU_DE_SHADE) ri = replay ? d.ri[i] : de_int(d, i, 0, TAG_RI, d.h);
    const double mcr = d.kind == BBGPU_DE_JADE ? d.mu[0] : d.MCR[ri];
    const double mf = d.kind == BBGPU_DE_JADE ? d.mu[1] : d.MF[ri];
    double z, c, F;
    if (replay) {
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
    {
        // std::max(0., std::min(v, 1.)) (jade.cpp:109, shade.cpp:106-107): a NaN v -- an L-SHADE memory slot that became 0 / 0
        // without repaircr -- passes std::min unchanged and std::max returns 0.  Explicit, because nvcc reorders min / max
        // chains (profiles/r02_nan_clamp.txt).
        const double v = __dadd_rn(__dmul_rn(z, 0.1), mcr);
        d.CR[i] = v != v ? 0.0 : fmax(0.0, fmin(v, 1.0));
    }
    d.Fm[i] = F;
    d.sig[4 * i + 0] = INT_MIN; // force a build in round 0
    d.st[0][i] = 0;
}

// CTA-wide: push flags -> exclusive prefix counts `pre` and archive targets `tgt` (both in smem, np
// entries each) and, per archive slot, the first and last candidate that writes it (minw / maxw);
// wsum[0..nwarps) scratch, wsum[32] = total pushes
__device__ __forceinline__ void de_scan_targets(const DeDev &d, const int *__restrict__ st, int larch0, unsigned short *pre,
        unsigned short *tgt, int *minw, int *maxw, int *wsum) {
    const int np = d.np, T = blockDim.x;
    const int per = (np + T - 1) / T;
    const int b = threadIdx.x * per, e = min(np, b + per);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int local = 0;
    unsigned mask = 0; // per <= 32 (np <= 8192 and T = 512)
    for (int i = b; i < e; i++) {
        const int s = st[i] & 3;
        const int push = d.archive && (d.kind == BBGPU_DE_JADE ? (s >= 1) : (s == 2));
        mask |= (unsigned) push << (i - b);
        local += push;
    }
    int incl = local;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= off) incl += v;
    }
    __syncthreads(); // previous readers of the tables are done
    if (lane == 31) wsum[wid] = incl;
    for (int a = threadIdx.x; a < np; a += T) { minw[a] = INT_MAX; maxw[a] = -1; }
    __syncthreads();
    if (wid == 0) { // exclusive scan of the warp totals
        const int nw = T >> 5;
        const int v = lane < nw ? wsum[lane] : 0;
        int x = v;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, x, off);
            if (lane >= off) x += y;
        }
        if (lane < nw) wsum[lane] = x - v;
        if (lane == 31) wsum[32] = x;
    }
    __syncthreads();
    int run = wsum[wid] + incl - local;
    for (int i = b; i < e; i++) {
        const int push = (mask >> (i - b)) & 1;
        pre[i] = (unsigned short) run;
        int t = 0xFFFF;
        if (push) {
            const int la = larch0 + run;
            t = la < np ? la : d.slot[i];
            // replayed slot draws exist only where the sequential loop drew one; a transient round may ask elsewhere
            if ((unsigned) t >= (unsigned) np) t = 0xFFFF;
            else { atomicMin(minw + t, i); atomicMax(maxw + t, i); }
        }
        tgt[i] = (unsigned short) t;
        run += push;
    }
    __syncthreads();
}

// Trial coordinate j of candidate i: binomial crossover + current-to-pbest/1 mutation + midpoint repair.
struct DeBuild {
    const double *xi, *xb, *x1, *x2, *lo, *hi, *U;
    double F, CR;
    int jr, replay;
};

__device__ __forceinline__ double de_trial_coord(const DeBuild &b, int j, double x, double vb, double v1, double v2,
        double lo, double hi, double u, int &cnt) {
    const bool cross = (j == b.jr) || (u < b.CR);
    double v = x;
    if (cross) {
        // x + F*(best - x) + F*(xr1 - xr2), left to right, no contraction (de/jade.cpp:253)
        v = __dadd_rn(__dadd_rn(x, __dmul_rn(b.F, __dsub_rn(vb, x))), __dmul_rn(b.F, __dsub_rn(v1, v2)));
        cnt++;
    }
    if (v < lo) v = __dmul_rn(__dadd_rn(lo, x), 0.5);       // midpoint repair (de/jade.cpp:150-156)
    else if (v > hi) v = __dmul_rn(__dadd_rn(hi, x), 0.5);
    return v;
}

__device__ __forceinline__ double u32_unit(uint32_t w) { return u32_to_double(w) * (1.0 / 4294967296.0); }

__device__ __forceinline__ Philox4 de_cross_block(const DeDev &d, int i, int quad) {
    return philox4x32_10((uint32_t) quad, (uint32_t) i, (uint32_t) d.gen,
            ((uint32_t) (d.gen >> 32)) ^ ((uint32_t) TAG_CROSS << 24), (uint32_t) d.seed, (uint32_t) (d.seed >> 32));
}

// A candidate whose inputs changed, queued for a rebuild in this round (48 bytes, shared memory)
struct DeItem {
    int i, j2, s0, s1, s2, mine /* its state word before the rebuild */;
    int rxi, rxb, rx1, rx2, rout /* pool rows: parent, pbest, r1, r2, trial */, jr;
};

__device__ __forceinline__ int st_succ(int w) { return w & 3; }
__device__ __forceinline__ int st_buf(int w) { return (w >> 2) & 1; }
__device__ __forceinline__ int st_ver(int w) { return w >> 3; }

// one THREAD per candidate: derive what the sequential loop would have shown candidate i from the
// previous round's states; unchanged -> carry the state over, changed -> queue a rebuild
template<bool REPLAY>
__device__ __forceinline__ void de_signature(const DeDev &d, int cj, int i, bool valid, int round, int larch0,
        const unsigned short *pre, const unsigned short *tgt, const int *minw, const int *maxw, DeItem *work, int *nwork) {
    const int cur = round & 1, nxt = cur ^ 1;
    const int np = d.np, lane = threadIdx.x & 31;
    const int *st = d.st[cur];
    int jb = 0, j1 = 0, j2 = 0, mine = 0, stb = 0, st1 = 0, st2 = 0, writer = -1, wa = 0, wlo = 0;
    int4 sg = make_int4(0, 0, 0, 0);
    bool search = false;
    if (valid) {
        jb = d.ibest[i]; j1 = d.r1[i];
        sg = *(const int4*) (d.sig + 4 * i);
        mine = st[i];
        stb = jb < i ? st[jb] : 0;
        st1 = j1 < i ? st[j1] : 0;
        const int larch_i = d.archive ? min(np, larch0 + pre[i]) : 0;
        if (REPLAY) {
            j2 = d.r2[i];
        } else {
            j2 = i;
            int t = 0;
            while ((j2 == i || j2 == j1) && (np + larch_i > 2)) j2 = de_int(d, i, t++, TAG_R2, np + larch_i);
        }
        if (j2 >= np) { // archive pick: the last candidate before i that writes this slot in the sequential loop
            wa = j2 - np;
            wlo = minw[wa];
            const int hi = maxw[wa];
            if (wlo >= i) writer = -1;      // nobody before i (INT_MAX: nobody at all)
            else if (hi < i) writer = hi;
            else search = true;             // several writers straddle i: rare
        } else {
            st2 = j2 < i ? st[j2] : 0;
        }
    }
    // the rare straddled slots are searched by the whole warp, one query at a time
    unsigned need = __ballot_sync(0xffffffffu, search);
    while (need) {
        const int src = __ffs(need) - 1;
        need &= need - 1;
        const int qa = __shfl_sync(0xffffffffu, wa, src), qi = __shfl_sync(0xffffffffu, i, src), qlo = __shfl_sync(0xffffffffu, wlo, src);
        int best = qlo;
        for (int j = qlo + 1 + lane; j < qi; j += 32)
            if (tgt[j] == qa) best = j;
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) best = max(best, __shfl_xor_sync(0xffffffffu, best, off));
        if (lane == src) writer = best;
    }
    if (!valid) return;
    const int s0 = st_succ(stb) ? st_ver(stb) : 0;
    const int s1 = st_succ(st1) ? st_ver(st1) : 0;
    const int s2 = j2 >= np ? -2 - writer : (st_succ(st2) ? st_ver(st2) : 0);
    if (sg.x == s0 && sg.y == s1 && sg.z == s2 && sg.w == j2) {
        d.st[nxt][i] = mine;
        return;
    }
    const int *srow = d.srow[cj];
    DeItem it;
    it.i = i; it.j2 = j2; it.s0 = s0; it.s1 = s1; it.s2 = s2; it.mine = mine;
    it.rxi = srow[i];
    it.rxb = st_succ(stb) ? d.trow[st_buf(stb)][jb] : srow[jb];
    it.rx1 = st_succ(st1) ? d.trow[st_buf(st1)][j1] : srow[j1];
    if (j2 >= np) it.rx2 = writer >= 0 ? srow[writer] : d.arch_row[j2 - np];
    else it.rx2 = st_succ(st2) ? d.trow[st_buf(st2)][j2] : srow[j2];
    it.rout = d.trow[1 - st_buf(mine)][i];
    it.jr = d.jrand[i];
    work[atomicAdd(nwork, 1)] = it;
}

// The queued candidates of one 128-thread group (items g, g + 4, ... of the CTA's work list): mutation + crossover +
// repair + objective + greedy test.  The four parent rows of a 512-coordinate tile arrive in shared memory by 1-D bulk
// copies (cp.async.bulk, one elected thread, completion on an mbarrier); the tile after it -- of this candidate or the
// next -- is already in flight while this one is computed, so the loads need no registers and never stall the math.
// The objective terms are written over the consumed parent rows of the stage and added by warp 0 in the objective's
// fixed order before the stage is refilled.
template<int ID, bool REPLAY>
__device__ __noinline__ void de_build_list(const DeDev &d, const DeLaunch &L, int cj, const DeItem *work, int todo, int round,
        int g, double *stage, unsigned long long *full, unsigned &uses, double *gred, int *changed) {
    const int nmine = todo > g ? (todo - g + DE_GROUPS - 1) / DE_GROUPS : 0;
    if (nmine == 0) return;
    const int nxt = (round & 1) ^ 1;
    const int gt = threadIdx.x & (DE_GROUP - 1), lane = threadIdx.x & 31, gwarp = gt >> 5;
    const int n = d.n;
    const int64_t ld = d.ld;
    const int ntiles = (n + DE_TILE - 1) / DE_TILE;
    const int U = nmine * ntiles;
    constexpr int id = ID;
    constexpr bool external = ID < 0; // the objective runs outside the kernel: only build the trial rows
    const int nt = external ? 0 : obj_term_count(id, n);
    constexpr bool rosen = id == BBGPU_OBJ_ROSENBROCK, ackley = id == BBGPU_OBJ_ACKLEY, ellip = id == BBGPU_OBJ_ELLIPSOID;
    const unsigned base = uses;
    auto issue = [&](int u) { // elected thread: the 4 parent rows of (item u / ntiles, tile u % ntiles) -> stage (base + u) & 1
        const DeItem &it = work[g + DE_GROUPS * (u / ntiles)];
        const int t0 = (u % ntiles) * DE_TILE;
        const int cnt = min(DE_TILE, (int) ld - t0); // padded row length: a multiple of 8 doubles
        const unsigned s = (base + u) & (DE_STAGES - 1);
        double *dst = stage + (size_t) s * 4 * DE_TILE;
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); // earlier generic writes to this stage (the terms)
        mbar_expect_tx(full + s, 4 * cnt * 8);
        bulk_load_1d(dst, d.POOL + (int64_t) it.rxi * ld + t0, cnt * 8, full + s);
        bulk_load_1d(dst + DE_TILE, d.POOL + (int64_t) it.rxb * ld + t0, cnt * 8, full + s);
        bulk_load_1d(dst + 2 * DE_TILE, d.POOL + (int64_t) it.rx1 * ld + t0, cnt * 8, full + s);
        bulk_load_1d(dst + 3 * DE_TILE, d.POOL + (int64_t) it.rx2 * ld + t0, cnt * 8, full + s);
    };
    if (gt == 0)
        for (int u = 0; u < DE_STAGES && u < U; u++) issue(u);
    DeBuild b;
    b.lo = d.lower; b.hi = d.upper; b.replay = REPLAY;
    int cnt = 0;
    double nrm = 0.0, acc = 0.0, acc2 = 0.0, fi = 0.0;
    for (int u = 0; u < U; u++) {
        const DeItem &it = work[g + DE_GROUPS * (u / ntiles)];
        const int t = u % ntiles, t0 = t * DE_TILE, i = it.i;
        const int tl = min(DE_TILE, n - t0);
        if (t == 0) {
            b.F = d.Fm[i]; b.CR = d.CR[i]; b.jr = it.jr; b.U = d.U + (int64_t) i * n;
            fi = d.F[cj][i];
            cnt = 0; nrm = 0.0; acc = 0.0; acc2 = 0.0;
        }
        const unsigned s = (base + u) & (DE_STAGES - 1);
        double *S = stage + (size_t) s * 4 * DE_TILE; // rows: parent, pbest, r1, r2
        double *out = d.POOL + (int64_t) it.rout * ld;
        mbar_wait(full + s, ((base + u) / DE_STAGES) & 1);
        double t0v = 0.0, t1v = 0.0, t2v = 0.0, t3v = 0.0, e0 = 0.0, e1 = 0.0, e2 = 0.0, e3 = 0.0;
        const int q = gt; // DE_TILE / 4 == DE_GROUP: one quad per thread and tile
        const bool mine = 4 * q < tl;
        if (mine) {
            const int j = t0 + 4 * q; // rows are 64-byte aligned and padded to a multiple of 8 doubles
            const double2 xa = *(const double2*) (S + 4 * q), xc = *(const double2*) (S + 4 * q + 2);
            const double2 ba = *(const double2*) (S + DE_TILE + 4 * q), bc = *(const double2*) (S + DE_TILE + 4 * q + 2);
            const double2 pa = *(const double2*) (S + 2 * DE_TILE + 4 * q), pc = *(const double2*) (S + 2 * DE_TILE + 4 * q + 2);
            const double2 qa = *(const double2*) (S + 3 * DE_TILE + 4 * q), qc = *(const double2*) (S + 3 * DE_TILE + 4 * q + 2);
            const double2 la = *(const double2*) (b.lo + j), lc = *(const double2*) (b.lo + j + 2);
            const double2 ha = *(const double2*) (b.hi + j), hc = *(const double2*) (b.hi + j + 2);
            double u0, u1, u2, u3;
            if (REPLAY) {
                u0 = b.U[j]; u1 = j + 1 < n ? b.U[j + 1] : 2.0; u2 = j + 2 < n ? b.U[j + 2] : 2.0; u3 = j + 3 < n ? b.U[j + 3] : 2.0;
            } else {
                const Philox4 r = de_cross_block(d, i, j >> 2);
                u0 = u32_unit(r.x); u1 = u32_unit(r.y); u2 = u32_unit(r.z); u3 = u32_unit(r.w);
            }
            double v0, v1 = 0.0, v2 = 0.0, v3 = 0.0;
            v0 = de_trial_coord(b, j, xa.x, ba.x, pa.x, qa.x, la.x, ha.x, u0, cnt);
            if (j + 1 < n) v1 = de_trial_coord(b, j + 1, xa.y, ba.y, pa.y, qa.y, la.y, ha.y, u1, cnt);
            if (j + 2 < n) v2 = de_trial_coord(b, j + 2, xc.x, bc.x, pc.x, qc.x, lc.x, hc.x, u2, cnt);
            if (j + 3 < n) v3 = de_trial_coord(b, j + 3, xc.y, bc.y, pc.y, qc.y, lc.y, hc.y, u3, cnt);
            *(double2*) (out + j) = make_double2(v0, v1);
            *(double2*) (out + j + 2) = make_double2(v2, v3);
            nrm = fma(v0, v0, fma(v1, v1, fma(v2, v2, fma(v3, v3, nrm))));
            double v4 = 0.0;
            if (rosen && j + 4 < n) { // the neighbour's first coordinate, rebuilt here (same inputs, same bits)
                double x4, b4, p4, q4, u4;
                if (4 * q + 4 < tl) { x4 = S[4 * q + 4]; b4 = S[DE_TILE + 4 * q + 4]; p4 = S[2 * DE_TILE + 4 * q + 4]; q4 = S[3 * DE_TILE + 4 * q + 4]; }
                else { // first coordinate of the next tile: not staged yet
                    x4 = d.POOL[(int64_t) it.rxi * ld + j + 4]; b4 = d.POOL[(int64_t) it.rxb * ld + j + 4];
                    p4 = d.POOL[(int64_t) it.rx1 * ld + j + 4]; q4 = d.POOL[(int64_t) it.rx2 * ld + j + 4];
                }
                if (REPLAY) u4 = b.U[j + 4];
                else u4 = u32_unit(de_cross_block(d, i, (j >> 2) + 1).x);
                int dummy = 0;
                v4 = de_trial_coord(b, j + 4, x4, b4, p4, q4, b.lo[j + 4], b.hi[j + 4], u4, dummy);
            }
            double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
            if (ellip) {
                const double2 ea = *(const double2*) (L.aux + j), ec = *(const double2*) (L.aux + j + 2);
                a0 = ea.x; a1 = ea.y; a2 = ec.x; a3 = ec.y;
            }
            if (!external) {
                t0v = obj_term(id, v0, v1, a0, e0);
                t1v = obj_term(id, v1, v2, a1, e1);
                t2v = obj_term(id, v2, v3, a2, e2);
                t3v = obj_term(id, v3, v4, a3, e3);
            }
        }
        if (rosen) group_sync(g); // neighbours have read their look-ahead coordinate from the parent rows
        if (mine && !external) { // terms over the consumed rows: T1 -> pbest row, T2 -> r1 row
            *(double2*) (S + DE_TILE + 4 * q) = make_double2(t0v, t1v);
            *(double2*) (S + DE_TILE + 4 * q + 2) = make_double2(t2v, t3v);
            if (ackley) {
                *(double2*) (S + 2 * DE_TILE + 4 * q) = make_double2(e0, e1);
                *(double2*) (S + 2 * DE_TILE + 4 * q + 2) = make_double2(e2, e3);
            }
        }
        const bool last = t == ntiles - 1;
        if (last) { // the group's crossover count and squared norm ride on the tile barrier
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) {
                cnt += __shfl_xor_sync(0xffffffffu, cnt, off);
                nrm += __shfl_xor_sync(0xffffffffu, nrm, off);
            }
            // double buffered by tile parity: warp 0 may still be reading the previous candidate's sums
            if (lane == 0) { gred[8 * (u & 1) + 2 * gwarp] = (double) cnt; gred[8 * (u & 1) + 2 * gwarp + 1] = nrm; }
        }
        group_sync(g);
        if (gwarp == 0) { // spec order: lane l adds term(i) for i = l, l+32, ... ascending (t0 is a multiple of 32)
            const int kend = min(tl, nt - t0);
            const double *T1 = S + DE_TILE, *T2 = S + 2 * DE_TILE;
            for (int k = lane; k < kend; k += 32) {
                acc = add(acc, T1[k]);
                if (ackley) acc2 = add(acc2, T2[k]);
            }
            __syncwarp();
            if (lane == 0 && u + DE_STAGES < U) issue(u + DE_STAGES); // the stage is free: refill it
            if (last && external) {
                if (lane == 0) { // the trial row is built; fitness and the greedy test follow in de_select_kernel
                    double cn = 0.0, nr = 0.0;
                    for (int w = 0; w < DE_GROUP / 32; w++) { cn += gred[8 * (u & 1) + 2 * w]; nr += gred[8 * (u & 1) + 2 * w + 1]; }
                    d.RT[i] = sqrt(nr);
                    d.CR1[i] = d.repaircr ? __ddiv_rn(cn, (double) n) : b.CR;
                    *(int4*) (d.sig + 4 * i) = make_int4(it.s0, it.s1, it.s2, it.j2);
                    d.r2[i] = it.j2;
                    L.pend[atomicAdd(L.pend_count, 1)] = i;
                }
            } else if (last) {
                const double sum1 = warp_sum(acc);
                const double sum2 = ackley ? warp_sum(acc2) : 0.0;
                const double ft = obj_finish(id, n, sum1, sum2);
                if (lane == 0) {
                    double cn = 0.0, nr = 0.0;
                    for (int w = 0; w < DE_GROUP / 32; w++) { cn += gred[8 * (u & 1) + 2 * w]; nr += gred[8 * (u & 1) + 2 * w + 1]; }
                    const int sx = ft <= fi ? (ft < fi ? 2 : 1) : 0;
                    d.FT[i] = ft;
                    d.DF[i] = fi - ft;
                    d.RT[i] = sqrt(nr);
                    d.CR1[i] = d.repaircr ? __ddiv_rn(cn, (double) n) : b.CR;
                    const int ver = st_ver(it.mine) + ((sx != st_succ(it.mine) || sx > 0) ? 1 : 0);
                    d.st[nxt][i] = (ver << 3) | ((1 - st_buf(it.mine)) << 2) | sx;
                    if (sx != st_succ(it.mine) || sx > 0) atomicAdd(changed, 1); // later candidates may see a change
                    *(int4*) (d.sig + 4 * i) = make_int4(it.s0, it.s1, it.s2, it.j2);
                    d.r2[i] = it.j2;
                }
            }
        }
    }
    uses = base + U;
}

// after the fixed point, one thread per successful candidate: the greedy replacement and the archive push
// as row-id swaps (de/jade.cpp:161-176).  Last writer per archive slot wins; the archived row is the OLD parent.
__device__ __forceinline__ void de_finalize_one(const DeDev &d, int cj, int i, int fin, const unsigned short *tgt, const int *maxw) {
    const int w = d.st[fin][i];
    if (st_succ(w) == 0) return;
    const int a = tgt[i];
    const int tb = st_buf(w);
    const int t = d.trow[tb][i], p = d.srow[cj][i];
    int freed = p;
    if (a != 0xFFFF && maxw[a] == i) { // the last writer of the slot archives its parent
        freed = d.arch_row[a];
        d.arch_row[a] = p;
    }
    d.srow[cj][i] = t;
    d.trow[tb][i] = freed;
    d.F[cj][i] = d.FT[i];
    d.R[cj][i] = d.RT[i];
}

// parameter adaptation (de/jade.cpp:185-205, de/shade.cpp:186-212)
__device__ __noinline__ void de_adapt_block(const DeDev &d, int fin, int *nsucc_out, double *red) {
    const int *st = d.st[fin];
    const bool jade = d.kind == BBGPU_DE_JADE;
    double cnt = 0, s_cr = 0, s_cr2 = 0, s_f = 0, s_f2 = 0, s_w = 0, s_wcr = 0, s_wf = 0, s_wf2 = 0, acc = 0;
    for (int i = threadIdx.x; i < d.np; i += blockDim.x) {
        const int s = st[i] & 3;
        if (s >= 1) acc += 1.0;
        if (jade ? (s >= 1) : (s == 2)) {
            const double cr = d.CR1[i], f = d.Fm[i];
            cnt += 1.0;
            s_cr += cr; s_cr2 += cr * cr; s_f += f; s_f2 += f * f;
            if (!jade) {
                const double w = d.DF[i];
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
            if ((st[i] & 3) >= 1) { const double t = d.CR1[i] - mean; dev2 += t * t; }
        dev2 = block_sum(dev2, red);
    }
    if (threadIdx.x == 0) {
        *nsucc_out = (int) acc;
        if (jade) {
            double meancr = 0.0, meanf = 0.0;
            if (cnt > 0) {
                const double sd = sqrt(dev2 / cnt);
                meancr = sd > d.sigma_cr ? sqrt(s_cr2 / cnt) : s_cr / cnt;
                meanf = (s_f2 / cnt) / (s_f / cnt);
            }
            d.mu[0] = (1.0 - d.cdamp) * d.mu[0] + d.cdamp * meancr;
            d.mu[1] = (1.0 - d.cdamp) * d.mu[1] + d.cdamp * meanf;
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

// archive trimming by random erase (de/shade.cpp:228-235) on the row-index indirection
__device__ __forceinline__ int de_trim_thread(const DeDev &d, int la, int replay, int npnew, int ntrim) {
    int t = 0;
    while (la > npnew) {
        int r;
        if (replay) r = t < ntrim ? d.trim[t] : 0;
        else r = de_int(d, 0, t, TAG_TRIM, la);
        t++;
        const int saved = d.arch_row[r];
        for (int q = r; q + 1 < la; q++) d.arch_row[q] = d.arch_row[q + 1];
        d.arch_row[la - 1] = saved;
        la--;
    }
    return la;
}

// counting rank of the members [first, np) stride `step` (one warp each) and id scatter:
// side `from` -> side `to`, ascending by (fitness, index)
__device__ __noinline__ void de_sort_ids(const DeDev &d, int from, int to, int first, int step) {
    const int lane = threadIdx.x & 31, np = d.np;
    const double *F = d.F[from];
    for (int i = first; i < np; i += step) {
        const double fi = F[i];
        const unsigned long long ki = sortable_key(fi);
        int cnt = 0;
        for (int j = lane; j < np; j += 32) {
            const unsigned long long kj = sortable_key(F[j]);
            cnt += (kj < ki) || (kj == ki && j < i);
        }
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, off);
        if (lane == 0) {
            d.F[to][cnt] = fi;
            d.R[to][cnt] = d.R[from][i];
            d.srow[to][cnt] = d.srow[from][i];
        }
    }
}

// end of the generation, one CTA per run: swarm statistics for converged() (de/jade.cpp:228-245), the
// report, L-SHADE's archive trimming, and the run's evolving scalars
__device__ __noinline__ void de_report_block(DeDev *dg, const DeDev &d, const DeLaunch &L, int side, int np_after, int fin,
        int rounds, double *red) {
    const double *radius = d.R[side], *F = d.F[side];
    double s = 0.0, fb = INFINITY;
    for (int i = threadIdx.x; i < np_after; i += blockDim.x) {
        s += radius[i];
        fb = fmin(fb, F[i]);
    }
    s = block_sum(s, red);
    fb = -block_max(-fb, red);
    const double mean = s / np_after;
    double m2 = 0.0;
    for (int i = threadIdx.x; i < np_after; i += blockDim.x) {
        const double t = radius[i] - mean;
        m2 += t * t;
    }
    m2 = block_sum(m2, red);
    if (threadIdx.x == 0) {
        int la = *d.larch_new;
        if (d.kind == BBGPU_DE_SHADE && d.archive) la = de_trim_thread(d, la, L.replay, L.npnew, L.ntrim);
        *d.larch = la;
        bbgpu_de_report *rep = d.rep;
        rep->np = np_after;
        rep->larch = la;
        rep->nsucc = *d.nsucc;
        rep->fev = d.fev + d.np;
        rep->rounds = rounds;
        rep->kslot = *d.kslot;
        rep->mucr = d.mu[0];
        rep->muf = d.mu[1];
        rep->fbest = fb;
        rep->radius_m2 = m2;
        dg->fev = d.fev + d.np;
        dg->np = np_after;
        dg->gen = d.gen + 1;
        dg->cur = side;
        dg->last_fin = fin;
        dg->last_rounds = rounds;
        dg->last_np = d.np;
    }
}

__global__ void __launch_bounds__(DE_CTA, 1) de_generation_kernel(DeDev *const *__restrict__ runs, DeLaunch L) {
    namespace cg = cooperative_groups;
    cg::grid_group grid = cg::this_grid();
    extern __shared__ __align__(16) unsigned char smraw[];
    __shared__ DeDev d;
    __shared__ double red[32];
    __shared__ double gredbuf[DE_GROUPS][2 * 2 * (DE_GROUP / 32)];
    __shared__ int wsum[33];
    double *stages = (double*) smraw;                          // [DE_GROUPS][DE_STAGES][4 rows][DE_TILE]
    DeItem *work = (DeItem*) (stages + DE_GROUPS * DE_STAGES * 4 * DE_TILE); // [DE_CHUNK]
    __shared__ __align__(8) unsigned long long fullbar[DE_GROUPS][DE_STAGES];
    unsigned uses = 0; // bulk-copy stages this thread's group has consumed so far (stage = uses & 1, parity = uses / 2 & 1)
    if (threadIdx.x == 0) {
        for (int q = 0; q < DE_GROUPS * DE_STAGES; q++) mbar_init(&fullbar[0][0] + q, 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();
    int *minw = (int*) (work + DE_CHUNK);                        // [np], then maxw [np], pre [np] u16, tgt [np] u16
    __shared__ int nwork;
    const int tid = threadIdx.x, warp = tid >> 5, g = tid / DE_GROUP;
    constexpr int WARPS = DE_CTA / 32;
    const int ntasks = L.nruns * L.cpr;
    int loaded = -1;
    auto load_run = [&](int run) {
        if (run == loaded) return;
        __syncthreads();
        const int *src = (const int*) runs[run];
        int *dst = (int*) &d;
        for (int w = tid; w < (int) (sizeof(DeDev) / sizeof(int)); w += DE_CTA) dst[w] = src[w];
        __syncthreads();
        loaded = run;
    };
    int mark = 0;
    auto stamp = [&]() {
        if (L.phase_ns && blockIdx.x == 0 && tid == 0 && mark < 32) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            L.phase_ns[mark] = t;
        }
        mark++;
    };
    auto fine = [&](int round, int k) { // [32 + 4*round + k], rounds 0..7: after scan / signatures / builds / grid.sync
        if (L.phase_ns && blockIdx.x == 0 && tid == 0 && round < 8) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            L.phase_ns[32 + 4 * round + k] = t;
        }
    };
    stamp(); // 0 start

    // one resolution round of one (run, part) task; `changed` counts the rebuilds later candidates can observe
    auto round_task = [&](int run, int part, int round, int *changed) {
        int *maxw = minw + d.np;
        unsigned short *pre = (unsigned short*) (maxw + d.np), *tgt = pre + d.np;
        const int cj = d.kind == BBGPU_DE_JADE ? d.cur ^ 1 : d.cur;
        const int larch0 = *d.larch;
        de_scan_targets(d, d.st[round & 1], larch0, pre, tgt, minw, maxw, wsum);
        fine(round, 0);
        // candidates of this CTA: part, part + cpr, ...; DE_CHUNK of them per pass
        for (int k0 = 0; part + (int64_t) k0 * L.cpr < d.np; k0 += DE_CHUNK) {
            __syncthreads(); // the previous chunk's work list has been consumed
            if (tid == 0) nwork = 0;
            __syncthreads();
            const int64_t i = part + (int64_t) (k0 + tid) * L.cpr;
            const bool valid = i < d.np && tid < DE_CHUNK;
            if (L.replay) de_signature<true>(d, cj, (int) min(i, (int64_t) d.np - 1), valid, round, larch0, pre, tgt, minw, maxw, work, &nwork);
            else de_signature<false>(d, cj, (int) min(i, (int64_t) d.np - 1), valid, round, larch0, pre, tgt, minw, maxw, work, &nwork);
            __syncthreads();
            fine(round, 1);
            const int todo = nwork;
            double *stage = stages + (size_t) g * DE_STAGES * 4 * DE_TILE;
#define DE_CASE(ID) case ID: if (L.replay) de_build_list<ID, true>(d, L, cj, work, todo, round, g, stage, fullbar[g], uses, gredbuf[g], changed); \
                         else de_build_list<ID, false>(d, L, cj, work, todo, round, g, stage, fullbar[g], uses, gredbuf[g], changed); break;
            switch (L.obj_id) {
            DE_CASE(BBGPU_OBJ_SPHERE) DE_CASE(BBGPU_OBJ_ELLIPSOID) DE_CASE(BBGPU_OBJ_ROSENBROCK)
            DE_CASE(BBGPU_OBJ_RASTRIGIN) DE_CASE(BBGPU_OBJ_ACKLEY) DE_CASE(-1)
            default: break;
            }
#undef DE_CASE
        }
    };
    auto finalize_task = [&](int part, int fin) {
        int *maxw = minw + d.np;
        unsigned short *pre = (unsigned short*) (maxw + d.np), *tgt = pre + d.np;
        const int cj = d.kind == BBGPU_DE_JADE ? d.cur ^ 1 : d.cur;
        const int larch0 = *d.larch;
        de_scan_targets(d, d.st[fin], larch0, pre, tgt, minw, maxw, wsum);
        if (part == 0) {
            // DF / CR1 / Fm, not the fitness being replaced below, feed the adaptation
            de_adapt_block(d, fin, d.nsucc, red);
            if (tid == 0) *d.larch_new = d.archive ? min(d.np, larch0 + wsum[32]) : 0;
        }
        for (int i = part * DE_CTA + tid; i < d.np; i += L.cpr * DE_CTA) de_finalize_one(d, cj, i, fin, tgt, maxw);
    };

    // ---- as many runs as CTAs (or more): a CTA owns whole runs.  Nothing is shared between CTAs, so every phase
    // boundary is a CTA barrier, runs drift apart freely and no CTA ever waits for another run's slowest candidate.
    if (L.mode == 0 && L.cpr == 1 && L.nruns >= (int) gridDim.x) {
        __shared__ int solo_changed[2];
        for (int run = blockIdx.x; run < L.nruns; run += gridDim.x) {
            load_run(run);
            if (d.kind == BBGPU_DE_JADE) de_sort_ids(d, d.cur, d.cur ^ 1, warp, WARPS);
            for (int i = tid; i < d.np; i += DE_CTA) de_params_one(d, L.replay, i);
            if (tid < 2) solo_changed[tid] = 0;
            __syncthreads();
            int round = 0;
            while (true) {
                round_task(run, 0, round, &solo_changed[round & 1]);
                __syncthreads();
                const int rebuilt = solo_changed[round & 1];
                round++;
                if (tid == 0) solo_changed[round & 1] = 0;
                __syncthreads();
                if (rebuilt == 0 || round >= DE_MAX_ROUNDS) break;
            }
            const int fin = round & 1;
            finalize_task(0, fin);
            __syncthreads();
            if (d.kind == BBGPU_DE_SHADE) {
                de_sort_ids(d, d.cur, d.cur ^ 1, warp, WARPS);
                __syncthreads();
            }
            const int np_after = d.kind == BBGPU_DE_SHADE ? min(d.np, L.npnew) : d.np;
            de_report_block(runs[run], d, L, d.cur ^ 1, np_after, fin, round, red);
            __syncthreads();
        }
        return;
    }

    // ---- phase 0: JADE sorts the swarm (de/jade.cpp:101); control parameters of the generation
    if (L.mode <= 1) {
        for (int task = blockIdx.x; task < ntasks; task += gridDim.x) {
            const int run = task / L.cpr, part = task % L.cpr;
            load_run(run);
            if (d.kind == BBGPU_DE_JADE) de_sort_ids(d, d.cur, d.cur ^ 1, part * WARPS + warp, L.cpr * WARPS);
            for (int i = part * DE_CTA + tid; i < d.np; i += L.cpr * DE_CTA) de_params_one(d, L.replay, i);
        }
        if (L.mode == 1) return;
        grid.sync();
    }
    stamp(); // 1 sorted + parameters
    if (L.mode == 2) { // one round with the objective outside: build the changed trials, list them, stop
        for (int task = blockIdx.x; task < ntasks; task += gridDim.x) {
            load_run(task / L.cpr);
            round_task(task / L.cpr, task % L.cpr, L.round, nullptr);
        }
        return;
    }

    // ---- resolution rounds
    int round = L.mode == 3 ? L.round : 0;
    while (L.mode == 0) {
        for (int task = blockIdx.x; task < ntasks; task += gridDim.x) {
            const int run = task / L.cpr, part = task % L.cpr;
            load_run(run);
            round_task(run, part, round, L.nrecomp + round);
        }
        fine(round, 2);
        grid.sync();
        fine(round, 3);
        const int rebuilt = *((volatile int*) (L.nrecomp + round));
        round++;
        if (round <= 12) stamp(); // 2.. after each of the first 12 rounds
        if (rebuilt == 0 || round >= DE_MAX_ROUNDS) break;
    }
    const int fin = round & 1;
    mark = 16;

    // ---- finalize: row-id swaps for replacement + archive, parameter adaptation
    for (int task = blockIdx.x; task < ntasks; task += gridDim.x) {
        const int run = task / L.cpr, part = task % L.cpr;
        load_run(run);
        finalize_task(part, fin);
    }
    if (blockIdx.x == 0 && tid == 0)
        for (int r = 0; r < round; r++) L.nrecomp[r] = 0;
    grid.sync();
    stamp(); // 16 finalized

    // ---- L-SHADE: sort, then the population shrinks from the tail (de/shade.cpp:215-225)
    if (L.kind == BBGPU_DE_SHADE) { // every run of a launch has the same kind
        for (int task = blockIdx.x; task < ntasks; task += gridDim.x) {
            const int run = task / L.cpr, part = task % L.cpr;
            load_run(run);
            de_sort_ids(d, d.cur, d.cur ^ 1, part * WARPS + warp, L.cpr * WARPS);
        }
        grid.sync();
    }
    stamp(); // 17 sorted

    for (int task = blockIdx.x; task < ntasks; task += gridDim.x) {
        const int run = task / L.cpr, part = task % L.cpr;
        if (part != 0) continue;
        load_run(run);
        const int np_after = d.kind == BBGPU_DE_SHADE ? min(d.np, L.npnew) : d.np;
        de_report_block(runs[run], d, L, d.cur ^ 1, np_after, fin, round, red);
    }
    stamp(); // 18 end
}

__global__ void __launch_bounds__(256) de_uniform_init_kernel(DeDev d) {
    const int64_t gid = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (int64_t) d.np * d.n) return;
    const int i = (int) (gid / d.n), j = (int) (gid % d.n);
    const Philox4 r = philox4x32_10((uint32_t) j, (uint32_t) i, 0u, (uint32_t) TAG_INIT << 24, (uint32_t) d.seed, (uint32_t) (d.seed >> 32));
    const double u = u53(r.x, r.y);
    d.POOL[(int64_t) i * d.ld + j] = u * (d.upper[j] - d.lower[j]) + d.lower[j];
}

// row ids of a fresh run: swarm rows [0, np), archive pool [np, 2np), trial buffers [2np, 4np)
__global__ void de_init_ids_kernel(DeDev d) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d.npcap) return;
    d.srow[0][i] = i;
    d.srow[1][i] = i;
    d.arch_row[i] = d.npcap + i;
    d.trow[0][i] = 2 * d.npcap + i;
    d.trow[1][i] = 3 * d.npcap + i;
}

// ||x_i||_2 of the initial members (contiguous pool rows)
__global__ void __launch_bounds__(256) de_norms_kernel(DeDev d) {
    const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (i >= d.np) return;
    const double *x = d.POOL + (int64_t) i * d.ld;
    double s = 0.0;
    for (int c = threadIdx.x & 31; c < d.n; c += 32) s = fma(x[c], x[c], s);
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) d.R[0][i] = sqrt(s);
}

// L-SHADE sorts the initial swarm (de/shade.cpp:93): ids only, one CTA
__global__ void __launch_bounds__(1024) de_init_sort_kernel(DeDev d) {
    de_sort_ids(d, 0, 1, threadIdx.x >> 5, 32);
}

// external-objective rounds: pool rows of the trials listed in `pend` (the buffer the candidate did NOT hold before)
__global__ void de_pend_rows_kernel(DeDev d, int round, const int *__restrict__ pend, int count, int *__restrict__ rows) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) return;
    const int i = pend[k];
    rows[k] = d.trow[1 - st_buf(d.st[round & 1][i])][i];
}

// external-objective rounds: fitness of the listed trials -> greedy test, next state word (de/jade.cpp:159-176)
__global__ void de_select_kernel(DeDev d, int round, const int *__restrict__ pend, int count, const double *__restrict__ ft_in,
        int *__restrict__ changed) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) return;
    const int i = pend[k];
    const int cj = d.kind == BBGPU_DE_JADE ? d.cur ^ 1 : d.cur;
    const int mine = d.st[round & 1][i];
    const double ft = ft_in[k], fi = d.F[cj][i];
    const int sx = ft <= fi ? (ft < fi ? 2 : 1) : 0;
    d.FT[i] = ft;
    d.DF[i] = fi - ft;
    const int ver = st_ver(mine) + ((sx != st_succ(mine) || sx > 0) ? 1 : 0);
    d.st[(round & 1) ^ 1][i] = (ver << 3) | ((1 - st_buf(mine)) << 2) | sx;
    if (sx != st_succ(mine) || sx > 0) atomicAdd(changed, 1);
}

// members in swarm order -> contiguous staging rows (for the getters)
__global__ void __launch_bounds__(256) de_gather_rows_kernel(const double *__restrict__ pool, const int *__restrict__ ids,
        double *__restrict__ dst, int rows, int n, int64_t ld) {
    const int r = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (r >= rows) return;
    const double *s = pool + (int64_t) ids[r] * ld;
    double *t = dst + (int64_t) r * n;
    for (int c = threadIdx.x & 31; c < n; c += 32) t[c] = s[c];
}

} // namespace bb

using namespace bb;

struct bbgpu_de {
    bbgpu_ctx *ctx = nullptr;
    bbgpu_de_config cfg;
    int n = 0, npcap = 0, np = 0, fev = 0;
    int64_t ld = 0;
    uint64_t gen = 0;
    DeDev d;                    // host mirror of the device struct (pointers + configuration)
    DeDev *d_dev = nullptr;     // the device-resident struct the kernel reads and advances
    DeDev **d_self = nullptr;   // device array of one pointer: the single-run launch
    DeDev **d_batch = nullptr;  // device array of pointers for the last batch this handle led
    std::vector<bbgpu_de*> batch_members;
    double *stage = nullptr, *Ubuf = nullptr, *zbuf = nullptr, *cbuf = nullptr;
    int *trimbuf = nullptr, *nrecomp = nullptr;
    int *pend = nullptr, *pend_rows = nullptr, *pend_count = nullptr; // external-objective rounds
    double *fpend = nullptr;
    int *h_counts = nullptr; // pinned [2]
    int trim_cap = 0;
    unsigned long long *phase_ns = nullptr;
    int blocks_per_sm = 0;
    bbgpu_de_report *d_rep = nullptr, *h_rep = nullptr;
    std::vector<void*> allocs;
};

namespace {

template<typename T> int de_alloc(bbgpu_de *h, T **p, size_t count, bool zero = true) {
    BB_TRY(dev_alloc(p, count));
    h->allocs.push_back(*p);
    if (zero) BB_CUDA(cudaMemsetAsync(*p, 0, count * sizeof(T), h->ctx->stream));
    return BBGPU_OK;
}

size_t de_smem_bytes(int np) {
    return (size_t) DE_GROUPS * DE_STAGES * 4 * DE_TILE * sizeof(double) + (size_t) DE_CHUNK * sizeof(DeItem) + (size_t) np * (2 * sizeof(int) + 2 * sizeof(unsigned short));
}

// the device struct as the kernel left it (np, cur, last_fin, ...)
int de_pull(bbgpu_de *h) {
    BB_CUDA(cudaMemcpyAsync(&h->d, h->d_dev, sizeof(DeDev), cudaMemcpyDeviceToHost, h->ctx->stream));
    BB_CUDA(cudaStreamSynchronize(h->ctx->stream));
    return BBGPU_OK;
}

int de_push(bbgpu_de *h) {
    BB_CUDA(cudaMemcpyAsync(h->d_dev, &h->d, sizeof(DeDev), cudaMemcpyHostToDevice, h->ctx->stream));
    BB_CUDA(cudaStreamSynchronize(h->ctx->stream)); // h->d is pageable and about to change again
    return BBGPU_OK;
}

int de_npnew(const bbgpu_de *h, int fev_after) {
    if (h->cfg.kind != BBGPU_DE_SHADE) return h->np;
    return (int) std::round((h->cfg.npmin - h->npcap) * ((1.0 * fev_after) / h->cfg.mfev) + h->npcap);
}

// one cooperative launch over `count` runs that share a configuration; lead = hs[0] owns the scratch
int de_launch(bbgpu_de *const *hs, int count, bbgpu_objective *obj, DeDev **d_runs, int replay, int ntrim, int mode = 0,
        int round = 0) {
    bbgpu_de *lead = hs[0];
    bbgpu_ctx *ctx = lead->ctx;
    const int np = lead->np;
    DeLaunch L;
    L.nruns = count;
    const int resident = lead->blocks_per_sm * ctx->sm_count;
    const int want = ceil_div(np, DE_GROUPS);
    L.cpr = std::max(1, std::min(want, resident / count));
    L.mode = mode; L.round = round; L.pend = lead->pend; L.pend_count = lead->pend_count;
    L.kind = lead->cfg.kind; L.obj_id = obj->mode == 0 ? obj->id : -1; L.aux = obj->d_aux; L.replay = replay; L.ntrim = ntrim;
    L.npnew = de_npnew(lead, lead->fev + np);
    L.nrecomp = lead->nrecomp;
    L.phase_ns = lead->phase_ns;
    const int grid = std::min(resident, count * L.cpr);
    const size_t smem = de_smem_bytes(np);
    void *args[] = { (void*) &d_runs, (void*) &L };
    {
        ProfScope prof(ctx, BBGPU_PROF_DE_ROUND);
        BB_CUDA(cudaLaunchCooperativeKernel((const void*) de_generation_kernel, dim3(grid), dim3(DE_CTA), args, smem, ctx->stream));
        BB_TRY(launched(ctx, "de_generation_kernel"));
    }
    if (mode == 1 || mode == 2) return BBGPU_OK; // the generation is not finished yet
    for (int r = 0; r < count; r++) {
        bbgpu_de *h = hs[r];
        h->fev += np;
        h->gen++;
        if (h->cfg.kind == BBGPU_DE_SHADE && L.npnew < h->np) h->np = L.npnew;
    }
    return BBGPU_OK;
}

// A generation whose objective is evaluated outside the fused kernel (user device function, host-callback bridge):
// begin -> [round: build the changed trials -> gather -> evaluate -> greedy test] until a round changes nothing -> finish.
// Same fixed point, same results as the fused path for the same fitness values.
int de_generation_external(bbgpu_de *h, bbgpu_objective *obj, int replay, int ntrim) {
    bbgpu_ctx *ctx = h->ctx;
    cudaStream_t st = ctx->stream;
    bbgpu_de *one[1] = { h };
    BB_TRY(de_launch(one, 1, obj, h->d_self, replay, ntrim, 1));
    BB_TRY(de_pull(h)); // the kernel-side view of cur / np for the helper kernels
    int round = 0;
    for (;; round++) {
        if (round >= DE_MAX_ROUNDS) return fail(BBGPU_ERR_ARG, "DE dependency resolution did not converge");
        BB_CUDA(cudaMemsetAsync(h->pend_count, 0, 2 * sizeof(int), st));
        BB_TRY(de_launch(one, 1, obj, h->d_self, replay, ntrim, 2, round));
        BB_CUDA(cudaMemcpyAsync(h->h_counts, h->pend_count, sizeof(int), cudaMemcpyDeviceToHost, st));
        BB_CUDA(cudaStreamSynchronize(st));
        const int count = h->h_counts[0];
        if (count == 0) { round++; break; }
        de_pend_rows_kernel<<<ceil_div(count, 256), 256, 0, st>>>(h->d, round, h->pend, count, h->pend_rows);
        BB_TRY(launched(ctx, "de_pend_rows_kernel"));
        de_gather_rows_kernel<<<ceil_div(count, 8), 256, 0, st>>>(h->d.POOL, h->pend_rows, h->stage, count, h->n, h->ld);
        BB_TRY(launched(ctx, "de_gather_rows_kernel"));
        BB_TRY(evaluate(ctx, obj, h->stage, h->n, count, h->fpend));
        de_select_kernel<<<ceil_div(count, 256), 256, 0, st>>>(h->d, round, h->pend, count, h->fpend, h->pend_count + 1);
        BB_TRY(launched(ctx, "de_select_kernel"));
        BB_CUDA(cudaMemcpyAsync(h->h_counts + 1, h->pend_count + 1, sizeof(int), cudaMemcpyDeviceToHost, st));
        BB_CUDA(cudaStreamSynchronize(st));
        if (h->h_counts[1] == 0) { round++; break; }
    }
    return de_launch(one, 1, obj, h->d_self, replay, ntrim, 3, round);
}

} // namespace

extern "C" {

int bbgpu_de_create(bbgpu_ctx *ctx, const bbgpu_de_config *cfg, const double *h_lower, const double *h_upper, bbgpu_de **out) {
    BB_ARG(ctx && cfg && h_lower && h_upper && out);
    BB_ARG(cfg->n >= 1 && cfg->np >= 4 && cfg->np <= 6144); // shared-memory tables of the generation kernel
    BB_ARG(cfg->kind == BBGPU_DE_JADE || cfg->kind == BBGPU_DE_SHADE);
    BB_ARG(cfg->kind == BBGPU_DE_JADE || (cfg->h >= 1 && cfg->npmin >= 4 && cfg->mfev > 0));
    BB_CUDA(cudaSetDevice(ctx->device));
    int supported = 0;
    cudaDeviceGetAttribute(&supported, cudaDevAttrCooperativeLaunch, ctx->device);
    if (!supported) return fail(BBGPU_ERR_NODEVICE, "the DE engine needs cooperative kernel launch");
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
    d.kind = cfg->kind; d.n = cfg->n; d.npcap = h->npcap; d.np = h->np; d.h = hh; d.archive = cfg->archive; d.repaircr = cfg->repaircr;
    d.npmin = cfg->npmin; d.mfev = cfg->mfev;
    d.ld = h->ld; d.seed = cfg->seed; d.pelite = cfg->pelite; d.cdamp = cfg->cdamp; d.sigma_cr = cfg->sigma;
    int rc = BBGPU_OK;
    double *lo = nullptr, *hi = nullptr;
#define DA(ptr, count) if (rc == BBGPU_OK) rc = de_alloc(h, &(ptr), (size_t) (count))
    DA(lo, ld); DA(hi, ld);
    DA(d.POOL, 4 * rows * ld); DA(h->stage, rows * (size_t) h->n);
    DA(d.srow[0], rows); DA(d.srow[1], rows); DA(d.trow[0], rows); DA(d.trow[1], rows); DA(d.arch_row, rows);
    DA(d.F[0], rows); DA(d.F[1], rows); DA(d.R[0], rows); DA(d.R[1], rows);
    DA(d.FT, rows); DA(d.RT, rows); DA(d.DF, rows); DA(d.CR, rows); DA(d.Fm, rows); DA(d.CR1, rows);
    DA(d.MCR, hh); DA(d.MF, hh); DA(d.mu, 2);
    DA(d.ibest, rows); DA(d.r1, rows); DA(d.r2, rows); DA(d.jrand, rows); DA(d.slot, rows); DA(d.ri, rows);
    DA(d.larch, 1); DA(d.larch_new, 1); DA(d.nsucc, 1); DA(d.kslot, 1); DA(h->phase_ns, 64);
    DA(d.st[0], rows); DA(d.st[1], rows);
    DA(d.sig, rows * 4); DA(h->nrecomp, DE_MAX_ROUNDS);
    DA(h->pend, rows); DA(h->pend_rows, rows); DA(h->pend_count, 2); DA(h->fpend, rows);
    DA(h->zbuf, rows); DA(h->cbuf, rows);
    DA(h->d_rep, 1); DA(h->d_dev, 1); DA(h->d_self, 1);
#undef DA
    if (rc == BBGPU_OK && cudaMallocHost((void**) &h->h_rep, sizeof(bbgpu_de_report)) != cudaSuccess) rc = fail(BBGPU_ERR_CUDA, "pinned alloc failed");
    if (rc == BBGPU_OK && cudaMallocHost((void**) &h->h_counts, 2 * sizeof(int)) != cudaSuccess) rc = fail(BBGPU_ERR_CUDA, "pinned alloc failed");
    if (rc != BBGPU_OK) { bbgpu_de_destroy(h); return rc; }
    d.lower = lo; d.upper = hi; d.rep = h->d_rep;
    d.zcr = h->zbuf; d.cauchy = h->cbuf;
    cudaStream_t st = ctx->stream;
    cudaMemcpyAsync(lo, h_lower, (size_t) h->n * 8, cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(hi, h_upper, (size_t) h->n * 8, cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(h->d_self, &h->d_dev, sizeof(DeDev*), cudaMemcpyHostToDevice, st);
    {
        // cooperative launch geometry: all CTAs must be co-resident
        const size_t smem = de_smem_bytes(h->npcap);
        if (smem > 48 * 1024 && cudaFuncSetAttribute(de_generation_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem) != cudaSuccess)
            rc = fail(BBGPU_ERR_ARG, "np = %d needs %zu bytes of shared memory per CTA", h->npcap, smem);
        if (rc == BBGPU_OK) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&h->blocks_per_sm, de_generation_kernel, DE_CTA, smem);
        if (rc == BBGPU_OK && h->blocks_per_sm < 1) rc = fail(BBGPU_ERR_ARG, "the DE generation kernel does not fit on an SM at np = %d", h->npcap);
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
    if (h->d_batch) cudaFree(h->d_batch);
    if (h->h_rep) cudaFreeHost(h->h_rep);
    if (h->h_counts) cudaFreeHost(h->h_counts);
    delete h;
    return BBGPU_OK;
}

int bbgpu_de_init(bbgpu_de *h, bbgpu_objective *obj, const double *h_x) {
    BB_ARG(h && obj && obj->n == h->n);
    bbgpu_ctx *ctx = h->ctx;
    cudaStream_t st = ctx->stream;
    DeDev &d = h->d;
    h->np = h->npcap;
    d.np = h->np;
    d.fev = 0; d.gen = 0; d.cur = 0; d.last_fin = 0; d.last_rounds = 0; d.last_np = h->np;
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
    BB_CUDA(cudaMemsetAsync(h->nrecomp, 0, DE_MAX_ROUNDS * sizeof(int), st));
    BB_CUDA(cudaStreamSynchronize(st)); // `half` / `one` are stack temporaries
    de_init_ids_kernel<<<ceil_div(h->npcap, 256), 256, 0, st>>>(d);
    BB_TRY(launched(ctx, "de_init_ids_kernel"));
    if (h_x) {
        BB_CUDA(cudaMemcpy2DAsync(d.POOL, (size_t) h->ld * 8, h_x, (size_t) h->n * 8, (size_t) h->n * 8, h->np, cudaMemcpyHostToDevice, st));
    } else {
        de_uniform_init_kernel<<<ceil_div((int64_t) h->np * h->n, 256), 256, 0, st>>>(d);
        BB_TRY(launched(ctx, "de_uniform_init_kernel"));
    }
    BB_TRY(evaluate(ctx, obj, d.POOL, h->ld, h->np, d.F[0]));
    de_norms_kernel<<<ceil_div(h->np, 8), 256, 0, st>>>(d);
    BB_TRY(launched(ctx, "de_norms_kernel"));
    h->fev += h->np;
    d.fev = h->fev;
    if (h->cfg.kind == BBGPU_DE_SHADE) { // de/shade.cpp:93
        de_init_sort_kernel<<<1, 1024, 0, st>>>(d);
        BB_TRY(launched(ctx, "de_init_sort_kernel"));
        d.cur = 1;
    }
    BB_TRY(de_push(h));
    return BBGPU_OK;
}

int bbgpu_de_generation(bbgpu_de *h, bbgpu_objective *obj, const bbgpu_de_draws *draws, bbgpu_de_report *out) {
    BB_ARG(h && obj && obj->n == h->n);
    bbgpu_ctx *ctx = h->ctx;
    cudaStream_t st = ctx->stream;
    DeDev &d = h->d;
    const int np = h->np, n = h->n;
    if (draws) {
        BB_ARG(draws->zcr && draws->cauchy && draws->ibest && draws->r1 && draws->r2 && draws->jrand && draws->slot && draws->u);
        BB_ARG(h->cfg.kind == BBGPU_DE_JADE || draws->ri);
        bool repush = false;
        if (!h->Ubuf) {
            BB_TRY(dev_alloc(&h->Ubuf, (size_t) h->npcap * n));
            d.U = h->Ubuf;
            repush = true;
        }
        if (draws->ntrim > h->trim_cap) {
            if (h->trimbuf) cudaFree(h->trimbuf);
            BB_TRY(dev_alloc(&h->trimbuf, (size_t) draws->ntrim));
            h->trim_cap = draws->ntrim;
            d.trim = h->trimbuf;
            repush = true;
        }
        if (repush) { // only the replay pointers changed; keep the scalars the kernel advanced
            DeDev live;
            BB_CUDA(cudaMemcpyAsync(&live, h->d_dev, sizeof(DeDev), cudaMemcpyDeviceToHost, st));
            BB_CUDA(cudaStreamSynchronize(st));
            live.U = d.U; live.trim = d.trim;
            d = live;
            BB_TRY(de_push(h));
        }
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
            BB_CUDA(cudaMemcpyAsync(h->trimbuf, draws->trim, (size_t) draws->ntrim * 4, cudaMemcpyHostToDevice, st));
        }
    }
    if (obj->mode != 0) {
        BB_TRY(de_generation_external(h, obj, draws != nullptr, draws ? draws->ntrim : 0));
    } else {
        bbgpu_de *one[1] = { h };
        BB_TRY(de_launch(one, 1, obj, h->d_self, draws != nullptr, draws ? draws->ntrim : 0));
    }
    if (out) {
        BB_CUDA(cudaMemcpyAsync(h->h_rep, h->d_rep, sizeof(bbgpu_de_report), cudaMemcpyDeviceToHost, st));
        BB_CUDA(cudaStreamSynchronize(st));
        *out = *h->h_rep;
    }
    return BBGPU_OK;
}

// One generation of `count` independent runs (same configuration, their own seeds and state) in ONE
// cooperative launch: the rows of all runs stream through the same kernel, so the chip is filled even
// when a single run's dependency rounds are latency bound.  Each run's result is bit-identical to
// bbgpu_de_generation on that handle.
int bbgpu_de_generation_batch(bbgpu_de *const *hs, int count, bbgpu_objective *obj, bbgpu_de_report *out) {
    BB_ARG(hs && count >= 1 && obj);
    if (obj->mode != 0) return fail(BBGPU_ERR_ARG, "the fused DE generation needs a registry objective");
    bbgpu_de *lead = hs[0];
    BB_ARG(lead && obj->n == lead->n);
    for (int r = 1; r < count; r++) {
        const bbgpu_de *h = hs[r];
        BB_ARG(h && h->ctx == lead->ctx && h->n == lead->n && h->np == lead->np && h->npcap == lead->npcap && h->fev == lead->fev);
        BB_ARG(h->cfg.kind == lead->cfg.kind && h->cfg.mfev == lead->cfg.mfev && h->cfg.npmin == lead->cfg.npmin);
    }
    cudaStream_t st = lead->ctx->stream;
    if ((int) lead->batch_members.size() != count || !std::equal(hs, hs + count, lead->batch_members.begin())) {
        std::vector<DeDev*> ptrs((size_t) count);
        for (int r = 0; r < count; r++) ptrs[r] = hs[r]->d_dev;
        BB_CUDA(cudaStreamSynchronize(st));
        if (lead->d_batch) cudaFree(lead->d_batch);
        BB_TRY(dev_alloc(&lead->d_batch, (size_t) count));
        BB_CUDA(cudaMemcpy(lead->d_batch, ptrs.data(), (size_t) count * sizeof(DeDev*), cudaMemcpyHostToDevice));
        lead->batch_members.assign(hs, hs + count);
    }
    BB_TRY(de_launch(hs, count, obj, lead->d_batch, 0, 0));
    if (out) {
        for (int r = 0; r < count; r++)
            BB_CUDA(cudaMemcpyAsync(out + r, hs[r]->d_rep, sizeof(bbgpu_de_report), cudaMemcpyDeviceToHost, st));
        BB_CUDA(cudaStreamSynchronize(st));
    }
    return BBGPU_OK;
}

int bbgpu_de_get_swarm(bbgpu_de *h, double *h_x, double *h_f) {
    BB_ARG(h);
    cudaStream_t st = h->ctx->stream;
    BB_TRY(de_pull(h));
    const DeDev &d = h->d;
    if (h_x) {
        de_gather_rows_kernel<<<ceil_div(d.np, 8), 256, 0, st>>>(d.POOL, d.srow[d.cur], h->stage, d.np, h->n, h->ld);
        BB_TRY(launched(h->ctx, "de_gather_rows_kernel"));
        BB_CUDA(cudaMemcpyAsync(h_x, h->stage, (size_t) d.np * h->n * 8, cudaMemcpyDeviceToHost, st));
    }
    if (h_f) BB_CUDA(cudaMemcpyAsync(h_f, d.F[d.cur], (size_t) d.np * 8, cudaMemcpyDeviceToHost, st));
    BB_CUDA(cudaStreamSynchronize(st));
    return BBGPU_OK;
}

int bbgpu_de_get_archive(bbgpu_de *h, double *h_arch, int *rows) {
    BB_ARG(h && rows);
    cudaStream_t st = h->ctx->stream;
    int larch = 0;
    BB_CUDA(cudaMemcpyAsync(&larch, h->d.larch, 4, cudaMemcpyDeviceToHost, st));
    BB_CUDA(cudaStreamSynchronize(st));
    *rows = larch;
    if (h_arch && larch > 0) {
        de_gather_rows_kernel<<<ceil_div(larch, 8), 256, 0, st>>>(h->d.POOL, h->d.arch_row, h->stage, larch, h->n, h->ld);
        BB_TRY(launched(h->ctx, "de_gather_rows_kernel"));
        BB_CUDA(cudaMemcpyAsync(h_arch, h->stage, (size_t) larch * h->n * 8, cudaMemcpyDeviceToHost, st));
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

// diagnostics: %globaltimer stamps (ns) of the last generation launch this handle led
// [0] start [1] sorted + parameters [2..13] rounds [16] finalized [17] L-SHADE sort [18] end;
// [32 + 4r + k], r < 8: round r after scan / signatures / builds / grid.sync (CTA 0)
int bbgpu_de_phase_times(bbgpu_de *h, uint64_t *out64) {
    BB_ARG(h && out64);
    BB_CUDA(cudaMemcpyAsync(out64, h->phase_ns, 64 * sizeof(uint64_t), cudaMemcpyDeviceToHost, h->ctx->stream));
    BB_CUDA(cudaStreamSynchronize(h->ctx->stream));
    return BBGPU_OK;
}

// per-candidate outcome of the generation just run, in the candidate order of that generation
int bbgpu_de_get_trace(bbgpu_de *h, int32_t *h_succ, double *h_ftrial, double *h_F, double *h_CR) {
    BB_ARG(h);
    cudaStream_t st = h->ctx->stream;
    BB_TRY(de_pull(h));
    const DeDev &d = h->d;
    const int np = d.last_np, fin = d.last_fin;
    if (h_succ) BB_CUDA(cudaMemcpyAsync(h_succ, d.st[fin], (size_t) np * 4, cudaMemcpyDeviceToHost, st));
    if (h_ftrial) BB_CUDA(cudaMemcpyAsync(h_ftrial, d.FT, (size_t) np * 8, cudaMemcpyDeviceToHost, st));
    if (h_F) BB_CUDA(cudaMemcpyAsync(h_F, d.Fm, (size_t) np * 8, cudaMemcpyDeviceToHost, st));
    if (h_CR) BB_CUDA(cudaMemcpyAsync(h_CR, d.CR, (size_t) np * 8, cudaMemcpyDeviceToHost, st));
    BB_CUDA(cudaStreamSynchronize(st));
    if (h_succ)
        for (int i = 0; i < np; i++) h_succ[i] &= 3; // state word -> success code
    return BBGPU_OK;
}

} // extern "C"
