// SSDE on the device (reference: de/ssde.cpp; Zhao et al. 2022: spherical search with an optional DE stage).
//
// Per candidate (ssde.cpp:148-294): a search direction z from live members (three random ones, a top-p one, the best), a
// random 0/1 mask b, and the trial y = x_i + c A^T diag(b) A z with A a random pairwise rotation (randA, :410-432: the
// identity with every consecutive pair of a shuffled index list turned into a [sin e, cos e; -cos e, sin e] block, e = 1e-12);
// greedy selection in place; with `usede` a candidate whose trial failed gets a second, DE/rand-to-best trial from the same
// live members.  Two things make the generation history dependent: the live rows (as in JADE / SaNSDE) and -- with `usede`
// -- R = fev / mfev, which the loop reads while fev counts the second-stage evaluations of the earlier candidates.
// Resolution by parallel fixed-point rounds, as in sansde.cu: a candidate rebuilds when a parent with a smaller index changed
// its accepted row or when the number of second-stage evaluations before it changed.  The reference applies A as two dense
// n x n products per candidate (2 n^2 np multiply-adds per generation, almost all with a zero factor); A has at most two
// non-zeros per row, and adding an exact zero never changes a partial sum, so the two-term form in ascending column order
// gives the same bits.  One thread then walks the candidates in order for the success memories (:296-342).
#include <algorithm>
#include <cmath>
#include <vector>

#include "objectives.cuh"
#include "philox.cuh"
#include "rank.cuh"

struct bbgpu_ssde {
    bbgpu_ctx *ctx = nullptr;
    int n = 0, npinit = 0, np = 0, np_trace = 0, H = 0, usede = 0, repaircr = 1;
    int64_t ld = 0;
    uint64_t seed = 0, gen = 0;
    double *X[2] = { nullptr, nullptr }, *f[2] = { nullptr, nullptr };
    int cur = 0;
    double *Y = nullptr, *W = nullptr, *fy = nullptr, *fu = nullptr, *lower = nullptr, *upper = nullptr;
    double *UB = nullptr, *UCR = nullptr, *RED = nullptr; // mask uniforms, crossover uniforms, redraw values (np x ld each)
    double *zrank = nullptr, *cv = nullptr, *zcr = nullptr, *prank = nullptr, *ci = nullptr, *cri = nullptr, *cr1 = nullptr;
    double *L1 = nullptr, *L2 = nullptr, *LCR = nullptr, *rnorm = nullptr, *rstat = nullptr;
    int *iL = nullptr, *pp = nullptr, *qq = nullptr, *rr = nullptr, *pb = nullptr, *j0 = nullptr, *partner = nullptr, *first = nullptr;
    int *out = nullptr, *changed = nullptr, *dirty = nullptr, *rebuilt = nullptr, *rebuilt2 = nullptr, *nf2 = nullptr, *nf2_prev = nullptr,
        *sortidx = nullptr, *ndirty = nullptr;
    void *evwork = nullptr; // scratch of evaluate_masked (user objectives only; allocated on first use)
    bbgpu_ssde_state *d_state = nullptr, *h_state = nullptr; // k, kCR, counters (pinned host copy)
    int *h_ndirty = nullptr;
};

namespace {

using namespace bb;

__global__ void __launch_bounds__(256) ss_uniform_kernel(double *__restrict__ X, int64_t ld, int n, int np, const double *__restrict__ lo,
        const double *__restrict__ hi, uint64_t seed, uint32_t tag, uint64_t gen, bool scale) {
    const int64_t gid = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    const int pairs = (n + 1) >> 1;
    if (gid >= (int64_t) np * pairs) return;
    const int i = (int) (gid / pairs), jp = (int) (gid % pairs);
    const Philox4 r = philox4x32_10((uint32_t) jp, (uint32_t) i, (uint32_t) gen, (uint32_t) (gen >> 32) ^ tag, (uint32_t) seed, (uint32_t) (seed >> 32));
    const int c = 2 * jp;
    const double u0 = u53(r.x, r.y), u1 = u53(r.z, r.w);
    X[(int64_t) i * ld + c] = scale ? fma(hi[c] - lo[c], u0, lo[c]) : u0;
    if (c + 1 < n) X[(int64_t) i * ld + c + 1] = scale ? fma(hi[c + 1] - lo[c + 1], u1, lo[c + 1]) : u1;
}

// opposition-based points (init, :59-132): x' = lower + upper - x
__global__ void __launch_bounds__(256) ss_opposite_kernel(double *__restrict__ Xo, const double *__restrict__ X, int64_t ld, int n,
        const double *__restrict__ lo, const double *__restrict__ hi) {
    const int i = blockIdx.x;
    for (int q = threadIdx.x; q < n; q += blockDim.x)
        Xo[(int64_t) i * ld + q] = __dsub_rn(__dadd_rn(lo[q], hi[q]), X[(int64_t) i * ld + q]);
}

__global__ void __launch_bounds__(256) ss_permute_kernel(double *__restrict__ Xn, const double *__restrict__ X, int64_t ld, int n,
        const int *__restrict__ idx) {
    const int pos = blockIdx.x, src = idx[pos];
    const double2 *s = reinterpret_cast<const double2*>(X + (int64_t) src * ld);
    double2 *d = reinterpret_cast<double2*>(Xn + (int64_t) pos * ld);
    for (int q = threadIdx.x; 2 * q < n; q += blockDim.x) d[q] = s[q];
}

struct SsStream {
    uint32_t i, g0, g1, k0, k1, slot;
    __device__ Philox4 block() { return philox4x32_10(slot++, i, g0, g1 ^ 0x53530002u, k0, k1); }
    __device__ double uniform() { const Philox4 r = block(); return u53(r.x, r.y); }
    __device__ double normal() { const Philox4 r = block(); double a, b; normal_pair(r, a, b); return a; }
    __device__ int index(int count) { const Philox4 r = block(); return min(count - 1, (int) (u53(r.x, r.y) * count)); }
};

// free-running scalar draws of candidate i (:151-170,225-229,253-258)
__global__ void __launch_bounds__(256) ss_draws_kernel(int np, int n, int H, int itop, uint64_t seed, uint64_t gen, const double *__restrict__ L2,
        int *__restrict__ iL, double *__restrict__ zrank, int *__restrict__ pp, int *__restrict__ qq, int *__restrict__ rr, int *__restrict__ pb,
        double *__restrict__ cv, double *__restrict__ zcr, int *__restrict__ j0) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= np) return;
    SsStream s { (uint32_t) i, (uint32_t) gen, (uint32_t) (gen >> 32), (uint32_t) seed, (uint32_t) (seed >> 32), 0u };
    const int l = s.index(H);
    iL[i] = l;
    zrank[i] = s.normal();
    int a = i, b = i, c = i;
    while (a == i) a = s.index(np);
    while (b == i || b == a) b = s.index(np);
    while (c == i || c == a || c == b) c = s.index(np);
    pp[i] = a; qq[i] = b; rr[i] = c;
    pb[i] = min(np - 1, s.index(itop + 1)); // (the reference's Random::get(0, itop) reads past the swarm when ptop * np >= np)
    double v = tan(3.14159265358979323846 * (s.uniform() - 0.5));
    while (!(v * 0.1 + L2[l] > 0.0)) v = tan(3.14159265358979323846 * (s.uniform() - 0.5));
    cv[i] = v;
    zcr[i] = s.normal();
    j0[i] = s.index(n);
}

// std::max(lo, std::min(hi, v)) as the reference writes its clamps.  A success memory can become 0 / 0 without repaircr, and
// std::min(hi, NaN) is hi; nvcc rewrites a max(lo, min(hi, v)) chain -- written with fmin / fmax or with comparisons -- into
// min(hi, max(lo, v)), which sends a NaN to lo instead (measured: profiles/r02_nan_clamp.txt), so the NaN case is explicit.
__device__ __forceinline__ double ref_clamp(double lo, double hi, double v) {
    if (v != v) return hi;
    const double t = v < hi ? v : hi;
    return lo < t ? t : lo;
}
__device__ __forceinline__ double ref_min(double hi, double v) { return v != v ? hi : (v < hi ? v : hi); }

// per-candidate parameters from the memories (:154-155,225-229,253-254) and the round state
__global__ void __launch_bounds__(256) ss_begin_kernel(int np, const int *__restrict__ iL, const double *__restrict__ zrank, const double *__restrict__ cv,
        const double *__restrict__ zcr, const double *__restrict__ L1, const double *__restrict__ L2, const double *__restrict__ LCR,
        double *__restrict__ prank, double *__restrict__ ci, double *__restrict__ cri, int *__restrict__ out, int *__restrict__ changed,
        int *__restrict__ dirty, int *__restrict__ nf2, int *__restrict__ nf2_prev) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= np) return;
    const int l = iL[i];
    prank[i] = ref_clamp(0.0, 1.0, __dadd_rn(__dmul_rn(zrank[i], 0.1), L1[l]));
    ci[i] = ref_min(1.0, __dadd_rn(__dmul_rn(cv[i], 0.1), L2[l]));
    cri[i] = ref_clamp(0.0, 1.0, __dadd_rn(__dmul_rn(zcr[i], 0.1), LCR[l]));
    out[i] = 0;
    changed[i] = 0;
    dirty[i] = 1;
    nf2[i] = 0;
    nf2_prev[i] = 0;
}

struct SsArgs {
    const double *X, *lo, *hi, *UB, *UCR, *RED, *prank, *ci, *cri;
    double *Y, *W, *cr1;
    const int *pp, *qq, *rr, *pb, *j0, *partner, *first, *out, *dirty, *nf2;
    int *rebuilt, *rebuilt2;
    int64_t ld;
    int n, np, usede, repaircr, fev0, mfev;
    double sinv, cosv;
};

// the row the sequential loop would have shown candidate i for member p
__device__ __forceinline__ const double *ss_row(const SsArgs &a, int p, int i) {
    if (p < i) {
        const int o = a.out[p];
        if (o == 1) return a.Y + (int64_t) p * a.ld;
        if (o == 2) return a.W + (int64_t) p * a.ld;
    }
    return a.X + (int64_t) p * a.ld;
}

// stage 1 (:173-234, computeTrialPoint :434-453): z, work = diag(b) A z, y = clip(x_i + c A^T work).  Everything the three
// phases need from global memory is fetched in the FIRST one (the kernel is a chain of dependent load latencies otherwise:
// 29 us for 1000 candidates, profiles/r02_ncu_new_kernels.txt): z, x_i and a packed word per coordinate -- partner index + 1,
// orientation of the pair, mask bit -- go to shared memory, the later phases read only that.
constexpr unsigned SS_FIRST = 1u << 30, SS_MASK = 1u << 31, SS_PARTNER = (1u << 30) - 1u;

__global__ void __launch_bounds__(128) ss_build1_kernel(const SsArgs a) {
    extern __shared__ double sh[]; // z[n], work[n], x_i[n], n packed words
    const int n = a.n;
    double *z = sh, *work = sh + n, *xs = sh + 2 * n;
    unsigned *code = reinterpret_cast<unsigned*>(sh + 3 * n);
    const int i = blockIdx.x;
    if (threadIdx.x == 0) a.rebuilt[i] = a.dirty[i];
    if (!a.dirty[i]) return;
    const double R = (1.0 * (a.fev0 + i + a.nf2[i])) / a.mfev;
    const double *xi = a.X + (int64_t) i * a.ld, *xp = ss_row(a, a.pp[i], i), *xq = ss_row(a, a.qq[i], i), *xr = ss_row(a, a.rr[i], i),
            *xb = ss_row(a, a.pb[i], i), *x0 = ss_row(a, 0, i);
    const double *lead;
    if (a.usede) lead = R < 0.333 ? xp : (R < 0.666 ? xb : x0);
    else lead = i < 0.5 * a.np ? xp : xb;
    const double pr = a.prank[i], c = a.ci[i];
    const double *ub = a.UB + (int64_t) i * a.ld;
    for (int d = threadIdx.x; d < n; d += blockDim.x) {
        const double xv = xi[d];
        double v = __dsub_rn(__dsub_rn(__dadd_rn(lead[d], xq[d]), xr[d]), xv);
        if (a.usede) v = __dadd_rn(v, __dmul_rn(R, __dsub_rn(xb[d], xq[d])));
        z[d] = v;
        xs[d] = xv;
        code[d] = (unsigned) (a.partner[d] + 1) | (a.first[d] ? SS_FIRST : 0u) | (ub[d] < pr ? SS_MASK : 0u);
    }
    __syncthreads();
    // row d of A: sin e on the diagonal, +cos e (first of its pair) or -cos e at the partner; ascending column order
    for (int d = threadIdx.x; d < n; d += blockDim.x) {
        const unsigned cd = code[d];
        const int q = (int) (cd & SS_PARTNER) - 1;
        double s = 0.0;
        if (q < 0) s = __dadd_rn(s, __dmul_rn(1.0, z[d]));
        else {
            const double aq = (cd & SS_FIRST) ? a.cosv : -a.cosv;
            if (d < q) {
                s = __dadd_rn(s, __dmul_rn(a.sinv, z[d]));
                s = __dadd_rn(s, __dmul_rn(aq, z[q]));
            } else {
                s = __dadd_rn(s, __dmul_rn(aq, z[q]));
                s = __dadd_rn(s, __dmul_rn(a.sinv, z[d]));
            }
        }
        work[d] = __dmul_rn(s, (cd & SS_MASK) ? 1.0 : 0.0);
    }
    __syncthreads();
    // column d of A (= row d of A^T): sin e on the diagonal; the partner's row holds -cos e if d is the first of the pair
    double *y = a.Y + (int64_t) i * a.ld;
    for (int d = threadIdx.x; d < n; d += blockDim.x) {
        const unsigned cd = code[d];
        const int q = (int) (cd & SS_PARTNER) - 1;
        double v = xs[d];
        if (q < 0) v = __dadd_rn(v, __dmul_rn(__dmul_rn(c, 1.0), work[d]));
        else {
            const double aq = (cd & SS_FIRST) ? -a.cosv : a.cosv; // A[q][d]
            if (d < q) {
                v = __dadd_rn(v, __dmul_rn(__dmul_rn(c, a.sinv), work[d]));
                v = __dadd_rn(v, __dmul_rn(__dmul_rn(c, aq), work[q]));
            } else {
                v = __dadd_rn(v, __dmul_rn(__dmul_rn(c, aq), work[q]));
                v = __dadd_rn(v, __dmul_rn(__dmul_rn(c, a.sinv), work[d]));
            }
        }
        y[d] = ref_clamp(a.lo[d], a.hi[d], v);
    }
}

// stage 2 (:250-276) for the candidates whose first trial failed: DE/rand-to-best with crossover and a uniform redraw of
// coordinates that leave the box
__global__ void __launch_bounds__(128) ss_build2_kernel(const SsArgs a, const double *__restrict__ f, const double *__restrict__ fy) {
    __shared__ double red[32];
    const int i = blockIdx.x;
    const bool second = a.rebuilt[i] && !(fy[i] <= f[i]); // rebuilt this round and the first stage was not accepted
    if (threadIdx.x == 0) a.rebuilt2[i] = second ? 1 : 0;
    if (!second) return;
    double *wout = a.W + (int64_t) i * a.ld;
    const int n = a.n;
    const double R = (1.0 * (a.fev0 + i + a.nf2[i])) / a.mfev;
    const double *xi = a.X + (int64_t) i * a.ld, *xp = ss_row(a, a.pp[i], i), *xq = ss_row(a, a.qq[i], i), *xr = ss_row(a, a.rr[i], i),
            *x0 = ss_row(a, 0, i);
    const double *u = a.UCR + (int64_t) i * a.ld, *rd = a.RED + (int64_t) i * a.ld;
    const double CR = a.cri[i];
    const int jr = a.j0[i];
    double count = 0.0;
    for (int d = threadIdx.x; d < n; d += blockDim.x) {
        double v;
        if (d == jr || u[d] <= CR) { // a NaN uniform (not drawn by the replayed run) compares false
            v = __dadd_rn(__dadd_rn(xp[d], __dmul_rn(R, __dsub_rn(x0[d], xq[d]))), __dmul_rn(R, __dsub_rn(x0[d], xr[d])));
            if (v < a.lo[d] || v > a.hi[d]) {
                const double t = rd[d];
                v = t == t ? t : 0.5 * (a.lo[d] + a.hi[d]); // placeholder in a transient round of a replayed run
            }
            count += 1.0;
        } else {
            v = xi[d];
        }
        wout[d] = v;
    }
    count = block_sum(count, red);
    if (threadIdx.x == 0) a.cr1[i] = a.repaircr ? __ddiv_rn(count, (double) n) : CR;
}

// outcomes of the round, what changed, the number of second-stage evaluations before every candidate, who must rebuild
__global__ void __launch_bounds__(1024) ss_select_kernel(int np, int usede, const double *__restrict__ f, const double *__restrict__ fy,
        const double *__restrict__ fu, const int *__restrict__ pp, const int *__restrict__ qq, const int *__restrict__ rr, const int *__restrict__ pb,
        int *__restrict__ out, int *__restrict__ changed, int *__restrict__ dirty, const int *__restrict__ rebuilt, int *__restrict__ nf2,
        int *__restrict__ nf2_prev, int *__restrict__ ndirty) {
    __shared__ int total, carry, wsum[32];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) { total = 0; carry = 0; }
    for (int i = tid; i < np; i += blockDim.x) {
        int now = 0;
        if (fy[i] <= f[i]) now = 1;
        else if (usede && fu[i] <= f[i]) now = 2;
        changed[i] = (now != out[i]) || (now != 0 && rebuilt[i]);
        out[i] = now;
    }
    __syncthreads();
    // exclusive prefix count of the candidates that spent a second evaluation (first stage failed), in chunks of 1024
    for (int k0 = 0; k0 < np; k0 += 1024) {
        const int i = k0 + tid;
        const int flag = (usede && i < np && out[i] != 1) ? 1 : 0;
        int x = flag;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, x, off);
            if (lane >= off) x += y;
        }
        if (lane == 31) wsum[wid] = x;
        __syncthreads();
        if (wid == 0) {
            int s = wsum[lane];
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const int y = __shfl_up_sync(0xffffffffu, s, off);
                if (lane >= off) s += y;
            }
            wsum[lane] = s;
        }
        __syncthreads();
        if (i < np) {
            nf2_prev[i] = nf2[i];
            nf2[i] = carry + (wid ? wsum[wid - 1] : 0) + x - flag;
        }
        __syncthreads();
        if (tid == 0) carry += wsum[31];
        __syncthreads();
    }
    int mine = 0;
    for (int i = tid; i < np; i += blockDim.x) {
        const int a = pp[i], b = qq[i], c = rr[i], e = pb[i];
        int d = (a < i && changed[a]) || (b < i && changed[b]) || (c < i && changed[c]) || (e < i && changed[e]) || (0 < i && changed[0]);
        d = d || (usede && nf2[i] != nf2_prev[i]);
        dirty[i] = d;
        mine += d;
    }
    atomicAdd(&total, mine);
    __syncthreads();
    if (tid == 0) *ndirty = total;
}

// the success memories in candidate order (:237-248,279-293,296-342), one thread; counts the evaluations of the generation
__global__ void ss_memory_kernel(int np, int usede, int H, const double *__restrict__ f, const double *__restrict__ fy, const double *__restrict__ fu,
        const int *__restrict__ out, const double *__restrict__ prank, const double *__restrict__ ci, const double *__restrict__ cr1,
        double *__restrict__ L1, double *__restrict__ L2, double *__restrict__ LCR, bbgpu_ssde_state *st) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double rn = 0.0, rd = 0.0, cn = 0.0, cd = 0.0, xn = 0.0, xd = 0.0;
    int lenS = 0, lenC = 0, evals = 0, nsucc = 0;
    for (int i = 0; i < np; i++) {
        evals++;
        const int o = out[i];
        if (o == 1) {
            nsucc++;
            if (fy[i] < f[i]) {
                const double w = __dsub_rn(f[i], fy[i]), sr = prank[i], sc = ci[i];
                rn = __dadd_rn(rn, __dmul_rn(__dmul_rn(w, sr), sr));
                rd = __dadd_rn(rd, __dmul_rn(w, sr));
                cn = __dadd_rn(cn, __dmul_rn(__dmul_rn(w, sc), sc));
                cd = __dadd_rn(cd, __dmul_rn(w, sc));
                lenS++;
            }
        } else if (usede) {
            evals++;
            if (o == 2) {
                nsucc++;
                if (fu[i] < f[i]) {
                    const double w = __dsub_rn(f[i], fu[i]), c = cr1[i];
                    xn = __dadd_rn(xn, __dmul_rn(__dmul_rn(w, c), c));
                    xd = __dadd_rn(xd, __dmul_rn(w, c));
                    lenC++;
                }
            }
        }
    }
    if (lenS > 0) {
        L1[st->k - 1] = __ddiv_rn(rn, rd);
        L2[st->k - 1] = __ddiv_rn(cn, cd);
        if (++st->k > H) st->k = 1;
    }
    if (usede && lenC > 0) {
        LCR[st->kcr - 1] = __ddiv_rn(xn, xd);
        if (++st->kcr > H) st->kcr = 1;
    }
    st->evals = evals;
    st->nsucc = nsucc;
}

__global__ void __launch_bounds__(128) ss_commit_kernel(double *__restrict__ X, double *__restrict__ f, const double *__restrict__ Y,
        const double *__restrict__ W, const double *__restrict__ fy, const double *__restrict__ fu, const int *__restrict__ out, int64_t ld, int n) {
    const int i = blockIdx.x, o = out[i];
    if (o == 0) return;
    const double2 *s = reinterpret_cast<const double2*>((o == 1 ? Y : W) + (int64_t) i * ld);
    double2 *d = reinterpret_cast<double2*>(X + (int64_t) i * ld);
    for (int q = threadIdx.x; 2 * q < n; q += blockDim.x) d[q] = s[q];
    if (threadIdx.x == 0) f[i] = o == 1 ? fy[i] : fu[i];
}

__global__ void __launch_bounds__(256) ss_norms_kernel(const double *__restrict__ X, int64_t ld, int n, int rows, double *__restrict__ out) {
    const int k = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (k >= rows) return;
    const double *x = X + (int64_t) k * ld;
    double s = 0.0;
    for (int i = lane; i < n; i += 32) s = fma(x[i], x[i], s);
    s = warp_sum(s);
    if (lane == 0) out[k] = sqrt(s);
}

__global__ void __launch_bounds__(1024) ss_radius_kernel(const double *__restrict__ r, int rows, double *__restrict__ out2) {
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

// end of a generation, one CTA: the population the schedule leaves (:354-360, the host's own formula on the evaluation count
// the memory kernel produced), the best fitness and the radius statistic of converged() (:383-408) over the members that stay
// -- everything the host needs from a generation comes back in ONE read
__global__ void __launch_bounds__(1024) ss_finish_kernel(const double *__restrict__ r, const double *__restrict__ fsorted, int np, int npinit,
        int npmin, int fev0, int mfev, bbgpu_ssde_state *st) {
    __shared__ double red[32];
    __shared__ int s_keep;
    if (threadIdx.x == 0) {
        const int fev = fev0 + st->evals;
        const int npnew = static_cast<int>(npinit + (npmin - npinit) * (1. * fev) / mfev);
        s_keep = np > npnew ? max(1, npnew) : np;
    }
    __syncthreads();
    const int keep = s_keep;
    double s = 0.0;
    for (int k = threadIdx.x; k < keep; k += blockDim.x) s += r[k];
    s = block_sum(s, red);
    const double mean = s / keep;
    __syncthreads();
    double m2 = 0.0;
    for (int k = threadIdx.x; k < keep; k += blockDim.x) {
        const double dlt = r[k] - mean;
        m2 = fma(dlt, dlt, m2);
    }
    m2 = block_sum(m2, red);
    if (threadIdx.x == 0) {
        st->np_next = keep;
        st->fbest = fsorted[0];
        st->rmean = mean;
        st->rm2 = m2;
    }
}

// sort `count` members of buffer cur into the other buffer
int ss_sort(bbgpu_ssde *h, int count) {
    bbgpu_ctx *ctx = h->ctx;
    const int c = h->cur, o = c ^ 1;
    BB_TRY(rank_fitness(ctx, h->f[c], count, h->sortidx, h->f[o]));
    ss_permute_kernel<<<count, 128, 0, ctx->stream>>>(h->X[o], h->X[c], h->ld, h->n, h->sortidx);
    BB_TRY(launched(ctx, "ss_permute_kernel"));
    h->cur = o;
    return BBGPU_OK;
}

} // namespace

using namespace bb;

extern "C" {

int bbgpu_ssde_create(bbgpu_ctx *ctx, int n, int npinit, int h_mem, int usede, int repaircr, uint64_t seed, const double *h_lower,
        const double *h_upper, bbgpu_ssde **out) {
    BB_ARG(ctx && out && h_lower && h_upper && n >= 1 && npinit >= 4 && h_mem >= 1);
    BB_CUDA(cudaSetDevice(ctx->device));
    bbgpu_ssde *h = new bbgpu_ssde();
    h->ctx = ctx; h->n = n; h->npinit = npinit; h->np = npinit; h->H = h_mem; h->usede = usede; h->repaircr = repaircr; h->seed = seed;
    h->ld = round_up(n, 8);
    const size_t ld = (size_t) h->ld, rows = (size_t) npinit, rows2 = 2 * rows; // init may hold 2 np points (opposition)
    int rc = BBGPU_OK;
    auto A = [&](double **p, size_t count) {
        if (rc != BBGPU_OK) return;
        rc = dev_alloc(p, count);
        if (rc == BBGPU_OK && cudaMemsetAsync(*p, 0, count * 8, ctx->stream) != cudaSuccess) rc = fail(BBGPU_ERR_CUDA, "memset failed");
    };
    auto I = [&](int **p, size_t count) { if (rc == BBGPU_OK) rc = dev_alloc(p, count); };
    for (int b = 0; b < 2; b++) { A(&h->X[b], rows2 * ld); A(&h->f[b], rows2); }
    A(&h->Y, rows * ld); A(&h->W, rows * ld); A(&h->fy, rows); A(&h->fu, rows); A(&h->lower, ld); A(&h->upper, ld);
    A(&h->UB, rows * ld); A(&h->UCR, rows * ld); A(&h->RED, rows * ld);
    A(&h->zrank, rows); A(&h->cv, rows); A(&h->zcr, rows); A(&h->prank, rows); A(&h->ci, rows); A(&h->cri, rows); A(&h->cr1, rows);
    A(&h->L1, (size_t) h_mem); A(&h->L2, (size_t) h_mem); A(&h->LCR, (size_t) h_mem); A(&h->rnorm, rows2); A(&h->rstat, 8);
    I(&h->iL, rows); I(&h->pp, rows); I(&h->qq, rows); I(&h->rr, rows); I(&h->pb, rows); I(&h->j0, rows); I(&h->partner, ld); I(&h->first, ld);
    I(&h->out, rows); I(&h->changed, rows); I(&h->dirty, rows); I(&h->rebuilt, rows); I(&h->rebuilt2, rows); I(&h->nf2, rows); I(&h->nf2_prev, rows);
    I(&h->sortidx, rows2); I(&h->ndirty, 8);
    if (rc == BBGPU_OK) rc = dev_alloc(&h->d_state, 1);
    if (rc == BBGPU_OK && (cudaMallocHost((void**) &h->h_ndirty, 8 * sizeof(int)) != cudaSuccess ||
            cudaMallocHost((void**) &h->h_state, sizeof(bbgpu_ssde_state)) != cudaSuccess)) rc = fail(BBGPU_ERR_CUDA, "cudaMallocHost failed");
    if (rc != BBGPU_OK) { bbgpu_ssde_destroy(h); return rc; }
    cudaMemcpyAsync(h->lower, h_lower, (size_t) n * 8, cudaMemcpyHostToDevice, ctx->stream);
    cudaMemcpyAsync(h->upper, h_upper, (size_t) n * 8, cudaMemcpyHostToDevice, ctx->stream);
    BB_CUDA(cudaStreamSynchronize(ctx->stream));
    *out = h;
    return BBGPU_OK;
}

int bbgpu_ssde_destroy(bbgpu_ssde *h) {
    if (!h) return BBGPU_OK;
    cudaStreamSynchronize(h->ctx->stream);
    double *bufs[] = { h->X[0], h->X[1], h->f[0], h->f[1], h->Y, h->W, h->fy, h->fu, h->lower, h->upper, h->UB, h->UCR, h->RED, h->zrank, h->cv,
            h->zcr, h->prank, h->ci, h->cri, h->cr1, h->L1, h->L2, h->LCR, h->rnorm, h->rstat };
    for (double *b : bufs) if (b) cudaFree(b);
    int *ibufs[] = { h->iL, h->pp, h->qq, h->rr, h->pb, h->j0, h->partner, h->first, h->out, h->changed, h->dirty, h->rebuilt, h->rebuilt2, h->nf2, h->nf2_prev,
            h->sortidx, h->ndirty };
    for (int *b : ibufs) if (b) cudaFree(b);
    if (h->d_state) cudaFree(h->d_state);
    if (h->evwork) cudaFree(h->evwork);
    if (h->h_ndirty) cudaFreeHost(h->h_ndirty);
    if (h->h_state) cudaFreeHost(h->h_state);
    delete h;
    return BBGPU_OK;
}

// init(): ssde.cpp:59-132.  np uniform points (h_x: replayed), with `usede` their opposites, all evaluated, sorted, the best np
// kept; memories = 0.5.  *fev_out: evaluations spent (np or 2 np).
int bbgpu_ssde_init(bbgpu_ssde *h, bbgpu_objective *obj, const double *h_x, int *fev_out) {
    BB_ARG(h && obj && obj->n == h->n && fev_out);
    bbgpu_ctx *ctx = h->ctx;
    cudaStream_t st = ctx->stream;
    const int np = h->npinit, n = h->n;
    h->np = np;
    h->cur = 0;
    h->gen = 0;
    if (h_x) {
        BB_CUDA(cudaMemcpy2DAsync(h->X[0], (size_t) h->ld * 8, h_x, (size_t) n * 8, (size_t) n * 8, np, cudaMemcpyHostToDevice, st));
    } else {
        const int64_t total = (int64_t) np * ((n + 1) / 2);
        ss_uniform_kernel<<<ceil_div(total, 256), 256, 0, st>>>(h->X[0], h->ld, n, np, h->lower, h->upper, h->seed, 0x53530001u, 0, true);
        BB_TRY(launched(ctx, "ss_uniform_kernel"));
    }
    int count = np;
    if (h->usede) {
        ss_opposite_kernel<<<np, 128, 0, st>>>(h->X[0] + (int64_t) np * h->ld, h->X[0], h->ld, n, h->lower, h->upper);
        BB_TRY(launched(ctx, "ss_opposite_kernel"));
        count = 2 * np;
    }
    BB_TRY(evaluate(ctx, obj, h->X[0], h->ld, count, h->f[0]));
    BB_TRY(ss_sort(h, count)); // the best np of them are the first np rows of the sorted buffer
    std::vector<double> half((size_t) h->H, 0.5);
    for (double *L : { h->L1, h->L2, h->LCR }) BB_CUDA(cudaMemcpyAsync(L, half.data(), (size_t) h->H * 8, cudaMemcpyHostToDevice, st));
    h->h_state->k = 1; h->h_state->kcr = 1; h->h_state->evals = 0; h->h_state->nsucc = 0;
    BB_CUDA(cudaMemcpyAsync(h->d_state, h->h_state, sizeof(bbgpu_ssde_state), cudaMemcpyHostToDevice, st));
    BB_CUDA(cudaStreamSynchronize(st));
    *fev_out = count;
    return BBGPU_OK;
}

// One generation (ssde.cpp:134-361): up to the sort as the reference runs it, then the population reduction (:354-360,
// npmin: its floor) and the statistics the host's tests need, all returned in `state`.  h_iwork: the shuffled index list behind A (randA, :413; n entries).  fev / mfev: the
// evaluation count at the start of the generation and the budget (R = fev / mfev is read inside the candidate loop).
// draws == NULL -> Philox.  state (out): memory cursors, evaluations spent, trials accepted.
int bbgpu_ssde_generation(bbgpu_ssde *h, bbgpu_objective *obj, const int32_t *h_iwork, int fev, int mfev, double ptop, int npmin,
        const bbgpu_ssde_draws *draws, bbgpu_ssde_state *state, int *rounds_out) {
    BB_ARG(h && obj && obj->n == h->n && h_iwork && state && mfev > 0 && npmin >= 1);
    bbgpu_ctx *ctx = h->ctx;
    cudaStream_t st = ctx->stream;
    const int np = h->np, n = h->n, c = h->cur;
    h->np_trace = np;
    ProfScope prof(ctx, BBGPU_PROF_DE_ROUND);
    // randA (:410-432): partner / orientation of every index from the shuffled list
    std::vector<int> partner((size_t) n, -1), first((size_t) n, 0);
    for (int ii = 1; ii <= n / 2; ii++) {
        const int a = h_iwork[2 * (ii - 1)], b = h_iwork[2 * (ii - 1) + 1];
        BB_ARG(a >= 0 && a < n && b >= 0 && b < n);
        partner[a] = b; partner[b] = a; first[a] = 1;
    }
    BB_CUDA(cudaMemcpyAsync(h->partner, partner.data(), (size_t) n * 4, cudaMemcpyHostToDevice, st));
    BB_CUDA(cudaMemcpyAsync(h->first, first.data(), (size_t) n * 4, cudaMemcpyHostToDevice, st));
    const int itop = std::max(1, (int) (ptop * np));
    if (obj->mode != 0 && h->evwork == nullptr) BB_CUDA(cudaMalloc(&h->evwork, evaluate_masked_work_bytes(n, h->npinit)));
    if (draws) {
        BB_ARG(draws->iL && draws->zrank && draws->p && draws->q && draws->r && draws->pbest && draws->cauchy && draws->ub);
        BB_ARG(!h->usede || (draws->zcr && draws->j0 && draws->ucr && draws->redraw));
        for (int i = 0; i < np; i++) { // replayed indices address device rows: check them here
            BB_ARG(draws->iL[i] >= 0 && draws->iL[i] < h->H);
            BB_ARG(draws->p[i] >= 0 && draws->p[i] < np && draws->q[i] >= 0 && draws->q[i] < np && draws->r[i] >= 0 && draws->r[i] < np);
            BB_ARG(draws->pbest[i] >= 0 && draws->pbest[i] < np);
            BB_ARG(!h->usede || (draws->j0[i] >= 0 && draws->j0[i] < n));
        }
        const struct { int *d; const int32_t *s; } iv[] = { { h->iL, draws->iL }, { h->pp, draws->p }, { h->qq, draws->q }, { h->rr, draws->r },
                { h->pb, draws->pbest }, { h->j0, draws->j0 } };
        for (const auto &v : iv) if (v.s) BB_CUDA(cudaMemcpyAsync(v.d, v.s, (size_t) np * 4, cudaMemcpyHostToDevice, st));
        BB_CUDA(cudaMemcpyAsync(h->zrank, draws->zrank, (size_t) np * 8, cudaMemcpyHostToDevice, st));
        BB_CUDA(cudaMemcpyAsync(h->cv, draws->cauchy, (size_t) np * 8, cudaMemcpyHostToDevice, st));
        BB_CUDA(cudaMemcpy2DAsync(h->UB, (size_t) h->ld * 8, draws->ub, (size_t) n * 8, (size_t) n * 8, np, cudaMemcpyHostToDevice, st));
        if (h->usede) {
            BB_CUDA(cudaMemcpyAsync(h->zcr, draws->zcr, (size_t) np * 8, cudaMemcpyHostToDevice, st));
            BB_CUDA(cudaMemcpy2DAsync(h->UCR, (size_t) h->ld * 8, draws->ucr, (size_t) n * 8, (size_t) n * 8, np, cudaMemcpyHostToDevice, st));
            BB_CUDA(cudaMemcpy2DAsync(h->RED, (size_t) h->ld * 8, draws->redraw, (size_t) n * 8, (size_t) n * 8, np, cudaMemcpyHostToDevice, st));
        }
    } else {
        ss_draws_kernel<<<ceil_div(np, 256), 256, 0, st>>>(np, n, h->H, itop, h->seed, h->gen, h->L2, h->iL, h->zrank, h->pp, h->qq, h->rr, h->pb,
                h->cv, h->zcr, h->j0);
        BB_TRY(launched(ctx, "ss_draws_kernel"));
        const int64_t total = (int64_t) np * ((n + 1) / 2);
        ss_uniform_kernel<<<ceil_div(total, 256), 256, 0, st>>>(h->UB, h->ld, n, np, h->lower, h->upper, h->seed, 0x53530003u, h->gen, false);
        BB_TRY(launched(ctx, "ss_uniform_kernel"));
        if (h->usede) {
            ss_uniform_kernel<<<ceil_div(total, 256), 256, 0, st>>>(h->UCR, h->ld, n, np, h->lower, h->upper, h->seed, 0x53530004u, h->gen, false);
            BB_TRY(launched(ctx, "ss_uniform_kernel"));
            ss_uniform_kernel<<<ceil_div(total, 256), 256, 0, st>>>(h->RED, h->ld, n, np, h->lower, h->upper, h->seed, 0x53530005u, h->gen, true);
            BB_TRY(launched(ctx, "ss_uniform_kernel"));
        }
    }
    h->gen++;
    ss_begin_kernel<<<ceil_div(np, 256), 256, 0, st>>>(np, h->iL, h->zrank, h->cv, h->zcr, h->L1, h->L2, h->LCR, h->prank, h->ci, h->cri, h->out,
            h->changed, h->dirty, h->nf2, h->nf2_prev);
    BB_TRY(launched(ctx, "ss_begin_kernel"));
    SsArgs a {};
    a.X = h->X[c]; a.lo = h->lower; a.hi = h->upper; a.UB = h->UB; a.UCR = h->UCR; a.RED = h->RED; a.prank = h->prank; a.ci = h->ci; a.cri = h->cri;
    a.Y = h->Y; a.W = h->W; a.cr1 = h->cr1; a.pp = h->pp; a.qq = h->qq; a.rr = h->rr; a.pb = h->pb; a.j0 = h->j0; a.partner = h->partner;
    a.first = h->first; a.out = h->out; a.dirty = h->dirty; a.nf2 = h->nf2; a.rebuilt = h->rebuilt; a.rebuilt2 = h->rebuilt2; a.ld = h->ld; a.n = n; a.np = np;
    a.usede = h->usede; a.repaircr = h->repaircr; a.fev0 = fev; a.mfev = mfev; a.sinv = std::sin(1e-12); a.cosv = std::cos(1e-12);
    const size_t smem = (size_t) 3 * n * 8 + (size_t) n * 4;
    if (smem > 200 * 1024) return fail(BBGPU_ERR_ARG, "SSDE: n = %d is beyond the shared memory of a CTA (%zu bytes)", n, smem);
    if (smem > 48 * 1024) BB_CUDA(cudaFuncSetAttribute(ss_build1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
    // Rounds.  A round past the fixed point changes nothing (no candidate is dirty: the builds return at once, the evaluation
    // and the selection reproduce their outputs), so the first SS_SPEC rounds -- what a generation usually needs -- are enqueued
    // back to back and the host reads their dirty counts in one go; further rounds, if any, one at a time.
    constexpr int SS_SPEC = 3;
    auto enqueue_round = [&](int slot) -> int {
        ss_build1_kernel<<<np, 128, smem, st>>>(a);
        BB_TRY(launched(ctx, "ss_build1_kernel"));
        BB_TRY(evaluate_masked(ctx, obj, h->Y, h->ld, np, h->fy, h->rebuilt, h->evwork)); // a user objective sees the rebuilt trials only
        if (h->usede) {
            ss_build2_kernel<<<np, 128, 0, st>>>(a, h->f[c], h->fy);
            BB_TRY(launched(ctx, "ss_build2_kernel"));
            BB_TRY(evaluate_masked(ctx, obj, h->W, h->ld, np, h->fu, h->rebuilt2, h->evwork));
        }
        ss_select_kernel<<<1, 1024, 0, st>>>(np, h->usede, h->f[c], h->fy, h->fu, h->pp, h->qq, h->rr, h->pb, h->out, h->changed, h->dirty,
                h->rebuilt, h->nf2, h->nf2_prev, h->ndirty + slot);
        return launched(ctx, "ss_select_kernel");
    };
    int rounds = 0;
    bool fixed = false;
    const int spec = obj->mode == 0 ? SS_SPEC : 1; // a user objective is not evaluated on speculation
    for (int r = 0; r < spec; r++) BB_TRY(enqueue_round(r));
    BB_CUDA(cudaMemcpyAsync(h->h_ndirty, h->ndirty, spec * sizeof(int), cudaMemcpyDeviceToHost, st));
    BB_CUDA(cudaStreamSynchronize(st));
    for (int r = 0; r < spec && !fixed; r++) {
        rounds++;
        fixed = h->h_ndirty[r] == 0;
    }
    while (!fixed) {
        BB_TRY(enqueue_round(0));
        BB_CUDA(cudaMemcpyAsync(h->h_ndirty, h->ndirty, sizeof(int), cudaMemcpyDeviceToHost, st));
        BB_CUDA(cudaStreamSynchronize(st));
        rounds++;
        fixed = h->h_ndirty[0] == 0;
        if (rounds > 2 * np + 2) return fail(BBGPU_ERR_CUDA, "SSDE rounds did not reach a fixed point");
    }
    ss_memory_kernel<<<1, 32, 0, st>>>(np, h->usede, h->H, h->f[c], h->fy, h->fu, h->out, h->prank, h->ci, h->cr1, h->L1, h->L2, h->LCR, h->d_state);
    BB_TRY(launched(ctx, "ss_memory_kernel"));
    ss_commit_kernel<<<np, 128, 0, st>>>(h->X[c], h->f[c], h->Y, h->W, h->fy, h->fu, h->out, h->ld, n);
    BB_TRY(launched(ctx, "ss_commit_kernel"));
    BB_TRY(ss_sort(h, np));
    ss_norms_kernel<<<ceil_div(np, 8), 256, 0, st>>>(h->X[h->cur], h->ld, n, np, h->rnorm);
    BB_TRY(launched(ctx, "ss_norms_kernel"));
    ss_finish_kernel<<<1, 1024, 0, st>>>(h->rnorm, h->f[h->cur], np, h->npinit, npmin, fev, mfev, h->d_state);
    BB_TRY(launched(ctx, "ss_finish_kernel"));
    BB_CUDA(cudaMemcpyAsync(h->h_state, h->d_state, sizeof(bbgpu_ssde_state), cudaMemcpyDeviceToHost, st));
    BB_CUDA(cudaStreamSynchronize(st));
    *state = *h->h_state;
    h->np = state->np_next; // the swarm is sorted: the worst members drop off the end
    if (rounds_out) *rounds_out = rounds;
    return BBGPU_OK;
}

// population reduction (:354-360): the swarm is sorted, the worst members drop off the end
int bbgpu_ssde_set_population(bbgpu_ssde *h, int np) {
    BB_ARG(h && np >= 1 && np <= h->np);
    h->np = np;
    return BBGPU_OK;
}

int bbgpu_ssde_get_swarm(bbgpu_ssde *h, double *h_x, double *h_f) {
    BB_ARG(h);
    cudaStream_t st = h->ctx->stream;
    const int c = h->cur;
    if (h_x) BB_CUDA(cudaMemcpy2DAsync(h_x, (size_t) h->n * 8, h->X[c], (size_t) h->ld * 8, (size_t) h->n * 8, h->np, cudaMemcpyDeviceToHost, st));
    if (h_f) BB_CUDA(cudaMemcpyAsync(h_f, h->f[c], (size_t) h->np * 8, cudaMemcpyDeviceToHost, st));
    BB_CUDA(cudaStreamSynchronize(st));
    return BBGPU_OK;
}

// success memories: H doubles each; any pointer may be NULL
int bbgpu_ssde_get_memory(bbgpu_ssde *h, double *h_L1, double *h_L2, double *h_LCR) {
    BB_ARG(h);
    cudaStream_t st = h->ctx->stream;
    if (h_L1) BB_CUDA(cudaMemcpyAsync(h_L1, h->L1, (size_t) h->H * 8, cudaMemcpyDeviceToHost, st));
    if (h_L2) BB_CUDA(cudaMemcpyAsync(h_L2, h->L2, (size_t) h->H * 8, cudaMemcpyDeviceToHost, st));
    if (h_LCR) BB_CUDA(cudaMemcpyAsync(h_LCR, h->LCR, (size_t) h->H * 8, cudaMemcpyDeviceToHost, st));
    BB_CUDA(cudaStreamSynchronize(st));
    return BBGPU_OK;
}

int bbgpu_ssde_get_trace(bbgpu_ssde *h, int32_t *h_succ, double *h_ftrial) {
    BB_ARG(h);
    cudaStream_t st = h->ctx->stream;
    const int np = h->np_trace;
    std::vector<int32_t> out((size_t) np);
    std::vector<double> fy((size_t) np), fu((size_t) np);
    BB_CUDA(cudaMemcpyAsync(out.data(), h->out, (size_t) np * 4, cudaMemcpyDeviceToHost, st));
    BB_CUDA(cudaMemcpyAsync(fy.data(), h->fy, (size_t) np * 8, cudaMemcpyDeviceToHost, st));
    BB_CUDA(cudaMemcpyAsync(fu.data(), h->fu, (size_t) np * 8, cudaMemcpyDeviceToHost, st));
    BB_CUDA(cudaStreamSynchronize(st));
    for (int i = 0; i < np; i++) {
        if (h_succ) h_succ[i] = out[i];
        if (h_ftrial) h_ftrial[i] = (out[i] == 1 || !h->usede) ? fy[i] : fu[i];
    }
    return BBGPU_OK;
}

int bbgpu_ssde_radius(bbgpu_ssde *h, double *h_mean, double *h_m2) {
    BB_ARG(h && h_mean && h_m2);
    cudaStream_t st = h->ctx->stream;
    ss_norms_kernel<<<ceil_div(h->np, 8), 256, 0, st>>>(h->X[h->cur], h->ld, h->n, h->np, h->rnorm);
    BB_TRY(launched(h->ctx, "ss_norms_kernel"));
    ss_radius_kernel<<<1, 1024, 0, st>>>(h->rnorm, h->np, h->rstat);
    BB_TRY(launched(h->ctx, "ss_radius_kernel"));
    double out[2];
    BB_CUDA(cudaMemcpyAsync(out, h->rstat, 16, cudaMemcpyDeviceToHost, st));
    BB_CUDA(cudaStreamSynchronize(st));
    *h_mean = out[0];
    *h_m2 = out[1];
    return BBGPU_OK;
}

} // extern "C"
