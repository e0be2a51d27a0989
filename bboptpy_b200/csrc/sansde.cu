// SaNSDE on the device (reference: de/sansde.cpp; Yang, Tang, Yao 2008).
//
// Like JADE the reference updates the swarm IN PLACE inside the candidate loop (:94-177): candidate i reads the best member
// swarm[0] and three random members, any of which an earlier candidate j < i may already have replaced.  Unlike JADE, the
// number and the values of the random draws of a candidate do not depend on what happened before it (no archive whose size
// feeds a range), so the draws of a whole generation are generated (or replayed) up front and only the PARENT ROWS are
// history dependent.  The generation is resolved by parallel fixed-point rounds:
//   round 0   every candidate builds its trial from the generation-start rows, all trials are evaluated, success flags set;
//   round r   a candidate one of whose parents (with a smaller index) changed its outcome or its accepted trial in round r-1
//             rebuilds from "trial of p if p succeeded else old row of p" and is re-evaluated;
//   stop      when nothing changed.  Candidate 0 has no earlier parent, so it is final after round 0, and candidate i once its
//             parents are: by induction the fixed point is the sequential result.  Depth = longest dependency chain (3 - 8
//             rounds at np = 1000).
// The adaptation sums (_pns, _crrec, _fpns ...) are then accumulated by ONE thread in candidate order, so they carry the
// reference's rounding, and the accepted trials are committed.  The sort at the start of iterate() permutes whole members
// (position, fitness and the member's CR) through the ranking kernels.
#include <algorithm>
#include <cmath>

#include "objectives.cuh"
#include "philox.cuh"
#include "rank.cuh"

struct bbgpu_sde {
    bbgpu_ctx *ctx = nullptr;
    int n = 0, np = 0, repaircr = 1;
    int64_t ld = 0;
    uint64_t seed = 0, gen = 0;
    double *X[2] = { nullptr, nullptr }, *f[2] = { nullptr, nullptr }, *cr[2] = { nullptr, nullptr }; // members, double buffered by the sort
    int cur = 0;
    double *T = nullptr, *ftrial = nullptr, *U = nullptr, *lower = nullptr, *upper = nullptr;
    double *zcr = nullptr, *F = nullptr, *cr1 = nullptr, *sortval = nullptr, *rnorm = nullptr, *rstat = nullptr;
    int *ifs = nullptr, *r1 = nullptr, *r2 = nullptr, *r3 = nullptr, *ist = nullptr, *jrand = nullptr, *sortidx = nullptr;
    int *succ = nullptr, *changed = nullptr, *dirty = nullptr, *rebuilt = nullptr, *ndirty = nullptr;
    bbgpu_sde_accum *d_acc = nullptr;
    int *h_ndirty = nullptr; // pinned
    bbgpu_sde_accum *h_acc = nullptr; // pinned
    int fev = 0;
    // the scalars of a generation live in device memory, so that a generation is the same work every time and can be replayed
    // as a CUDA graph
    struct Params { double crm, fp, p; unsigned long long gen; int refresh, pad; };
    Params *d_par = nullptr, *h_par = nullptr; // device / pinned
    // two captured halves of a free-running generation (up to the dirty counts of the speculative rounds; from the adaptation
    // sums to the radius statistic), per parity of the swarm buffer the generation starts from
    cudaGraphExec_t g_head[2] = { nullptr, nullptr }, g_tail[2] = { nullptr, nullptr };
    int64_t g_head_kernels[2] = { 0, 0 }, g_tail_kernels[2] = { 0, 0 }, g_head_evals[2] = { 0, 0 };
    bbgpu_objective *g_obj = nullptr;
    int g_obj_id = -1;            // (a new objective handle can reuse the address of a destroyed one)
    void *evwork = nullptr;       // scratch of evaluate_masked (user objectives only; allocated on first use)
    const double *g_obj_aux = nullptr;
};

namespace {

using namespace bb;

// x = lower + (upper - lower) u, the shape of uniform_real_distribution (libstdc++: (b - a) * canonical + a)
__global__ void __launch_bounds__(256) sde_uniform_kernel(double *__restrict__ X, int64_t ld, int n, int np, const double *__restrict__ lo,
        const double *__restrict__ hi, uint64_t seed) {
    const int64_t gid = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    const int pairs = (n + 1) >> 1;
    if (gid >= (int64_t) np * pairs) return;
    const int i = (int) (gid / pairs), jp = (int) (gid % pairs);
    const Philox4 r = philox4x32_10((uint32_t) jp, (uint32_t) i, 0u, 0x53440001u, (uint32_t) seed, (uint32_t) (seed >> 32));
    const int c = 2 * jp;
    X[(int64_t) i * ld + c] = fma(hi[c] - lo[c], u53(r.x, r.y), lo[c]);
    if (c + 1 < n) X[(int64_t) i * ld + c + 1] = fma(hi[c + 1] - lo[c + 1], u53(r.z, r.w), lo[c + 1]);
}

// whole members follow the ranking: Xn[pos] = X[idx[pos]] (f comes sorted from the ranking itself)
__global__ void __launch_bounds__(256) sde_permute_kernel(double *__restrict__ Xn, double *__restrict__ crn, const double *__restrict__ X,
        const double *__restrict__ cr, int64_t ld, int n, const int *__restrict__ idx) {
    const int pos = blockIdx.x, src = idx[pos];
    const double2 *s = reinterpret_cast<const double2*>(X + (int64_t) src * ld);
    double2 *d = reinterpret_cast<double2*>(Xn + (int64_t) pos * ld);
    for (int q = threadIdx.x; 2 * q < n; q += blockDim.x) d[q] = s[q];
    if (threadIdx.x == 0) crn[pos] = cr[src];
}

// free-running draws of candidate i (sansde.cpp:105-146): a private Philox stream, one block per request
struct SdeStream {
    uint32_t i, g0, g1, k0, k1, slot;
    __device__ Philox4 block() { return philox4x32_10(slot++, i, g0, g1 ^ 0x53440002u, k0, k1); }
    __device__ double uniform() { const Philox4 r = block(); return u53(r.x, r.y); }
    __device__ double normal() { const Philox4 r = block(); double a, b; normal_pair(r, a, b); return a; }
    __device__ int index(int count) { const Philox4 r = block(); return min(count - 1, (int) (u53(r.x, r.y) * count)); }
};

__global__ void __launch_bounds__(256) sde_draws_kernel(int np, int n, uint64_t seed, const bbgpu_sde::Params *__restrict__ par,
        double *__restrict__ zcr, int *__restrict__ ifs, double *__restrict__ F, int *__restrict__ r1, int *__restrict__ r2,
        int *__restrict__ r3, int *__restrict__ ist, int *__restrict__ jrand) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= np) return;
    const uint64_t gen = par->gen;
    const double fp = par->fp, p = par->p;
    SdeStream s { (uint32_t) i, (uint32_t) gen, (uint32_t) (gen >> 32), (uint32_t) seed, (uint32_t) (seed >> 32), 0u };
    zcr[i] = s.normal();
    const int ifstrat = s.uniform() < fp ? 0 : 1;
    double f = -1.0;
    while (f < 0.0) {
        f = ifstrat == 0 ? s.normal() * 0.3 + 0.5 : tan(3.14159265358979323846 * (s.uniform() - 0.5));
        f = fmin(f, 1.0);
    }
    int a = i, b = i, c = i;
    while (a == i) a = s.index(np);
    while (b == i || b == a) b = s.index(np);
    while (c == i || c == a || c == b) c = s.index(np);
    ifs[i] = ifstrat; F[i] = f; r1[i] = a; r2[i] = b; r3[i] = c;
    ist[i] = s.uniform() < p ? 0 : 1;
    jrand[i] = s.index(n);
}

// crossover uniforms of the generation (free-running)
__global__ void __launch_bounds__(256) sde_u_kernel(double *__restrict__ U, int64_t ld, int n, int np, uint64_t seed,
        const bbgpu_sde::Params *__restrict__ par) {
    const uint64_t gen = par->gen;
    const int64_t gid = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    const int pairs = (n + 1) >> 1;
    if (gid >= (int64_t) np * pairs) return;
    const int i = (int) (gid / pairs), jp = (int) (gid % pairs);
    const Philox4 r = philox4x32_10((uint32_t) jp, (uint32_t) i, (uint32_t) gen, (uint32_t) (gen >> 32) ^ 0x53440003u, (uint32_t) seed,
            (uint32_t) (seed >> 32));
    U[(int64_t) i * ld + 2 * jp] = u53(r.x, r.y);
    if (2 * jp + 1 < n) U[(int64_t) i * ld + 2 * jp + 1] = u53(r.z, r.w);
}

// start of a generation: CR refresh (:108-111) and the round state
__global__ void __launch_bounds__(256) sde_begin_kernel(int np, const bbgpu_sde::Params *__restrict__ par, const double *__restrict__ zcr,
        double *__restrict__ cr, int *__restrict__ succ, int *__restrict__ changed, int *__restrict__ dirty) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= np) return;
    const double crm = par->crm;
    if (par->refresh) cr[i] = fmax(0.0, fmin(__dadd_rn(__dmul_rn(zcr[i], 0.1), crm), 1.0));
    succ[i] = 0;
    changed[i] = 0;
    dirty[i] = 1;
}

// trial of candidate i (mutate() :255-286 + bound control :149-156) from the rows the sequential loop would have shown it
__global__ void __launch_bounds__(128) sde_build_kernel(double *__restrict__ T, const double *__restrict__ X, const double *__restrict__ U,
        int64_t ld, int n, const double *__restrict__ lo, const double *__restrict__ hi, const double *__restrict__ cr,
        const double *__restrict__ F, const int *__restrict__ r1, const int *__restrict__ r2, const int *__restrict__ r3,
        const int *__restrict__ ist, const int *__restrict__ jrand, const int *__restrict__ succ, const int *__restrict__ dirty,
        int *__restrict__ rebuilt, double *__restrict__ cr1, int repaircr) {
    __shared__ double red[32];
    const int i = blockIdx.x;
    if (threadIdx.x == 0) rebuilt[i] = dirty[i];
    if (!dirty[i]) return;
    auto row = [&](int p) -> const double* { return (p < i && succ[p]) ? T + (int64_t) p * ld : X + (int64_t) p * ld; };
    const double *x = X + (int64_t) i * ld, *best = row(0), *x1 = row(r1[i]), *x2 = row(r2[i]), *x3 = row(r3[i]);
    const double *u = U + (int64_t) i * ld;
    double *out = T + (int64_t) i * ld;
    const double f = F[i], CR = cr[i];
    const int strat = ist[i], jr = jrand[i];
    double count = 0.0;
    for (int q = threadIdx.x; q < n; q += blockDim.x) {
        const double xq = x[q];
        double v;
        if (strat == 0) v = __dadd_rn(x1[q], __dmul_rn(f, __dsub_rn(x2[q], x3[q])));
        else v = __dadd_rn(__dadd_rn(xq, __dmul_rn(f, __dsub_rn(best[q], xq))), __dmul_rn(f, __dsub_rn(x1[q], x2[q])));
        if (u[q] <= CR || q == jr) count += 1.0;
        else v = xq;
        if (v < lo[q]) v = __ddiv_rn(__dadd_rn(lo[q], xq), 2.0);
        else if (v > hi[q]) v = __ddiv_rn(__dadd_rn(hi[q], xq), 2.0);
        out[q] = v;
    }
    count = block_sum(count, red); // small integers: exact in any order
    if (threadIdx.x == 0) cr1[i] = repaircr ? __ddiv_rn(count, (double) n) : CR;
}

// after the evaluation of a round: outcomes, what changed, who must rebuild; one CTA
__global__ void __launch_bounds__(1024) sde_select_kernel(int np, const double *__restrict__ f, const double *__restrict__ ftrial,
        const int *__restrict__ r1, const int *__restrict__ r2, const int *__restrict__ r3, const int *__restrict__ ist, int *__restrict__ succ,
        int *__restrict__ changed, int *__restrict__ dirty, const int *__restrict__ rebuilt, int *__restrict__ ndirty) {
    __shared__ int total;
    if (threadIdx.x == 0) total = 0;
    for (int i = threadIdx.x; i < np; i += blockDim.x) {
        const int now = ftrial[i] < f[i] ? 1 : 0;
        changed[i] = (now != succ[i]) || (now && rebuilt[i]);
        succ[i] = now;
    }
    __syncthreads();
    int mine = 0;
    for (int i = threadIdx.x; i < np; i += blockDim.x) {
        const int a = r1[i], b = r2[i], c = r3[i], strat = ist[i];
        int d = (a < i && changed[a]) || (b < i && changed[b]);
        if (strat == 0) d = d || (c < i && changed[c]);
        else d = d || (0 < i && changed[0]);
        dirty[i] = d;
        mine += d;
    }
    atomicAdd(&total, mine);
    __syncthreads();
    if (threadIdx.x == 0) *ndirty = total;
}

// the adaptation sums in candidate order, as the sequential loop accumulates them (:160-176)
__global__ void sde_accumulate_kernel(int np, const double *__restrict__ f, const double *__restrict__ ftrial, const int *__restrict__ succ,
        const int *__restrict__ ist, const int *__restrict__ ifs, const double *__restrict__ F, const double *__restrict__ cr1,
        bbgpu_sde_accum *acc) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    bbgpu_sde_accum a = *acc;
    int nsucc = 0;
    for (int i = 0; i < np; i++) {
        if (succ[i]) {
            nsucc++;
            a.pns[ist[i]]++;
            const double deltaf = __dsub_rn(f[i], ftrial[i]);
            a.crrec = __dadd_rn(a.crrec, __dmul_rn(cr1[i], deltaf));
            a.crdeltaf = __dadd_rn(a.crdeltaf, deltaf);
            a.fpns[ifs[i]] = __dadd_rn(a.fpns[ifs[i]], F[i]);
        } else {
            a.pnf[ist[i]]++;
            a.fpnf[ifs[i]] = __dadd_rn(a.fpnf[ifs[i]], F[i]);
        }
    }
    a.nsucc = nsucc;
    *acc = a;
}

__global__ void __launch_bounds__(256) sde_commit_kernel(double *__restrict__ X, double *__restrict__ f, const double *__restrict__ T,
        const double *__restrict__ ftrial, const int *__restrict__ succ, int64_t ld, int n) {
    const int i = blockIdx.x;
    if (!succ[i]) return;
    const double2 *s = reinterpret_cast<const double2*>(T + (int64_t) i * ld);
    double2 *d = reinterpret_cast<double2*>(X + (int64_t) i * ld);
    for (int q = threadIdx.x; 2 * q < n; q += blockDim.x) d[q] = s[q];
    if (threadIdx.x == 0) f[i] = ftrial[i];
}

// converged() (:235-253): ||x_i||_2 per member, then mean and M2 of the radii
__global__ void __launch_bounds__(256) sde_norms_kernel(const double *__restrict__ X, int64_t ld, int n, int rows, double *__restrict__ out) {
    const int k = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (k >= rows) return;
    const double *x = X + (int64_t) k * ld;
    double s = 0.0;
    for (int i = lane; i < n; i += 32) s = fma(x[i], x[i], s);
    s = warp_sum(s);
    if (lane == 0) out[k] = sqrt(s);
}

__global__ void __launch_bounds__(1024) sde_radius_kernel(const double *__restrict__ r, int rows, double *__restrict__ out2) {
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

int sde_sort(bbgpu_sde *h) {
    bbgpu_ctx *ctx = h->ctx;
    const int c = h->cur, o = c ^ 1;
    BB_TRY(rank_fitness(ctx, h->f[c], h->np, h->sortidx, h->f[o]));
    sde_permute_kernel<<<h->np, 128, 0, ctx->stream>>>(h->X[o], h->cr[o], h->X[c], h->cr[c], h->ld, h->n, h->sortidx);
    BB_TRY(launched(ctx, "sde_permute_kernel"));
    h->cur = o;
    return BBGPU_OK;
}

} // namespace

using namespace bb;

extern "C" {

int bbgpu_sde_create(bbgpu_ctx *ctx, int n, int np, int repaircr, uint64_t seed, const double *h_lower, const double *h_upper,
        bbgpu_sde **out) {
    BB_ARG(ctx && out && h_lower && h_upper && n >= 1 && np >= 4);
    BB_CUDA(cudaSetDevice(ctx->device));
    bbgpu_sde *h = new bbgpu_sde();
    h->ctx = ctx; h->n = n; h->np = np; h->repaircr = repaircr; h->seed = seed;
    h->ld = round_up(n, 8);
    const size_t ld = (size_t) h->ld, rows = (size_t) np;
    int rc = BBGPU_OK;
    auto A = [&](double **p, size_t count) {
        if (rc != BBGPU_OK) return;
        rc = dev_alloc(p, count);
        if (rc == BBGPU_OK && cudaMemsetAsync(*p, 0, count * 8, ctx->stream) != cudaSuccess) rc = fail(BBGPU_ERR_CUDA, "memset failed");
    };
    auto I = [&](int **p, size_t count) { if (rc == BBGPU_OK) rc = dev_alloc(p, count); };
    for (int b = 0; b < 2; b++) { A(&h->X[b], rows * ld); A(&h->f[b], rows); A(&h->cr[b], rows); }
    A(&h->T, rows * ld); A(&h->U, rows * ld); A(&h->ftrial, rows); A(&h->lower, ld); A(&h->upper, ld);
    A(&h->zcr, rows); A(&h->F, rows); A(&h->cr1, rows); A(&h->sortval, rows); A(&h->rnorm, rows); A(&h->rstat, 8);
    I(&h->ifs, rows); I(&h->r1, rows); I(&h->r2, rows); I(&h->r3, rows); I(&h->ist, rows); I(&h->jrand, rows); I(&h->sortidx, rows);
    I(&h->succ, rows); I(&h->changed, rows); I(&h->dirty, rows); I(&h->rebuilt, rows); I(&h->ndirty, 8);
    if (rc == BBGPU_OK) rc = dev_alloc(&h->d_acc, 1);
    if (rc == BBGPU_OK) rc = dev_alloc(&h->d_par, 1);
    if (rc == BBGPU_OK && (cudaMallocHost((void**) &h->h_ndirty, 8 * sizeof(int)) != cudaSuccess ||
            cudaMallocHost((void**) &h->h_acc, sizeof(bbgpu_sde_accum)) != cudaSuccess ||
            cudaMallocHost((void**) &h->h_par, sizeof(bbgpu_sde::Params)) != cudaSuccess)) rc = fail(BBGPU_ERR_CUDA, "cudaMallocHost failed");
    if (rc != BBGPU_OK) { bbgpu_sde_destroy(h); return rc; }
    cudaMemcpyAsync(h->lower, h_lower, (size_t) n * 8, cudaMemcpyHostToDevice, ctx->stream);
    cudaMemcpyAsync(h->upper, h_upper, (size_t) n * 8, cudaMemcpyHostToDevice, ctx->stream);
    BB_CUDA(cudaStreamSynchronize(ctx->stream));
    *out = h;
    return BBGPU_OK;
}

int bbgpu_sde_destroy(bbgpu_sde *h) {
    if (!h) return BBGPU_OK;
    cudaStreamSynchronize(h->ctx->stream);
    double *bufs[] = { h->X[0], h->X[1], h->f[0], h->f[1], h->cr[0], h->cr[1], h->T, h->U, h->ftrial, h->lower, h->upper, h->zcr, h->F,
            h->cr1, h->sortval, h->rnorm, h->rstat };
    for (double *b : bufs) if (b) cudaFree(b);
    int *ibufs[] = { h->ifs, h->r1, h->r2, h->r3, h->ist, h->jrand, h->sortidx, h->succ, h->changed, h->dirty, h->rebuilt, h->ndirty };
    for (int *b : ibufs) if (b) cudaFree(b);
    if (h->d_acc) cudaFree(h->d_acc);
    if (h->d_par) cudaFree(h->d_par);
    if (h->evwork) cudaFree(h->evwork);
    if (h->h_par) cudaFreeHost(h->h_par);
    for (int q = 0; q < 2; q++) {
        if (h->g_head[q]) cudaGraphExecDestroy(h->g_head[q]);
        if (h->g_tail[q]) cudaGraphExecDestroy(h->g_tail[q]);
    }
    if (h->h_ndirty) cudaFreeHost(h->h_ndirty);
    if (h->h_acc) cudaFreeHost(h->h_acc);
    delete h;
    return BBGPU_OK;
}

// init(): sansde.cpp:54-92.  h_x == NULL: uniform Philox points; otherwise np*n replayed points.  Evaluates the swarm; every
// member starts with CR = 0.5.
int bbgpu_sde_init(bbgpu_sde *h, bbgpu_objective *obj, const double *h_x) {
    BB_ARG(h && obj && obj->n == h->n);
    bbgpu_ctx *ctx = h->ctx;
    cudaStream_t st = ctx->stream;
    h->cur = 0;
    h->gen = 0;
    if (h_x) {
        BB_CUDA(cudaMemcpy2DAsync(h->X[0], (size_t) h->ld * 8, h_x, (size_t) h->n * 8, (size_t) h->n * 8, h->np, cudaMemcpyHostToDevice, st));
    } else {
        const int64_t total = (int64_t) h->np * ((h->n + 1) / 2);
        sde_uniform_kernel<<<ceil_div(total, 256), 256, 0, st>>>(h->X[0], h->ld, h->n, h->np, h->lower, h->upper, h->seed);
        BB_TRY(launched(ctx, "sde_uniform_kernel"));
    }
    std::vector<double> half((size_t) h->np, 0.5);
    BB_CUDA(cudaMemcpyAsync(h->cr[0], half.data(), (size_t) h->np * 8, cudaMemcpyHostToDevice, st));
    BB_TRY(evaluate(ctx, obj, h->X[0], h->ld, h->np, h->f[0]));
    BB_CUDA(cudaStreamSynchronize(st));
    h->fev = h->np;
    return BBGPU_OK;
}

// One generation (sansde.cpp:94-177) with the reference's sequential in-place semantics.  acc: the running adaptation sums,
// in and out (the host applies the periodic updates :180-211).  draws == NULL: Philox.
//
// A free-running generation with a registry objective is the same ~25 launches every time -- only four scalars and the Philox
// counter change, and those sit in device memory -- so it is captured ONCE per parity of the swarm buffer as two CUDA graphs
// (head: sort, draws, the speculative rounds, the dirty counts; tail: adaptation sums, commit, radius statistic) and replayed:
// two graph launches and two host reads per generation instead of ~25 launches.  Replayed draws, user objectives and profiled
// runs take the launch-by-launch path (BBGPU_SDE_GRAPH=0 forces it).
} // extern "C"

namespace {

constexpr int SDE_SPEC = 3; // rounds enqueued before the host looks at the dirty counts

int sde_enqueue_round(bbgpu_sde *h, bbgpu_objective *obj, int c, int slot) {
    bbgpu_ctx *ctx = h->ctx;
    cudaStream_t st = ctx->stream;
    const int np = h->np, n = h->n;
    sde_build_kernel<<<np, 128, 0, st>>>(h->T, h->X[c], h->U, h->ld, n, h->lower, h->upper, h->cr[c], h->F, h->r1, h->r2, h->r3, h->ist,
            h->jrand, h->succ, h->dirty, h->rebuilt, h->cr1, h->repaircr);
    BB_TRY(launched(ctx, "sde_build_kernel"));
    BB_TRY(evaluate_masked(ctx, obj, h->T, h->ld, np, h->ftrial, h->rebuilt, h->evwork)); // a user objective sees the rebuilt trials only
    sde_select_kernel<<<1, 1024, 0, st>>>(np, h->f[c], h->ftrial, h->r1, h->r2, h->r3, h->ist, h->succ, h->changed, h->dirty, h->rebuilt,
            h->ndirty + slot);
    return launched(ctx, "sde_select_kernel");
}

// head of a generation: parameters, sort, draws (Philox, unless replayed ones were uploaded), round state, the speculative rounds
int sde_enqueue_head(bbgpu_sde *h, bbgpu_objective *obj, bool philox) {
    bbgpu_ctx *ctx = h->ctx;
    cudaStream_t st = ctx->stream;
    const int np = h->np, n = h->n;
    BB_CUDA(cudaMemcpyAsync(h->d_par, h->h_par, sizeof(bbgpu_sde::Params), cudaMemcpyHostToDevice, st));
    BB_TRY(sde_sort(h));
    const int c = h->cur;
    if (philox) {
        sde_draws_kernel<<<ceil_div(np, 256), 256, 0, st>>>(np, n, h->seed, h->d_par, h->zcr, h->ifs, h->F, h->r1, h->r2, h->r3, h->ist,
                h->jrand);
        BB_TRY(launched(ctx, "sde_draws_kernel"));
        const int64_t total = (int64_t) np * ((n + 1) / 2);
        sde_u_kernel<<<ceil_div(total, 256), 256, 0, st>>>(h->U, h->ld, n, np, h->seed, h->d_par);
        BB_TRY(launched(ctx, "sde_u_kernel"));
    }
    sde_begin_kernel<<<ceil_div(np, 256), 256, 0, st>>>(np, h->d_par, h->zcr, h->cr[c], h->succ, h->changed, h->dirty);
    BB_TRY(launched(ctx, "sde_begin_kernel"));
    // Rounds.  A round past the fixed point changes nothing (no candidate is dirty: the builds return at once, the evaluation
    // and the selection reproduce their outputs), so the first SDE_SPEC rounds -- what a generation usually needs -- are enqueued
    // back to back and the host reads their dirty counts in one go; further rounds, if any, one at a time.
    const int spec = obj->mode == 0 ? SDE_SPEC : 1; // a user objective is not evaluated on speculation
    for (int r = 0; r < spec; r++) BB_TRY(sde_enqueue_round(h, obj, c, r));
    BB_CUDA(cudaMemcpyAsync(h->h_ndirty, h->ndirty, spec * sizeof(int), cudaMemcpyDeviceToHost, st));
    return BBGPU_OK;
}

// tail: the adaptation sums in candidate order, the accepted trials, the radius statistic of converged()
int sde_enqueue_tail(bbgpu_sde *h) {
    bbgpu_ctx *ctx = h->ctx;
    cudaStream_t st = ctx->stream;
    const int np = h->np, n = h->n, c = h->cur;
    BB_CUDA(cudaMemcpyAsync(h->d_acc, h->h_acc, sizeof(bbgpu_sde_accum), cudaMemcpyHostToDevice, st));
    sde_accumulate_kernel<<<1, 32, 0, st>>>(np, h->f[c], h->ftrial, h->succ, h->ist, h->ifs, h->F, h->cr1, h->d_acc);
    BB_TRY(launched(ctx, "sde_accumulate_kernel"));
    sde_commit_kernel<<<np, 128, 0, st>>>(h->X[c], h->f[c], h->T, h->ftrial, h->succ, h->ld, n);
    BB_TRY(launched(ctx, "sde_commit_kernel"));
    // converged() (:235-253) reads the swarm as this generation leaves it: its statistic travels with the sums
    sde_norms_kernel<<<ceil_div(np, 8), 256, 0, st>>>(h->X[c], h->ld, n, np, h->rnorm);
    BB_TRY(launched(ctx, "sde_norms_kernel"));
    sde_radius_kernel<<<1, 1024, 0, st>>>(h->rnorm, np, &h->d_acc->rmean);
    BB_TRY(launched(ctx, "sde_radius_kernel"));
    BB_CUDA(cudaMemcpyAsync(h->h_acc, h->d_acc, sizeof(bbgpu_sde_accum), cudaMemcpyDeviceToHost, st));
    return BBGPU_OK;
}

// capture `enqueue` on the ctx stream into an executable graph; *kernels: launches it holds
template<typename F>
int sde_capture(bbgpu_sde *h, F enqueue, cudaGraphExec_t *exec, int64_t *kernels) {
    bbgpu_ctx *ctx = h->ctx;
    const int64_t before = ctx->launches;
    BB_CUDA(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
    const int rc = enqueue();
    cudaGraph_t graph = nullptr;
    const cudaError_t e = cudaStreamEndCapture(ctx->stream, &graph);
    *kernels = ctx->launches - before;
    ctx->launches = before; // nothing ran yet: the replays are counted
    if (rc != BBGPU_OK || e != cudaSuccess || graph == nullptr) {
        if (graph) cudaGraphDestroy(graph);
        return rc != BBGPU_OK ? rc : fail(BBGPU_ERR_CUDA, "stream capture failed: %s", cudaGetErrorString(e));
    }
    const cudaError_t ei = cudaGraphInstantiate(exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ei != cudaSuccess) return fail(BBGPU_ERR_CUDA, "cudaGraphInstantiate failed: %s", cudaGetErrorString(ei));
    return BBGPU_OK;
}

void sde_drop_graphs(bbgpu_sde *h) {
    for (int q = 0; q < 2; q++) {
        if (h->g_head[q]) cudaGraphExecDestroy(h->g_head[q]);
        if (h->g_tail[q]) cudaGraphExecDestroy(h->g_tail[q]);
        h->g_head[q] = h->g_tail[q] = nullptr;
    }
    h->g_obj = nullptr;
}

} // namespace

extern "C" {

int bbgpu_sde_generation(bbgpu_sde *h, bbgpu_objective *obj, int refresh_cr, double crm, double fp, double p, const bbgpu_sde_draws *draws,
        bbgpu_sde_accum *acc, int *rounds_out) {
    BB_ARG(h && obj && obj->n == h->n && acc);
    bbgpu_ctx *ctx = h->ctx;
    cudaStream_t st = ctx->stream;
    const int np = h->np, n = h->n;
    ProfScope prof(ctx, BBGPU_PROF_DE_ROUND);
    if (obj->mode != 0 && h->evwork == nullptr) BB_CUDA(cudaMalloc(&h->evwork, evaluate_masked_work_bytes(n, np)));
    static const bool graphs_on = [] { const char *e = getenv("BBGPU_SDE_GRAPH"); return !(e && e[0] == '0'); }();
    // (the first two generations run launch by launch: lazily created buffers of the ranking exist afterwards)
    const bool graph = graphs_on && draws == nullptr && obj->mode == 0 && !ctx->profiling && h->gen >= 2;
    *h->h_par = bbgpu_sde::Params { crm, fp, p, (unsigned long long) h->gen, refresh_cr, 0 };
    if (draws) {
        BB_ARG(draws->ifstrat && draws->F && draws->r1 && draws->r2 && draws->r3 && draws->istrat && draws->jrand && draws->u);
        BB_ARG(!refresh_cr || draws->zcr);
        for (int i = 0; i < np; i++) { // replayed indices address device rows and the two-entry sums: check them here
            BB_ARG(draws->r1[i] >= 0 && draws->r1[i] < np && draws->r2[i] >= 0 && draws->r2[i] < np && draws->r3[i] >= 0 && draws->r3[i] < np);
            BB_ARG((draws->ifstrat[i] | 1) == 1 && (draws->istrat[i] | 1) == 1 && draws->jrand[i] >= 0 && draws->jrand[i] < n);
        }
        if (draws->zcr) BB_CUDA(cudaMemcpyAsync(h->zcr, draws->zcr, (size_t) np * 8, cudaMemcpyHostToDevice, st));
        BB_CUDA(cudaMemcpyAsync(h->F, draws->F, (size_t) np * 8, cudaMemcpyHostToDevice, st));
        const struct { int *d; const int32_t *s; } iv[] = { { h->ifs, draws->ifstrat }, { h->r1, draws->r1 }, { h->r2, draws->r2 },
                { h->r3, draws->r3 }, { h->ist, draws->istrat }, { h->jrand, draws->jrand } };
        for (const auto &v : iv) BB_CUDA(cudaMemcpyAsync(v.d, v.s, (size_t) np * 4, cudaMemcpyHostToDevice, st));
        BB_CUDA(cudaMemcpy2DAsync(h->U, (size_t) h->ld * 8, draws->u, (size_t) n * 8, (size_t) n * 8, np, cudaMemcpyHostToDevice, st));
    }
    if (graph && (h->g_obj != obj || h->g_obj_id != obj->id || h->g_obj_aux != obj->d_aux)) sde_drop_graphs(h); // another objective: the evaluation launches differ
    const int c0 = h->cur;
    if (graph) {
        if (h->g_head[c0] == nullptr) {
            const int64_t evals0 = obj->evals;
            BB_TRY(sde_capture(h, [&] { return sde_enqueue_head(h, obj, true); }, &h->g_head[c0], &h->g_head_kernels[c0]));
            h->g_head_evals[c0] = obj->evals - evals0;
            obj->evals = evals0;
            h->g_obj = obj;
            h->g_obj_id = obj->id;
            h->g_obj_aux = obj->d_aux;
            h->cur = c0; // the capture walked through the sort
        }
        BB_CUDA(cudaGraphLaunch(h->g_head[c0], st));
        ctx->launches += h->g_head_kernels[c0];
        obj->evals += h->g_head_evals[c0];
        h->cur = c0 ^ 1;
    } else {
        BB_TRY(sde_enqueue_head(h, obj, draws == nullptr));
    }
    h->gen++;
    const int c = h->cur;
    BB_CUDA(cudaStreamSynchronize(st));
    int rounds = 0;
    bool fixed = false;
    for (int r = 0; r < (obj->mode == 0 ? SDE_SPEC : 1) && !fixed; r++) {
        rounds++;
        fixed = h->h_ndirty[r] == 0;
    }
    while (!fixed) {
        BB_TRY(sde_enqueue_round(h, obj, c, 0));
        BB_CUDA(cudaMemcpyAsync(h->h_ndirty, h->ndirty, sizeof(int), cudaMemcpyDeviceToHost, st));
        BB_CUDA(cudaStreamSynchronize(st));
        rounds++;
        fixed = h->h_ndirty[0] == 0;
        if (rounds > np + 1) return fail(BBGPU_ERR_CUDA, "SaNSDE rounds did not reach a fixed point");
    }
    *h->h_acc = *acc;
    if (graph) {
        if (h->g_tail[c] == nullptr) BB_TRY(sde_capture(h, [&] { return sde_enqueue_tail(h); }, &h->g_tail[c], &h->g_tail_kernels[c]));
        BB_CUDA(cudaGraphLaunch(h->g_tail[c], st));
        ctx->launches += h->g_tail_kernels[c];
    } else {
        BB_TRY(sde_enqueue_tail(h));
    }
    BB_CUDA(cudaStreamSynchronize(st));
    *acc = *h->h_acc;
    h->fev += np;
    if (rounds_out) *rounds_out = rounds;
    return BBGPU_OK;
}

// std::sort(_swarm) as solution() does it (:214-217)
int bbgpu_sde_sort(bbgpu_sde *h) {
    BB_ARG(h);
    return sde_sort(h);
}

// members in their current order: x (np*n), f (np), cr (np); any pointer may be NULL
int bbgpu_sde_get_swarm(bbgpu_sde *h, double *h_x, double *h_f, double *h_cr) {
    BB_ARG(h);
    cudaStream_t st = h->ctx->stream;
    const int c = h->cur;
    if (h_x) BB_CUDA(cudaMemcpy2DAsync(h_x, (size_t) h->n * 8, h->X[c], (size_t) h->ld * 8, (size_t) h->n * 8, h->np, cudaMemcpyDeviceToHost, st));
    if (h_f) BB_CUDA(cudaMemcpyAsync(h_f, h->f[c], (size_t) h->np * 8, cudaMemcpyDeviceToHost, st));
    if (h_cr) BB_CUDA(cudaMemcpyAsync(h_cr, h->cr[c], (size_t) h->np * 8, cudaMemcpyDeviceToHost, st));
    BB_CUDA(cudaStreamSynchronize(st));
    return BBGPU_OK;
}

// outcome of the last generation by candidate: success flags, trial fitness, realised crossover rates; any may be NULL
int bbgpu_sde_get_trace(bbgpu_sde *h, int32_t *h_succ, double *h_ftrial, double *h_cr1) {
    BB_ARG(h);
    cudaStream_t st = h->ctx->stream;
    if (h_succ) BB_CUDA(cudaMemcpyAsync(h_succ, h->succ, (size_t) h->np * 4, cudaMemcpyDeviceToHost, st));
    if (h_ftrial) BB_CUDA(cudaMemcpyAsync(h_ftrial, h->ftrial, (size_t) h->np * 8, cudaMemcpyDeviceToHost, st));
    if (h_cr1) BB_CUDA(cudaMemcpyAsync(h_cr1, h->cr1, (size_t) h->np * 8, cudaMemcpyDeviceToHost, st));
    BB_CUDA(cudaStreamSynchronize(st));
    return BBGPU_OK;
}

// converged() (:235-253): mean and M2 = sum (r_i - mean)^2 of the members' Euclidean norms
int bbgpu_sde_radius(bbgpu_sde *h, double *h_mean, double *h_m2) {
    BB_ARG(h && h_mean && h_m2);
    cudaStream_t st = h->ctx->stream;
    sde_norms_kernel<<<ceil_div(h->np, 8), 256, 0, st>>>(h->X[h->cur], h->ld, h->n, h->np, h->rnorm);
    BB_TRY(launched(h->ctx, "sde_norms_kernel"));
    sde_radius_kernel<<<1, 1024, 0, st>>>(h->rnorm, h->np, h->rstat);
    BB_TRY(launched(h->ctx, "sde_radius_kernel"));
    double out[2];
    BB_CUDA(cudaMemcpyAsync(out, h->rstat, 16, cudaMemcpyDeviceToHost, st));
    BB_CUDA(cudaStreamSynchronize(st));
    *h_mean = out[0];
    *h_m2 = out[1];
    return BBGPU_OK;
}

int bbgpu_sde_evaluations(bbgpu_sde *h) { return h ? h->fev : 0; }

} // extern "C"
