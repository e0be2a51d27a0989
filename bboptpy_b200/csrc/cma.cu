// CMA-ES device engine: state + the per-generation phases behind the C ABI (include/bbgpu.h).
//
// Reference being replaced (paths under /root/reference/src/multivariate/cma):
//   sample      cmaes.cpp:65-80, sep_cmaes.cpp:69-79
//   evaluate    base_cmaes.cpp:211-219            rank   base_cmaes.cpp:221-229
//   update      cmaes.cpp:82-142, active_cmaes.cpp:71-161, sep_cmaes.cpp:81-131
//   eig         cmaes.cpp:229-283 (tred2/tql2 -> cuSOLVER syevd, or a replayed basis)
//   probe       cmaes.cpp:184-229, sep_cmaes.cpp:170-205 (the O(n) parts of converged())
//
// HBM layout (fp64, row-major, row pitch ldn = n rounded up to 8 doubles = 64 B; pad columns
// stay zero):  X, Z [lambda x ldn] candidate-major like the reference's _arx;  C, Bt, BDt,
// BDit, isC [n x ldn];  Ysel [2mu x ldn] the selected, centred, sigma-scaled steps that feed the
// covariance SYRK.  Bt is the TRANSPOSE of the reference's _b (row j = eigenvector j): that is what
// syevd leaves in memory and what both tensor-core operand forms want (k-major rows).
#include <cusolverDn.h>

#include <algorithm>
#include <cmath>

#include "chol.cuh"
#include "lmcma.cuh"
#include "xnes.cuh"
#include "gemm_tma.cuh"
#include "objectives.cuh"
#include "philox.cuh"
#include "rank.cuh"

namespace bb { struct PeerRegion; }

struct bbgpu_cma {
    bbgpu_ctx *ctx = nullptr;
    bbgpu_cma_config cfg;
    int n = 0, lambda = 0, mu = 0, nsel = 0;
    int64_t ldn = 0;
    bool full = false, active = false, chol = false;
    // Cholesky-CMA-ES (chol.cuh): stacked SYRK operand [A^T ; Y], the matrix being factored, the panel transposed, -1 coefficients,
    // candidate radii + their statistics
    double *Ystk = nullptr, *Mch = nullptr, *Lt = nullptr, *negone = nullptr, *rnorm = nullptr, *rstat = nullptr;
    cudaEvent_t ev_ch_fork = nullptr, ev_ch_join = nullptr;
    // LM-CMA-ES (lmcma.cuh): direction vectors and their scalars, the base draws of the pairs, live-row order, subset starts
    bool xnes = false; // xNES (xnes.cuh): Bt = the transform B (transposed), C = G / its scaled form, BDt / BDit / isC scratch
    bool lm = false;
    int lm_mem = 0, lm_pairs = 0, lm_slabs = 1, lm_part_ld = 0;
    double *lm_part = nullptr;    // partial inner products of the slab-split sampling: pairs x slabs x lm_part_ld
    int *lm_flags = nullptr;      // rebuild wavefront: one flag per memory row + a give-up word
    double *lm_tglob = nullptr;   // rebuild scratch (one row per CTA) when a row does not fit the registers of a CTA
    double *PCmat = nullptr, *Vmat = nullptr, *lmb = nullptr, *lmd = nullptr, *Zp = nullptr;
    int *lm_jarr = nullptr, *lm_i0 = nullptr;
    // vectors
    double *xmean = nullptr, *xold = nullptr, *pc = nullptr, *ps = nullptr, *D = nullptr, *weights = nullptr,
            *lower = nullptr, *upper = nullptr, *dm = nullptr, *W = nullptr, *coef = nullptr, *ycoeff = nullptr,
            *ssq = nullptr, *dscal = nullptr, *tmpn = nullptr;
    // matrices
    double *C = nullptr, *Bt = nullptr, *BDt = nullptr, *BDit = nullptr, *isC = nullptr;
    double *X = nullptr, *Z = nullptr, *Ysel = nullptr, *fit = nullptr, *rank_val = nullptr;
    int *rank_idx = nullptr;
    // reductions
    double *part = nullptr; // [splits][2][n] partial sums for mean / sep variance
    int splits = 1;
    double *rowsq_part = nullptr; // [tiles_n][mu]
    int rowsq_tiles_cap = 0;
    // eigensolver workspace
    double *solver_work = nullptr;
    int solver_lwork = 0;
    double *own_ws = nullptr;     // own eigensolver: d, e, tau, reflectors (n*n), sytrd workspace, ormtr workspace
    int own_lwork = 0;
    int *own_info = nullptr;
    cudaEvent_t ev_eig_fork = nullptr, ev_eig_join = nullptr;
    int *dev_info = nullptr;
    // draw prefetch: Z is double buffered; the draws of generation g+1 are generated on the ctx's side
    // stream while generation g runs its update (counter-based Philox: they depend on nothing)
    double *Zbuf[2] = { nullptr, nullptr };
    int zcur = 0;
    long long zgen[2] = { -1, -1 };      // generation whose draws a buffer holds (or is being filled with)
    cudaEvent_t ev_z[2] = { nullptr, nullptr };    // buffer filled (recorded on aux)
    cudaEvent_t ev_used[2] = { nullptr, nullptr }; // buffer consumed by sample() (recorded on the main stream)
    bool used_rec[2] = { false, false };
    // candidate sharding of one run over several GPUs: this handle owns candidates
    // [row0, row0 + lam_loc); X / Z hold only those rows, fitness and ranking cover all lambda
    int shards = 1, shard = 0, lam_loc = 0, row0 = 0;
    bbgpu_comm *comm = nullptr;
    bb::PeerRegion *peer = nullptr; // the rank-mu partial buffers of all ranks, mapped (NVLink peer memory); null: NCCL
    double last_sigma = 0.0;
    double *Yloc = nullptr, *coef_loc = nullptr, *wm_loc = nullptr, *norm_all = nullptr, *red = nullptr, *rowsq_loc = nullptr;
    double *Ycomp = nullptr, *coef_comp = nullptr; // the selected local steps, compacted (non-active kind: half the rows carry no weight)
    int *sel_idx = nullptr, *ksel = nullptr;
    int rowsq_loc_tiles = 0;
    // report
    bbgpu_cma_report *d_report = nullptr;
    bbgpu_cma_report *h_report = nullptr; // pinned
};

namespace bb {

// comm.cu
int comm_allreduce_sum(bbgpu_comm *c, double *d_buf, size_t count);
int comm_allgather(bbgpu_comm *c, double *d_buf, size_t per);
int comm_allgather2(bbgpu_comm *c, double *d_a, double *d_b, size_t per);
int comm_broadcast(bbgpu_comm *c, double *d_buf, size_t count, int root);
// sytrd.cu / stedc.cu: our own dense symmetric eigensolver (tridiagonalisation + divide & conquer)
bool sytrd_fits(bbgpu_ctx *ctx, int n);
size_t sytrd_workspace_doubles(int n);
int q_rows_device(bbgpu_ctx *ctx, cudaStream_t st, int n, const double *d_Vcm, const double *d_tau, double *d_R, int ldr);
int sytrd_device(bbgpu_ctx *ctx, int n, const double *d_A, int ldA, double *d_d, double *d_e, double *d_tau, double *d_Vcm, double *ws);
int stedc_device(bbgpu_ctx *ctx, int n, const double *d, const double *e, double *W, double *Zt_out, int ldz, int *d_info);
struct PeerRegion;
int comm_peer_region_create(bbgpu_comm *c, double *d_local, size_t bytes, PeerRegion **out);
int comm_peer_region_destroy(PeerRegion *r);
int comm_allreduce_sum_p2p(PeerRegion *r, size_t count);
int comm_rank(const bbgpu_comm *c);
int comm_world(const bbgpu_comm *c);
bbgpu_ctx *comm_ctx(const bbgpu_comm *c);

void destroy_solver(bbgpu_ctx *ctx) {
    if (ctx->solver) {
        cusolverDnDestroy((cusolverDnHandle_t) ctx->solver);
        ctx->solver = nullptr;
    }
}

static int get_solver(bbgpu_ctx *ctx, cusolverDnHandle_t *out) {
    if (!ctx->solver) {
        cusolverDnHandle_t h;
        if (cusolverDnCreate(&h) != CUSOLVER_STATUS_SUCCESS) return fail(BBGPU_ERR_SOLVER, "cusolverDnCreate failed");
        if (cusolverDnSetStream(h, ctx->stream) != CUSOLVER_STATUS_SUCCESS) {
            cusolverDnDestroy(h);
            return fail(BBGPU_ERR_SOLVER, "cusolverDnSetStream failed");
        }
        ctx->solver = (cusolverDnContext*) h;
    }
    *out = (cusolverDnHandle_t) ctx->solver;
    return BBGPU_OK;
}

// ------------------------------------------------------------------------------------------
// sampling
// ------------------------------------------------------------------------------------------

// Z[k][2jp], Z[k][2jp+1] <- Box-Muller pair of Philox block (jp, k0+k, generation)
__global__ void __launch_bounds__(256) philox_normal_kernel(double *__restrict__ Z, int64_t ld, int rows, int cols,
        int row0, uint64_t seed, uint64_t generation, uint32_t stream_tag) {
    const int pairs = (cols + 1) >> 1;
    const int64_t gid = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (int64_t) rows * pairs) return;
    const int k = (int) (gid / pairs), jp = (int) (gid % pairs);
    const Philox4 r = philox4x32_10((uint32_t) jp, (uint32_t) (row0 + k), (uint32_t) generation,
            (uint32_t) (generation >> 32) ^ stream_tag, (uint32_t) seed, (uint32_t) (seed >> 32));
    double z0, z1;
    normal_pair(r, z0, z1);
    double *dst = Z + (int64_t) k * ld + 2 * jp;
    dst[0] = z0;
    if (2 * jp + 1 < cols) dst[1] = z1;
}

// SepCmaes::samplePopulation (sep_cmaes.cpp:69-79): x = m + sigma d z, clipped. HBM-bound.
__global__ void __launch_bounds__(256) sep_sample_kernel(double *__restrict__ X, const double *__restrict__ Z,
        int64_t ld, int rows, int n, const double *__restrict__ m, const double *__restrict__ d, double sigma,
        const double *__restrict__ lo, const double *__restrict__ hi) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double mi = m[i], sd = __dmul_rn(sigma, d[i]);
    const bool clip = lo != nullptr;
    const double l = clip ? lo[i] : 0.0, h = clip ? hi[i] : 0.0;
    for (int k = blockIdx.y; k < rows; k += gridDim.y) {
        double v = __dadd_rn(mi, __dmul_rn(sd, Z[(int64_t) k * ld + i]));
        if (clip) v = fmax(l, fmin(v, h));
        X[(int64_t) k * ld + i] = v;
    }
}

// SepCmaes with free-running draws, the whole sample -> evaluate front of the generation in ONE pass:
// Philox block -> Box-Muller pair -> x = m + sigma d z (clipped) -> X written once (16-byte stores) -> objective.
// One CTA (8 warps) per candidate: the warps share the Philox / Box-Muller work of a 1024-column tile (lane l of a warp
// owns columns 2l, 2l+1 of a 64-column block) and stage the objective terms in shared memory; warp 0 then adds them in
// the objective's fixed order (lane a: term(i) for i = a, a + 32, ... ascending; objectives.cuh) while the other warps
// already work on the next tile, so the fitness bits equal eval_rows_kernel's on the same X.  HBM traffic: 8 n bytes
// per candidate, the algorithmic minimum (the draws never touch memory).  Objectives whose terms read a neighbour
// coordinate (Rosenbrock) take the unfused path.
constexpr int SEP_TILE = 1024;

template<int ID>
__global__ void __launch_bounds__(256) sep_sample_eval_kernel(double *__restrict__ X, int64_t ld, int rows, int n, int row0,
        const double *__restrict__ m, const double *__restrict__ d, double sigma, const double *__restrict__ lo,
        const double *__restrict__ hi, const double *__restrict__ aux, uint64_t seed, uint64_t generation,
        double *__restrict__ f) {
    constexpr int NB = ID == BBGPU_OBJ_ACKLEY ? 2 : 1;
    __shared__ __align__(16) double T[2][NB][SEP_TILE];
    const int k = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool clip = lo != nullptr;
    double *xrow = X + (int64_t) k * ld;
    double acc = 0.0, acc2 = 0.0;
    int buf = 0;
    for (int tile0 = 0; tile0 < n; tile0 += SEP_TILE, buf ^= 1) {
        const int tl = min(SEP_TILE, n - tile0);
        for (int j = warp; 64 * j < tl; j += 16) {
            // two 64-column blocks (j and j + 8) per trip: two independent Philox -> Box-Muller chains in flight per lane
            int c[2];
            bool ok[2];
            double z[2][2];
#pragma unroll
            for (int u = 0; u < 2; u++) {
                c[u] = tile0 + 64 * (j + 8 * u) + 2 * lane;
                ok[u] = 64 * (j + 8 * u) < tl && c[u] < n;
                const Philox4 r = philox4x32_10((uint32_t) (c[u] >> 1), (uint32_t) (row0 + k), (uint32_t) generation,
                        (uint32_t) (generation >> 32), (uint32_t) seed, (uint32_t) (seed >> 32));
                normal_pair(r, z[u][0], z[u][1]);
            }
#pragma unroll
            for (int u = 0; u < 2; u++) {
                if (!ok[u]) continue;
                const int c0 = c[u];
                const bool two = c0 + 1 < n;
                double x0 = __dadd_rn(m[c0], __dmul_rn(__dmul_rn(sigma, d[c0]), z[u][0]));
                double x1 = two ? __dadd_rn(m[c0 + 1], __dmul_rn(__dmul_rn(sigma, d[c0 + 1]), z[u][1])) : 0.0;
                if (clip) {
                    x0 = fmax(lo[c0], fmin(x0, hi[c0]));
                    if (two) x1 = fmax(lo[c0 + 1], fmin(x1, hi[c0 + 1]));
                }
                if (two) *reinterpret_cast<double2*>(xrow + c0) = make_double2(x0, x1); // c0 even, ld multiple of 8
                else xrow[c0] = x0;
                double e0 = 0.0, e1 = 0.0;
                const double t0 = obj_term(ID, x0, 0.0, ID == BBGPU_OBJ_ELLIPSOID ? aux[c0] : 0.0, e0);
                const double t1 = two ? obj_term(ID, x1, 0.0, ID == BBGPU_OBJ_ELLIPSOID ? aux[c0 + 1] : 0.0, e1) : 0.0;
                *reinterpret_cast<double2*>(&T[buf][0][c0 - tile0]) = make_double2(t0, t1);
                if (ID == BBGPU_OBJ_ACKLEY) *reinterpret_cast<double2*>(&T[buf][NB - 1][c0 - tile0]) = make_double2(e0, e1);
            }
        }
        __syncthreads();
        if (warp == 0) { // tile0 is a multiple of 32: lane a keeps adding the terms of its residue class, ascending
            for (int q = lane; q < tl; q += 32) {
                acc = add(acc, T[buf][0][q]);
                if (ID == BBGPU_OBJ_ACKLEY) acc2 = add(acc2, T[buf][NB - 1][q]);
            }
        }
        // the buffer written two tiles from now is this one: warp 0 is past the next barrier by then
    }
    if (warp == 0) {
        const double s1 = warp_sum(acc);
        const double s2 = ID == BBGPU_OBJ_ACKLEY ? warp_sum(acc2) : 0.0;
        if (lane == 0) f[k] = obj_finish(ID, n, s1, s2);
    }
}

// BDt[j][i] = D[j] Bt[j][i],  BDit[j][i] = Bt[j][i] / D[j]
__global__ void __launch_bounds__(256) scale_basis_kernel(const double *__restrict__ Bt, const double *__restrict__ D,
        double *__restrict__ BDt, double *__restrict__ BDit, int n, int64_t ld) {
    const int j = blockIdx.x;
    const double d = D[j];
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const double b = Bt[(int64_t) j * ld + i];
        BDt[(int64_t) j * ld + i] = d * b;
        BDit[(int64_t) j * ld + i] = b / d;
    }
}

// ------------------------------------------------------------------------------------------
// ranking report
// ------------------------------------------------------------------------------------------
__global__ void rank_report_kernel(const int *__restrict__ idx, const double *__restrict__ val, int lambda, int ik,
        bbgpu_cma_report *rep) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const int pos[4] = { 0, lambda > 1 ? 1 : 0, lambda > 1 ? lambda - 2 : 0, lambda - 1 };
    for (int q = 0; q < 4; q++) {
        rep->ibw[q] = idx[pos[q]];
        rep->ybw[q] = val[pos[q]];
    }
    rep->f_kth = val[ik < lambda ? ik : lambda - 1];
}

// ------------------------------------------------------------------------------------------
// update: selection gather, recombination, evolution paths
// ------------------------------------------------------------------------------------------

// Ysel[r][:] = (X[src(r)][:] - m_old) / sigma.  r < mu: r-th best.  r >= mu (active): (r-mu)-th worst.
__global__ void __launch_bounds__(256) gather_steps_kernel(double *__restrict__ Ysel, const double *__restrict__ X,
        int64_t ld, int n, int mu, int lambda, const int *__restrict__ rank_idx, const double *__restrict__ mold,
        double sigma) {
    const int r = blockIdx.x;
    const int src = r < mu ? rank_idx[r] : rank_idx[lambda - 1 - (r - mu)];
    const double *x = X + (int64_t) src * ld;
    double *y = Ysel + (int64_t) r * ld;
    const double rs = __drcp_rn(sigma);
    if (div_by_ok(sigma)) {
        for (int i = threadIdx.x; i < n; i += blockDim.x) y[i] = div_by(__dsub_rn(x[i], mold[i]), sigma, rs);
    } else {
        for (int i = threadIdx.x; i < n; i += blockDim.x) y[i] = __ddiv_rn(__dsub_rn(x[i], mold[i]), sigma);
    }
}

// Full-covariance kinds, one pass over the selected rows: writes the centred, sigma-scaled steps
// Ysel[r][:] = (X[src(r)][:] - m_old)/sigma (the SYRK operand; r < mu: r-th best, r >= mu: (r-mu)-th worst,
// Active only) and the recombination partials  part[s][i] = sum_{k in split s, k < mu} w_k X[idx_k][i].
__global__ void __launch_bounds__(256) select_steps_kernel(double *__restrict__ part, double *__restrict__ Ysel,
        const double *__restrict__ X, int64_t ld, int n, int mu, int nsel, int lambda, int splits,
        const int *__restrict__ rank_idx, const double *__restrict__ w, const double *__restrict__ mold, double sigma) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int s = blockIdx.y;
    if (i >= n) return;
    const int per = (nsel + splits - 1) / splits;
    const int r0 = s * per, r1 = min(nsel, r0 + per);
    const double mo = mold[i];
    double am = 0.0;
    const double rs = __drcp_rn(sigma);
    if (div_by_ok(sigma)) { // the common case, without a division in the loop
#pragma unroll 4
        for (int r = r0; r < r1; r++) {
            const int src = r < mu ? rank_idx[r] : rank_idx[lambda - 1 - (r - mu)];
            const double x = X[(int64_t) src * ld + i];
            Ysel[(int64_t) r * ld + i] = div_by(__dsub_rn(x, mo), sigma, rs);
            if (r < mu) am = __dadd_rn(am, __dmul_rn(w[r], x));
        }
    } else {
        for (int r = r0; r < r1; r++) {
            const int src = r < mu ? rank_idx[r] : rank_idx[lambda - 1 - (r - mu)];
            const double x = X[(int64_t) src * ld + i];
            Ysel[(int64_t) r * ld + i] = __ddiv_rn(__dsub_rn(x, mo), sigma);
            if (r < mu) am = __dadd_rn(am, __dmul_rn(w[r], x));
        }
    }
    part[((int64_t) s * 2 + 0) * n + i] = am;
}

// partial[s][i] = sum over rows k of split s of w_k X[idx_k][i]  (+ sep: w_k ((x - m_old)/sigma)^2)
template<bool SEP>
__global__ void __launch_bounds__(256) recombine_partial_kernel(double *__restrict__ part, const double *__restrict__ X,
        int64_t ld, int n, int mu, int splits, const int *__restrict__ rank_idx, const double *__restrict__ w,
        const double *__restrict__ mold, double sigma) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int s = blockIdx.y;
    if (i >= n) return;
    const int per = (mu + splits - 1) / splits;
    const int k0 = s * per, k1 = min(mu, k0 + per);
    double am = 0.0, ac = 0.0;
    const double mo = SEP ? mold[i] : 0.0;
    const double rs = __drcp_rn(sigma);
    if (!SEP || div_by_ok(sigma)) {
#pragma unroll 8
        for (int k = k0; k < k1; k++) {
            const double x = X[(int64_t) rank_idx[k] * ld + i];
            const double wk = w[k];
            am = __dadd_rn(am, __dmul_rn(wk, x));
            if (SEP) {
                // the reference's (x - xold) / sigma, bit for bit (long replays stay aligned), without the generic division
                const double y = div_by(__dsub_rn(x, mo), sigma, rs);
                ac = __dadd_rn(ac, __dmul_rn(wk, __dmul_rn(y, y)));
            }
        }
    } else {
        for (int k = k0; k < k1; k++) {
            const double x = X[(int64_t) rank_idx[k] * ld + i];
            const double wk = w[k];
            am = __dadd_rn(am, __dmul_rn(wk, x));
            const double y = __ddiv_rn(__dsub_rn(x, mo), sigma);
            ac = __dadd_rn(ac, __dmul_rn(wk, __dmul_rn(y, y)));
        }
    }
    part[((int64_t) s * 2 + 0) * n + i] = am;
    if (SEP) part[((int64_t) s * 2 + 1) * n + i] = ac;
}

// m_new = sum_s partial (active: m_old (1-cm) + sum cm), clip, dm = (m_new - m_old)/(cm sigma).
// sep additionally: ps_i = (1-cs) ps_i + csc C_i dm_i  (sep_cmaes.cpp:99-102: the VARIANCE c_i, kept)
// and tmpn_i = sum_s partial variance term.
template<bool SEP>
__global__ void __launch_bounds__(256) recombine_final_kernel(const double *__restrict__ part, int n, int splits,
        double *__restrict__ xmean, const double *__restrict__ xold, double *__restrict__ dm, double cm, double sigma,
        const double *__restrict__ lo, const double *__restrict__ hi, bool active, double *__restrict__ ps,
        const double *__restrict__ Cdiag, double cs, double csc, double *__restrict__ tmpn) {
    // 8 columns per warp: lanes = 4 split groups x 8 columns, so a warp reads 64-byte row segments of `part`
    const int lane = threadIdx.x & 31, warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int i = warp * 8 + (lane & 7), sg = lane >> 3;
    double sm = 0.0, sc = 0.0;
    if (i < n) {
        for (int s = sg; s < splits; s += 4) {
            sm = __dadd_rn(sm, part[((int64_t) s * 2 + 0) * n + i]);
            if (SEP) sc = __dadd_rn(sc, part[((int64_t) s * 2 + 1) * n + i]);
        }
    }
    sm = __dadd_rn(sm, __shfl_xor_sync(0xffffffffu, sm, 8));
    sm = __dadd_rn(sm, __shfl_xor_sync(0xffffffffu, sm, 16));
    if (SEP) {
        sc = __dadd_rn(sc, __shfl_xor_sync(0xffffffffu, sc, 8));
        sc = __dadd_rn(sc, __shfl_xor_sync(0xffffffffu, sc, 16));
    }
    if (i >= n || sg != 0) return;
    const double mo = xold[i];
    double mn = active ? __dadd_rn(__dmul_rn(mo, 1.0 - cm), __dmul_rn(sm, cm)) : sm;
    if (lo != nullptr) mn = fmax(lo[i], fmin(mn, hi[i]));
    xmean[i] = mn;
    const double step = __ddiv_rn(__dsub_rn(mn, mo), active ? __dmul_rn(cm, sigma) : sigma);
    dm[i] = step;
    if (SEP) {
        ps[i] = __dadd_rn(__dmul_rn(ps[i], 1.0 - cs), __dmul_rn(__dmul_rn(csc, Cdiag[i]), step));
        tmpn[i] = sc;
    }
}

// p_sigma <- (1-cs) p_sigma + csc C^-1/2 dm   (cmaes.cpp:99-105); 4 warps per row, 2 rows per CTA
__global__ void __launch_bounds__(256) ps_gemv_kernel(const double *__restrict__ isC, int64_t ld, int n,
        const double *__restrict__ dm, double *__restrict__ ps, double cs, double csc) {
    __shared__ double red[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int row = blockIdx.x * 2 + (warp >> 2), q = warp & 3;
    double acc = 0.0;
    if (row < n) {
        const double *a = isC + (int64_t) row * ld;
        for (int j = q * 32 + lane; j < n; j += 128) acc = fma(a[j], dm[j], acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) red[warp] = acc;
    __syncthreads();
    if (threadIdx.x < 2) {
        const int r = blockIdx.x * 2 + threadIdx.x;
        if (r < n) {
            const double t = ((red[4 * threadIdx.x] + red[4 * threadIdx.x + 1]) + red[4 * threadIdx.x + 2]) + red[4 * threadIdx.x + 3];
            ps[r] = __dadd_rn(__dmul_rn(ps[r], 1.0 - cs), __dmul_rn(csc, t));
        }
    }
}

// ||p_sigma||, hsig, p_c (cmaes.cpp:107-122); sep also finishes c_i, d_i (sep_cmaes.cpp:118-131).
// denom = 1 - (1-cs)^(2 fev/lambda) is computed by the host with libm pow, as the reference does.
template<bool SEP>
__global__ void __launch_bounds__(1024) paths_kernel(int n, const double *__restrict__ ps, double *__restrict__ pc,
        const double *__restrict__ dm, double denom, double chi, double cc, double ccc, double *__restrict__ dscal,
        bbgpu_cma_report *rep, double *__restrict__ Cdiag, double *__restrict__ D, const double *__restrict__ tmpn,
        double ccov, double mueff, double *__restrict__ coef, const double *__restrict__ w, int mu, double cmu) {
    __shared__ double red[32];
    // non-active full covariance: SYRK coefficients cmu * w_k (cmaes.cpp:139)
    if (coef != nullptr)
        for (int k = threadIdx.x; k < mu; k += blockDim.x) coef[k] = __dmul_rn(cmu, w[k]);
    double s = 0.0;
#pragma unroll 4
    for (int i = threadIdx.x; i < n; i += blockDim.x) s = fma(ps[i], ps[i], s);
    s = block_sum(s, red);
    const double pslen = sqrt(s);
    const int hsig = (pslen / sqrt(denom) / chi < 1.4 + 2.0 / (n + 1.0)) ? 1 : 0;
    if (threadIdx.x == 0) {
        rep->pslen = pslen;
        rep->hsig = hsig;
        *dscal = (1.0 - hsig) * cc * (2.0 - cc);
    }
#pragma unroll 4
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const double p = __dadd_rn(__dmul_rn(1.0 - cc, pc[i]), __dmul_rn(__dmul_rn((double) hsig, ccc), dm[i]));
        pc[i] = p;
        if (SEP) {
            double c = __dadd_rn(__dmul_rn(1.0 - ccov, Cdiag[i]), __dmul_rn(__dmul_rn(ccov / mueff, p), p));
            c = __dadd_rn(c, __dmul_rn(ccov * (1.0 - 1.0 / mueff), tmpn[i]));
            Cdiag[i] = c;
            D[i] = sqrt(c);
        }
    }
}

// active: ycoeff and the signed SYRK coefficients (active_cmaes.cpp:115-132,146-158)
__global__ void __launch_bounds__(256) active_coef_kernel(const double *__restrict__ rowsq_part, int tiles_n, int mu,
        double sigma, const double *__restrict__ w, double cmu1, double cneg, double *__restrict__ ycoeff,
        double *__restrict__ coef) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= mu) return;
    // Ysel row mu+q is the q-th worst candidate; rowsq holds ||C^-1/2 y||^2 with y = (x-m_old)/sigma
    auto ssq = [&](int q) {
        double s = 0.0;
        for (int tnl = 0; tnl < tiles_n; tnl++) s = __dadd_rn(s, rowsq_part[(int64_t) tnl * mu + q]);
        return __dmul_rn(__dmul_rn(sigma, sigma), s);
    };
    const double top = ssq(mu - 1 - i);   // rank lambda-mu+i
    const double bot = fmax(ssq(i), 1e-8);// rank lambda-1-i
    const double yc = __ddiv_rn(top, bot);
    ycoeff[i] = yc;
    coef[i] = __dmul_rn(cmu1, w[i]);
    coef[mu + i] = -__dmul_rn(__dmul_rn(cneg, w[i]), yc);
}

__global__ void __launch_bounds__(256) plain_coef_kernel(int mu, const double *__restrict__ w, double cmu,
        double *__restrict__ coef) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < mu) coef[i] = __dmul_rn(cmu, w[i]);
}

// ------------------------------------------------------------------------------------------
// candidate-sharded update (one run over several GPUs)
// ------------------------------------------------------------------------------------------

// Y[k][:] = (X[k][:] - m)/sigma for every local candidate (the SYRK / norm operand)
__global__ void __launch_bounds__(256) shard_steps_kernel(double *__restrict__ Y, const double *__restrict__ X, int64_t ld,
        int n, int rows, const double *__restrict__ m, double sigma) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double mo = m[i];
    for (int k = blockIdx.y; k < rows; k += gridDim.y)
        Y[(int64_t) k * ld + i] = __ddiv_rn(__dsub_rn(X[(int64_t) k * ld + i], mo), sigma);
}

// norm_all[row0 + k] = sigma^2 * sum over N tiles of the row-sum-of-squares partials = ||C^-1/2 (x_k - m)||^2
__global__ void __launch_bounds__(256) shard_norms_kernel(const double *__restrict__ part, int tiles_n, int rows, double sigma,
        double *__restrict__ out) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= rows) return;
    double s = 0.0;
    for (int t = 0; t < tiles_n; t++) s = __dadd_rn(s, part[(int64_t) t * rows + k]);
    out[k] = __dmul_rn(__dmul_rn(sigma, sigma), s);
}

// From the (replicated) ranking: recombination weight and covariance coefficient of every LOCAL candidate
// (cmaes.cpp:139; active_cmaes.cpp:115-132,146-158).  One thread per sorted position r.
__global__ void __launch_bounds__(256) shard_coef_kernel(const int *__restrict__ rank_idx, int lambda, int mu, int row0,
        int lam_loc, const double *__restrict__ w, bool active, double cpos, double cneg, const double *__restrict__ norm_all,
        double *__restrict__ coef_loc, double *__restrict__ wm_loc, double *__restrict__ ycoeff) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= lambda) return;
    double yc = 0.0;
    const bool worst = active && r >= lambda - mu;
    const int i = lambda - 1 - r; // i-th worst
    if (worst) {
        const double top = norm_all[rank_idx[lambda - mu + i]];
        const double bot = fmax(norm_all[rank_idx[lambda - 1 - i]], 1e-8);
        yc = __ddiv_rn(top, bot);
        ycoeff[i] = yc;
    }
    const int loc = rank_idx[r] - row0;
    if (loc < 0 || loc >= lam_loc) return;
    double c = 0.0, wm = 0.0;
    if (r < mu) { wm = w[r]; c = __dmul_rn(cpos, w[r]); }
    else if (worst) c = -__dmul_rn(__dmul_rn(cneg, w[i]), yc);
    coef_loc[loc] = c;
    wm_loc[loc] = wm;
}

// Compaction of the local selection (non-active kind): candidates with a non-zero covariance coefficient, in local order,
// get consecutive positions; their coefficients follow, the tail of the coefficient array is zeroed, the count goes to
// *ksel.  One CTA: an inclusive scan over at most 64 K flags in chunks of 1024.
__global__ void __launch_bounds__(1024) shard_compact_kernel(const double *__restrict__ coef_loc, int rows, int *__restrict__ sel_idx,
        double *__restrict__ coef_comp, int coef_cap, int *__restrict__ ksel) {
    __shared__ int wsum[32];
    __shared__ int base_s;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) base_s = 0;
    __syncthreads();
    for (int k0 = 0; k0 < rows; k0 += 1024) {
        const int k = k0 + tid;
        const double c = k < rows ? coef_loc[k] : 0.0;
        const int flag = c != 0.0 ? 1 : 0;
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
        const int pos = base_s + (wid ? wsum[wid - 1] : 0) + x - flag;
        if (flag) {
            sel_idx[pos] = k;
            coef_comp[pos] = c;
        }
        __syncthreads();
        if (tid == 0) base_s += wsum[31];
        __syncthreads();
    }
    const int total = base_s;
    for (int q = total + tid; q < coef_cap; q += 1024) coef_comp[q] = 0.0;
    if (tid == 0) *ksel = total;
}

// Ycomp[r][:] = Y[sel_idx[r]][:] for r < *ksel
__global__ void __launch_bounds__(256) shard_gather_kernel(double *__restrict__ Ycomp, const double *__restrict__ Y, int64_t ld, int n,
        const int *__restrict__ sel_idx, const int *__restrict__ ksel) {
    const int r = blockIdx.x;
    if (r >= *ksel) return;
    const double2 *src = reinterpret_cast<const double2*>(Y + (int64_t) sel_idx[r] * ld);
    double2 *dst = reinterpret_cast<double2*>(Ycomp + (int64_t) r * ld);
    for (int i = threadIdx.x; 2 * i < n; i += blockDim.x) dst[i] = src[i]; // ld is a multiple of 8: pad columns copy as zeros
}

// part[s][i] = sum over the local rows k of split s of wm[k] X[k][i]
__global__ void __launch_bounds__(256) shard_mean_partial_kernel(double *__restrict__ part, const double *__restrict__ X,
        int64_t ld, int n, int rows, int splits, const double *__restrict__ wm) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int s = blockIdx.y;
    if (i >= n) return;
    const int per = (rows + splits - 1) / splits;
    const int k0 = s * per, k1 = min(rows, k0 + per);
    double am = 0.0;
    for (int k = k0; k < k1; k++) {
        const double wk = wm[k];
        if (wk != 0.0) am = __dadd_rn(am, __dmul_rn(wk, X[(int64_t) k * ld + i]));
    }
    part[(int64_t) s * n + i] = am;
}

__global__ void __launch_bounds__(256) shard_mean_final_kernel(const double *__restrict__ part, int n, int splits,
        double *__restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double sm = 0.0;
    for (int s = 0; s < splits; s++) sm = __dadd_rn(sm, part[(int64_t) s * n + i]);
    out[i] = sm;
}

// C (lower) <- decay C + c1 (pc pc^T + c2 C) + S, S = the all-reduced sum of the per-GPU rank-mu partials
__global__ void __launch_bounds__(256) shard_cov_finish_kernel(double *__restrict__ C, const double *__restrict__ S, int64_t ld,
        int n, const double *__restrict__ pc, double decay, double c1, const double *__restrict__ dscal) {
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    const int row = blockIdx.y;
    if (col > row || col >= n) return;
    const double c2 = *dscal;
    const int64_t o = (int64_t) row * ld + col;
    const double old = C[o];
    C[o] = (decay * old + c1 * (pc[row] * pc[col] + c2 * old)) + S[o];
}

// ------------------------------------------------------------------------------------------
// eigen decomposition helpers
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) symmetrize_kernel(const double *__restrict__ C, double *__restrict__ V, int n,
        int64_t ld) {
    const int i = blockIdx.y, j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    V[(int64_t) i * ld + j] = j <= i ? C[(int64_t) i * ld + j] : C[(int64_t) j * ld + i];
}

// eigenvalue clamps + D = sqrt (cmaes.cpp:249-271); single CTA
__global__ void __launch_bounds__(1024) eig_post_kernel(double *__restrict__ W, double *__restrict__ D,
        double *__restrict__ C, int n, int64_t ld, const int *__restrict__ info, bbgpu_cma_report *rep) {
    __shared__ double shift_s;
    if (threadIdx.x == 0) {
        rep->eig_info = *info;
        shift_s = 0.0;
    }
    __syncthreads();
    if (W[0] <= 0.0) {
        __syncthreads();
        for (int i = threadIdx.x; i < n; i += blockDim.x) W[i] = fmax(W[i], 0.0);
        __syncthreads();
        const double sh = W[n - 1] / 1e14;
        __syncthreads();
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            C[(int64_t) i * ld + i] += sh;
            W[i] += sh;
        }
        __syncthreads();
    }
    if (W[n - 1] > 1e14 * W[0]) {
        const double sh = W[n - 1] / 1e14 - W[0];
        __syncthreads();
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            C[(int64_t) i * ld + i] += sh;
            W[i] += sh;
        }
        __syncthreads();
    }
    for (int i = threadIdx.x; i < n; i += blockDim.x) D[i] = sqrt(W[i]);
    __syncthreads();
    if (threadIdx.x == 0) {
        rep->d_min = D[0];
        rep->d_max = D[n - 1];
    }
}

__global__ void ends_report_kernel(const double *__restrict__ D, int n, bbgpu_cma_report *rep) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        rep->d_min = D[0];
        rep->d_max = D[n - 1];
    }
}

__global__ void __launch_bounds__(256) transpose_kernel(const double *__restrict__ in, double *__restrict__ out, int n,
        int64_t ld) {
    __shared__ double tile[32][33];
    const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += 8) {
        const int i = by + r, j = bx + threadIdx.x;
        tile[r][threadIdx.x] = (i < n && j < n) ? in[(int64_t) i * ld + j] : 0.0;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += 8) {
        const int i = bx + r, j = by + threadIdx.x;
        if (i < n && j < n) out[(int64_t) i * ld + j] = tile[threadIdx.x][r];
    }
}

// ------------------------------------------------------------------------------------------
// convergence probe (the O(n) tests of converged())
// ------------------------------------------------------------------------------------------
template<bool SEP>
__global__ void __launch_bounds__(1024) probe_kernel(int n, int64_t ld, const double *__restrict__ xmean,
        const double *__restrict__ pc, const double *__restrict__ C, const double *__restrict__ D,
        const double *__restrict__ Bt, double sigma, int iaxis, bbgpu_cma_report *rep) {
    __shared__ double red[32];
    double tol = -INFINITY;
    int axis_all = 1, coor_any = 0;
    const double da = D[iaxis];
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const double m = xmean[i];
        const double spread = SEP ? D[i] : sqrt(C[(int64_t) i * ld + i]);
        tol = fmax(tol, fmax(pc[i], spread));
        if (!SEP) {
            // row iaxis of the reference's _b == column iaxis of Bt (cmaes.cpp:211: _b[iaxis][i])
            const double b = Bt[(int64_t) i * ld + iaxis];
            const double step = __dmul_rn(__dmul_rn(__dmul_rn(0.1, sigma), da), b);
            if (m != __dadd_rn(m, step)) axis_all = 0;
        }
        const double cstep = __dmul_rn(__dmul_rn(0.2, sigma), spread);
        if (m == __dadd_rn(m, cstep)) coor_any = 1;
    }
    tol = block_max(tol, red);
    axis_all = __syncthreads_and(axis_all);
    coor_any = __syncthreads_or(coor_any);
    if (threadIdx.x == 0) {
        rep->tolx_max = tol;
        if (SEP) {
            const double m = xmean[iaxis];
            rep->noeffect_axis = (m == __dadd_rn(m, __dmul_rn(__dmul_rn(0.1, sigma), da))) ? 1 : 0;
        } else {
            rep->noeffect_axis = axis_all;
        }
        rep->noeffect_coor = coor_any;
        rep->d_min = D[0];
        rep->d_max = D[n - 1];
    }
}

__global__ void __launch_bounds__(256) fill_identity_kernel(double *__restrict__ M, int n, int64_t ld, bool keep_offdiag) {
    const int i = blockIdx.y, j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    if (i == j) M[(int64_t) i * ld + j] = 1.0;
    else if (!keep_offdiag) M[(int64_t) i * ld + j] = 0.0;
}

__global__ void __launch_bounds__(256) fill_kernel(double *__restrict__ v, int n, double x) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[i] = x;
}

// ------------------------------------------------------------------------------------------
// host-side phase drivers
// ------------------------------------------------------------------------------------------
static int refresh_basis_products(bbgpu_cma *h) {
    bbgpu_ctx *ctx = h->ctx;
    const int n = h->n;
    scale_basis_kernel<<<n, 256, 0, ctx->stream>>>(h->Bt, h->D, h->BDt, h->BDit, n, h->ldn);
    BB_TRY(launched(ctx, "scale_basis_kernel"));
    // C^-1/2[i][l] = sum_j Bt[j][i] * (Bt[j][l] / D[j])          (cmaes.cpp:274-282)
    GemmArgs g {};
    g.A = h->Bt; g.lda = h->ldn; g.B = h->BDit; g.ldb = h->ldn;
    g.M = n; g.N = n; g.K = n; g.kscale = nullptr;
    g.C = h->isC; g.ldc = h->ldn;
    return launch_gemm<2, EPI_STORE>(ctx, g);
}

static int do_sample(bbgpu_cma *h, double sigma) {
    bbgpu_ctx *ctx = h->ctx;
    h->last_sigma = sigma;
    ProfScope prof(ctx, (h->full || h->chol || h->xnes) ? BBGPU_PROF_SAMPLE_GEMM : BBGPU_PROF_SEP_SAMPLE);
    const double *lo = h->cfg.bound ? h->lower : nullptr, *hi = h->cfg.bound ? h->upper : nullptr;
    if (!h->full && !h->chol && !h->xnes) {
        dim3 grid(ceil_div(h->n, 256), std::min(h->lam_loc, 64));
        sep_sample_kernel<<<grid, 256, 0, ctx->stream>>>(h->X, h->Z, h->ldn, h->lam_loc, h->n, h->xmean, h->D, sigma, lo, hi);
        return launched(ctx, "sep_sample_kernel");
    }
    // X[k][i] = m_i + sigma * sum_j Z[k][j] * (D_j Bt[j][i]), clipped in the epilogue
    GemmArgs g {};
    // Cholesky kind: x = m + sigma A z, the operand is A^T (rows [0, n) of the stacked buffer; cholesky_cmaes.cpp:120-135)
    g.A = h->Z; g.lda = h->ldn; g.B = h->chol ? h->Ystk : (h->xnes ? h->Bt : h->BDt); g.ldb = h->ldn;
    g.M = h->lam_loc; g.N = h->n; g.K = h->n;
    g.C = h->X; g.ldc = h->ldn; g.vec = h->xmean; g.lo = lo; g.hi = hi; g.s0 = sigma;
    return launch_gemm<1, EPI_SAMPLE>(ctx, g);
}

static int launch_philox(bbgpu_cma *h, cudaStream_t st, double *Z, uint64_t generation) {
    // counters carry the GLOBAL candidate id: a sharded run draws exactly what the unsharded run draws
    const int64_t total = (int64_t) h->lam_loc * ((h->n + 1) / 2);
    philox_normal_kernel<<<ceil_div(total, 256), 256, 0, st>>>(Z, h->ldn, h->lam_loc, h->n, h->row0, h->cfg.seed, generation, 0u);
    return launched(h->ctx, "philox_normal_kernel");
}

// make h->Z hold the draws of `generation` (prefetched on the side stream if possible)
static int do_philox(bbgpu_cma *h, uint64_t generation) {
    bbgpu_ctx *ctx = h->ctx;
    ProfScope prof(ctx, BBGPU_PROF_PHILOX);
    for (int b = 0; b < 2; b++) {
        if (h->zgen[b] == (long long) generation) {
            BB_CUDA(cudaStreamWaitEvent(ctx->stream, h->ev_z[b], 0));
            h->zcur = b;
            h->Z = h->Zbuf[b];
            return BBGPU_OK;
        }
    }
    const int b = h->zcur ^ 1;
    if (h->zgen[b] >= 0) BB_CUDA(cudaStreamWaitEvent(ctx->stream, h->ev_z[b], 0)); // an unused prefetch may still be writing it
    h->zgen[b] = -1;
    BB_TRY(launch_philox(h, ctx->stream, h->Zbuf[b], generation));
    h->zcur = b;
    h->Z = h->Zbuf[b];
    return BBGPU_OK;
}

// after sample(generation) has been enqueued: start generating the next generation's draws
static int prefetch_philox(bbgpu_cma *h, uint64_t generation) {
    bbgpu_ctx *ctx = h->ctx;
    if (!ctx->aux) BB_CUDA(cudaStreamCreateWithFlags(&ctx->aux, cudaStreamNonBlocking));
    const int cur = h->zcur, other = cur ^ 1;
    BB_CUDA(cudaEventRecord(h->ev_used[cur], ctx->stream));
    h->used_rec[cur] = true;
    if (h->used_rec[other]) BB_CUDA(cudaStreamWaitEvent(ctx->aux, h->ev_used[other], 0));
    BB_TRY(launch_philox(h, ctx->aux, h->Zbuf[other], generation + 1));
    BB_CUDA(cudaEventRecord(h->ev_z[other], ctx->aux));
    h->zgen[other] = (long long) generation + 1;
    return BBGPU_OK;
}

static int do_rank(bbgpu_cma *h, int ik) {
    ProfScope prof(h->ctx, BBGPU_PROF_RANK);
    if (h->lambda <= RANK_COUNT_MAX) {
        // ibw / ybw / f_kth are written by the scatter kernel itself (base_cmaes.cpp:222-229)
        RankEnds ends { h->d_report->ibw, h->d_report->ybw, &h->d_report->f_kth, ik };
        return rank_fitness(h->ctx, h->fit, h->lambda, h->rank_idx, h->rank_val, &ends);
    }
    BB_TRY(rank_fitness(h->ctx, h->fit, h->lambda, h->rank_idx, h->rank_val));
    rank_report_kernel<<<1, 32, 0, h->ctx->stream>>>(h->rank_idx, h->rank_val, h->lambda, ik, h->d_report);
    return launched(h->ctx, "rank_report_kernel");
}

// M (lower, in Mch) -> its Cholesky factor in place: 64-column panels; the diagonal block and the panel solve in
// chol_panel_kernel, the trailing update M22 -= L21 L21^T on the DMMA SYRK kernel (operand = the panel transposed)
static int chol_factor(bbgpu_cma *h) {
    bbgpu_ctx *ctx = h->ctx;
    const int n = h->n;
    for (int k0 = 0; k0 < n; k0 += CH_NB) {
        const int nb = std::min(CH_NB, n - k0), rem = n - k0 - nb;
        chol_panel_kernel<<<std::max(1, ceil_div(rem, CH_NB)), 256, 0, ctx->stream>>>(h->Mch, h->ldn, n, k0, nb, h->Lt, h->dev_info);
        BB_TRY(launched(ctx, "chol_panel_kernel"));
        if (rem <= 0) break;
        GemmArgs g {};
        g.A = h->Lt + (k0 + nb); g.lda = h->ldn; g.B = g.A; g.ldb = h->ldn;
        g.M = rem; g.N = rem; g.K = nb; g.kscale = h->negone; g.kscale_padded = true;
        g.C = h->Mch + (int64_t) (k0 + nb) * (h->ldn + 1); g.ldc = h->ldn; g.vec = h->pc; g.s0 = 1.0; g.s1 = 0.0; g.dscal = h->dscal;
        BB_TRY((launch_gemm<2, EPI_SYRK>(ctx, g)));
    }
    return BBGPU_OK;
}

// CholeskyCmaes::updateDistribution up to and excluding updateSigma() (cholesky_cmaes.cpp:54-114)
static int do_update_chol(bbgpu_cma *h, double sigma) {
    bbgpu_ctx *ctx = h->ctx;
    const bbgpu_cma_config &c = h->cfg;
    const int n = h->n, mu = h->mu;
    cudaStream_t st = ctx->stream;
    const double *lo = c.bound ? h->lower : nullptr, *hi = c.bound ? h->upper : nullptr;
    const double csc = std::sqrt(c.cs * (2.0 - c.cs) * c.mueff);
    const double ccc = std::sqrt(c.cc * (2.0 - c.cc) * c.mueff);
    const double acoeff = std::sqrt(1.0 - c.c1 - c.cmu);
    if (!ctx->aux) BB_CUDA(cudaStreamCreateWithFlags(&ctx->aux, cudaStreamNonBlocking));
    {
        ProfScope prof(ctx, BBGPU_PROF_UPDATE_MISC);
        BB_CUDA(cudaMemcpyAsync(h->xold, h->xmean, (size_t) n * 8, cudaMemcpyDeviceToDevice, st));
        dim3 pgrid(ceil_div(n, 256), h->splits);
        recombine_partial_kernel<false><<<pgrid, 256, 0, st>>>(h->part, h->X, h->ldn, n, mu, h->splits, h->rank_idx, h->weights,
                h->xold, sigma);
        BB_TRY(launched(ctx, "recombine_partial_kernel"));
        recombine_final_kernel<false><<<ceil_div(n, 64), 256, 0, st>>>(h->part, n, h->splits, h->xmean, h->xold, h->dm, 1.0, sigma,
                lo, hi, false, h->ps, nullptr, c.cs, csc, nullptr);
        BB_TRY(launched(ctx, "recombine_final_kernel"));
        const int coef_len = (int) round_up(n + mu, 32) + 32;
        chol_paths_kernel<<<ceil_div(coef_len, 256), 256, 0, st>>>(n, mu, h->pc, h->dm, c.cc, ccc, acoeff, c.cmu, h->weights, h->coef,
                coef_len);
        BB_TRY(launched(ctx, "chol_paths_kernel"));
        // p_sigma needs A(t)^-1 dmean with the OLD factor (:93-104): on the side stream, beside the SYRK and the factorisation
        BB_CUDA(cudaEventRecord(h->ev_ch_fork, st));
        BB_CUDA(cudaStreamWaitEvent(ctx->aux, h->ev_ch_fork, 0));
        if ((size_t) (n + 64) * 8 > 48 * 1024)
            BB_CUDA(cudaFuncSetAttribute(chol_ps_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) ((n + 64) * 8)));
        chol_ps_kernel<<<1, 512, (size_t) (n + 64) * 8, ctx->aux>>>(h->Ystk, h->ldn, n, h->dm, h->ps, c.cs, csc, h->d_report);
        BB_TRY(launched(ctx, "chol_ps_kernel"));
        BB_CUDA(cudaEventRecord(h->ev_ch_join, ctx->aux));
        chol_steps_kernel<<<mu, 256, 0, st>>>(h->Ystk + (int64_t) n * h->ldn, h->X, h->ldn, n, h->xmean, sigma);
        BB_TRY(launched(ctx, "chol_steps_kernel"));
    }
    {
        // M (lower) = c1 pc pc^T + sum_k coef_k Ystk[k]^T Ystk[k]  over the n columns of A and the mu steps
        GemmArgs g {};
        g.A = h->Ystk; g.lda = h->ldn; g.B = h->Ystk; g.ldb = h->ldn;
        g.M = n; g.N = n; g.K = n + mu; g.kscale = h->coef; g.kscale_padded = true;
        g.C = h->Mch; g.ldc = h->ldn; g.vec = h->pc; g.s0 = 0.0; g.s1 = c.c1; g.dscal = h->dscal;
        ProfScope prof(ctx, BBGPU_PROF_COV_SYRK);
        BB_TRY((launch_gemm<2, EPI_SYRK>(ctx, g)));
    }
    ProfScope prof(ctx, BBGPU_PROF_EIG); // the factorisation takes the decomposition's place in the accounting
    BB_TRY(chol_factor(h));
    BB_CUDA(cudaStreamWaitEvent(st, h->ev_ch_join, 0)); // the old factor has been read
    dim3 tgrid(ceil_div(n, 32), ceil_div(n, 32));
    chol_store_kernel<<<tgrid, 256, 0, st>>>(h->Mch, h->Ystk, n, h->ldn, h->dev_info, h->d_report);
    return launched(ctx, "chol_store_kernel");
}

static int do_update(bbgpu_cma *h, double sigma, int fev) {
    if (h->chol) return do_update_chol(h, sigma);
    bbgpu_ctx *ctx = h->ctx;
    const bbgpu_cma_config &c = h->cfg;
    const int n = h->n, mu = h->mu;
    cudaStream_t st = ctx->stream;
    const double *lo = c.bound ? h->lower : nullptr, *hi = c.bound ? h->upper : nullptr;
    BB_CUDA(cudaMemcpyAsync(h->xold, h->xmean, (size_t) n * 8, cudaMemcpyDeviceToDevice, st));
    const double csc = std::sqrt(c.cs * (2.0 - c.cs) * c.mueff);
    const double ccc = std::sqrt(c.cc * (2.0 - c.cc) * c.mueff);
    const double denom = 1.0 - std::pow(1.0 - c.cs, 2.0 * fev / h->lambda); // cmaes.cpp:109
    dim3 pgrid(ceil_div(n, 256), h->splits);
    ProfScope *misc = new ProfScope(ctx, BBGPU_PROF_UPDATE_MISC);
    struct MiscGuard { ProfScope *&p; ~MiscGuard() { delete p; p = nullptr; } } guard { misc };
    if (!h->full) {
        recombine_partial_kernel<true><<<pgrid, 256, 0, st>>>(h->part, h->X, h->ldn, n, mu, h->splits, h->rank_idx,
                h->weights, h->xold, sigma);
        BB_TRY(launched(ctx, "recombine_partial_kernel<sep>"));
        recombine_final_kernel<true><<<ceil_div(n, 64), 256, 0, st>>>(h->part, n, h->splits, h->xmean, h->xold, h->dm, 1.0,
                sigma, lo, hi, false, h->ps, h->C, c.cs, csc, h->tmpn);
        BB_TRY(launched(ctx, "recombine_final_kernel<sep>"));
        paths_kernel<true><<<1, 1024, 0, st>>>(n, h->ps, h->pc, h->dm, denom, c.chi, c.cc, ccc, h->dscal, h->d_report, h->C,
                h->D, h->tmpn, c.ccov, c.mueff, nullptr, nullptr, 0, 0.0);
        return launched(ctx, "paths_kernel<sep>");
    }
    select_steps_kernel<<<pgrid, 256, 0, st>>>(h->part, h->Ysel, h->X, h->ldn, n, mu, h->nsel, h->lambda, h->splits,
            h->rank_idx, h->weights, h->xold, sigma);
    BB_TRY(launched(ctx, "select_steps_kernel"));
    recombine_final_kernel<false><<<ceil_div(n, 64), 256, 0, st>>>(h->part, n, h->splits, h->xmean, h->xold, h->dm, c.cm,
            sigma, lo, hi, h->active, h->ps, nullptr, c.cs, csc, nullptr);
    BB_TRY(launched(ctx, "recombine_final_kernel"));
    ps_gemv_kernel<<<ceil_div(n, 2), 256, 0, st>>>(h->isC, h->ldn, n, h->dm, h->ps, c.cs, csc);
    BB_TRY(launched(ctx, "ps_gemv_kernel"));
    paths_kernel<false><<<1, 1024, 0, st>>>(n, h->ps, h->pc, h->dm, denom, c.chi, c.cc, ccc, h->dscal, h->d_report, nullptr,
            nullptr, nullptr, 0.0, c.mueff, h->active ? nullptr : h->coef, h->weights, mu, c.cmu);
    BB_TRY(launched(ctx, "paths_kernel"));
    delete misc;
    misc = nullptr;
    double decay;
    if (h->active) {
        // Mahalanobis norms of the mu worst steps: rows [mu, 2mu) of Ysel times C^-1/2
        GemmArgs g {};
        g.A = h->Ysel + (int64_t) mu * h->ldn; g.lda = h->ldn; g.B = h->isC; g.ldb = h->ldn;
        g.M = mu; g.N = n; g.K = n; g.partial = h->rowsq_part;
        int tiles_n = 0;
        {
            ProfScope prof(ctx, BBGPU_PROF_NORM_GEMM);
            BB_TRY((launch_gemm<1, EPI_ROWSQ>(ctx, g, &tiles_n)));
        }
        if (tiles_n > h->rowsq_tiles_cap) return fail(BBGPU_ERR_ARG, "rowsq partial buffer too small");
        const double cmu1 = c.cmu + c.cneg * (1.0 - c.alphaold);
        active_coef_kernel<<<ceil_div(mu, 256), 256, 0, st>>>(h->rowsq_part, tiles_n, mu, sigma, h->weights, cmu1, c.cneg,
                h->ycoeff, h->coef);
        BB_TRY(launched(ctx, "active_coef_kernel"));
        decay = 1.0 - c.c1 - c.cmu + c.cneg * c.alphaold;
    } else {
        decay = 1.0 - c.c1 - c.cmu;
    }
    // C (lower) <- decay C + c1 (pc pc^T + c2 C) + sum_r coef_r y_r y_r^T
    GemmArgs g {};
    g.A = h->Ysel; g.lda = h->ldn; g.B = h->Ysel; g.ldb = h->ldn;
    g.M = n; g.N = n; g.K = h->nsel; g.kscale = h->coef; g.kscale_padded = true;
    g.C = h->C; g.ldc = h->ldn; g.vec = h->pc; g.s0 = decay; g.s1 = c.c1; g.dscal = h->dscal;
    ProfScope prof(ctx, BBGPU_PROF_COV_SYRK);
    return launch_gemm<2, EPI_SYRK>(ctx, g);
}

// ---- candidate-sharded generation: the three local phases around the two exchange steps ----

// after sampling: fitness of the local candidates into their slice of the full fitness array, the steps
// Y = (X - m)/sigma, and (Active) the Mahalanobis norms of all local candidates
static int shard_evaluate(bbgpu_cma *h, bbgpu_objective *obj) {
    bbgpu_ctx *ctx = h->ctx;
    cudaStream_t st = ctx->stream;
    const int n = h->n;
    if (h->last_sigma <= 0.0) return fail(BBGPU_ERR_ARG, "sharded evaluate needs a preceding sample()");
    {
        ProfScope prof(ctx, BBGPU_PROF_EVAL);
        BB_TRY(evaluate(ctx, obj, h->X, h->ldn, h->lam_loc, h->fit + h->row0));
    }
    ProfScope prof(ctx, BBGPU_PROF_UPDATE_MISC);
    dim3 grid(ceil_div(n, 256), std::min(h->lam_loc, 128));
    shard_steps_kernel<<<grid, 256, 0, st>>>(h->Yloc, h->X, h->ldn, n, h->lam_loc, h->xmean, h->last_sigma);
    BB_TRY(launched(ctx, "shard_steps_kernel"));
    if (h->active) {
        GemmArgs g {};
        g.A = h->Yloc; g.lda = h->ldn; g.B = h->isC; g.ldb = h->ldn;
        g.M = h->lam_loc; g.N = n; g.K = n; g.partial = h->rowsq_loc;
        int tiles_n = 0;
        BB_TRY((launch_gemm<1, EPI_ROWSQ>(ctx, g, &tiles_n)));
        if (tiles_n > h->rowsq_loc_tiles) return fail(BBGPU_ERR_ARG, "rowsq partial buffer too small");
        shard_norms_kernel<<<ceil_div(h->lam_loc, 256), 256, 0, st>>>(h->rowsq_loc, tiles_n, h->lam_loc, h->last_sigma,
                h->norm_all + h->row0);
        BB_TRY(launched(ctx, "shard_norms_kernel"));
    }
    return BBGPU_OK;
}

// exchange step 1: everybody gets every candidate's fitness (and Active norm)
static int shard_exchange_fitness(bbgpu_cma *h) {
    if (!h->comm) return fail(BBGPU_ERR_ARG, "sharded handle without a communicator: exchange BBGPU_CMA_FITNESS yourself");
    ProfScope prof(h->ctx, BBGPU_PROF_COMM);
    if (h->active) return comm_allgather2(h->comm, h->fit, h->norm_all, (size_t) h->lam_loc);
    return comm_allgather(h->comm, h->fit, (size_t) h->lam_loc);
}

// after the (replicated) ranking: this GPU's share of sum w x and of the rank-mu (and active) sum
static int shard_partial(bbgpu_cma *h) {
    bbgpu_ctx *ctx = h->ctx;
    cudaStream_t st = ctx->stream;
    const bbgpu_cma_config &c = h->cfg;
    const int n = h->n;
    {
        ProfScope prof(ctx, BBGPU_PROF_UPDATE_MISC);
        const double cpos = h->active ? c.cmu + c.cneg * (1.0 - c.alphaold) : c.cmu;
        shard_coef_kernel<<<ceil_div(h->lambda, 256), 256, 0, st>>>(h->rank_idx, h->lambda, h->mu, h->row0, h->lam_loc, h->weights,
                h->active, cpos, c.cneg, h->norm_all, h->coef_loc, h->wm_loc, h->ycoeff);
        BB_TRY(launched(ctx, "shard_coef_kernel"));
        dim3 pgrid(ceil_div(n, 256), h->splits);
        shard_mean_partial_kernel<<<pgrid, 256, 0, st>>>(h->part, h->X, h->ldn, n, h->lam_loc, h->splits, h->wm_loc);
        BB_TRY(launched(ctx, "shard_mean_partial_kernel"));
        shard_mean_final_kernel<<<ceil_div(n, 256), 256, 0, st>>>(h->part, n, h->splits, h->red + (size_t) n * h->ldn);
        BB_TRY(launched(ctx, "shard_mean_final_kernel"));
        BB_CUDA(cudaMemsetAsync(h->red, 0, (size_t) n * h->ldn * 8, st));
    }
    // S_local (lower) = sum_k coef_k y_k y_k^T over the local candidates.  Active: every candidate carries a coefficient.
    // Otherwise only the best mu of all lambda do -- about half of the local rows --, so they are compacted first (their
    // number is data dependent: it reaches the kernel through device memory) and the SYRK runs over them alone.
    GemmArgs g {};
    g.A = h->Yloc; g.lda = h->ldn; g.B = h->Yloc; g.ldb = h->ldn;
    g.M = n; g.N = n; g.K = h->lam_loc; g.kscale = h->coef_loc; g.kscale_padded = true;
    if (h->Ycomp != nullptr) {
        ProfScope prof(ctx, BBGPU_PROF_UPDATE_MISC);
        shard_compact_kernel<<<1, 1024, 0, st>>>(h->coef_loc, h->lam_loc, h->sel_idx, h->coef_comp, (int) round_up(h->lam_loc, 32) + 32,
                h->ksel);
        BB_TRY(launched(ctx, "shard_compact_kernel"));
        shard_gather_kernel<<<h->lam_loc, 256, 0, st>>>(h->Ycomp, h->Yloc, h->ldn, n, h->sel_idx, h->ksel);
        BB_TRY(launched(ctx, "shard_gather_kernel"));
        g.A = h->Ycomp; g.B = h->Ycomp; g.kscale = h->coef_comp; g.Kdev = h->ksel;
    }
    g.C = h->red; g.ldc = h->ldn; g.vec = h->pc; g.s0 = 0.0; g.s1 = 0.0; g.dscal = h->dscal;
    ProfScope prof(ctx, BBGPU_PROF_COV_SYRK);
    return launch_gemm<2, EPI_SYRK>(ctx, g);
}

// exchange step 2: sum of the partials over the GPUs (n*ldn covariance entries + the n-vector)
static int shard_exchange_partials(bbgpu_cma *h) {
    if (!h->comm) return fail(BBGPU_ERR_ARG, "sharded handle without a communicator: exchange BBGPU_CMA_SHARD_RED yourself");
    ProfScope prof(h->ctx, BBGPU_PROF_COMM);
    if (h->peer) return comm_allreduce_sum_p2p(h->peer, (size_t) h->n * h->ldn + h->ldn);
    return comm_allreduce_sum(h->comm, h->red, (size_t) h->n * h->ldn + h->ldn);
}

// from the summed partials: mean, paths, covariance -- identical on every GPU
static int shard_finish(bbgpu_cma *h, double sigma, int fev) {
    bbgpu_ctx *ctx = h->ctx;
    cudaStream_t st = ctx->stream;
    const bbgpu_cma_config &c = h->cfg;
    const int n = h->n;
    const double *lo = c.bound ? h->lower : nullptr, *hi = c.bound ? h->upper : nullptr;
    ProfScope prof(ctx, BBGPU_PROF_UPDATE_MISC);
    BB_CUDA(cudaMemcpyAsync(h->xold, h->xmean, (size_t) n * 8, cudaMemcpyDeviceToDevice, st));
    const double csc = std::sqrt(c.cs * (2.0 - c.cs) * c.mueff);
    const double ccc = std::sqrt(c.cc * (2.0 - c.cc) * c.mueff);
    const double denom = 1.0 - std::pow(1.0 - c.cs, 2.0 * fev / h->lambda); // cmaes.cpp:109
    recombine_final_kernel<false><<<ceil_div(n, 64), 256, 0, st>>>(h->red + (size_t) n * h->ldn, n, 1, h->xmean, h->xold, h->dm,
            c.cm, sigma, lo, hi, h->active, h->ps, nullptr, c.cs, csc, nullptr);
    BB_TRY(launched(ctx, "recombine_final_kernel"));
    ps_gemv_kernel<<<ceil_div(n, 2), 256, 0, st>>>(h->isC, h->ldn, n, h->dm, h->ps, c.cs, csc);
    BB_TRY(launched(ctx, "ps_gemv_kernel"));
    paths_kernel<false><<<1, 1024, 0, st>>>(n, h->ps, h->pc, h->dm, denom, c.chi, c.cc, ccc, h->dscal, h->d_report, nullptr,
            nullptr, nullptr, 0.0, c.mueff, nullptr, h->weights, h->mu, c.cmu);
    BB_TRY(launched(ctx, "paths_kernel"));
    const double decay = h->active ? 1.0 - c.c1 - c.cmu + c.cneg * c.alphaold : 1.0 - c.c1 - c.cmu;
    dim3 grid(ceil_div(n, 256), n);
    shard_cov_finish_kernel<<<grid, 256, 0, st>>>(h->C, h->red, h->ldn, n, h->pc, decay, c.c1, h->dscal);
    return launched(ctx, "shard_cov_finish_kernel");
}

// Symmetric eigendecomposition of Cin (n x n, pitch ldn, lower triangle read): eigenvalues ascending in Wout, eigenvectors
// as the ROWS of Vt.  Own solver at 256 <= n <= ~1700, cuSOLVER syevd otherwise; no clamping of the spectrum.
static int eig_core(bbgpu_cma *h, const double *Cin, double *Vt, double *Wout) {
    bbgpu_ctx *ctx = h->ctx;
    const int n = h->n;
    cudaStream_t st = ctx->stream;
    cusolverDnHandle_t sv;
    BB_TRY(get_solver(ctx, &sv));
    // Own solver where it pays and fits (cuSOLVER's syevd spends 80 % of its 12 ms at n = 1000 in a latency-bound sytrd):
    // cooperative sytrd -> tridiagonal divide & conquer -> cuSOLVER ormtr for the back-transformation.
    // BBGPU_EIG=cusolver forces the library path.
    static const bool force_lib = [] { const char *e = getenv("BBGPU_EIG"); return e && e[0] == 'c'; }();
    if (!force_lib && n >= 256 && sytrd_fits(ctx, n)) {
        const size_t nn = (size_t) n;
        // Back-transformation: Q^T is formed row by row from the reflectors (q_rows_kernel: no exchange between CTAs) on
        // the side stream WHILE the divide & conquer runs (its kernels are small and latency-bound), then
        // B^T = Z^T Q^T is one DMMA GEMM.  cuSOLVER's ormtr (0.5 ms at n = 1000: 16 dependent block steps after the
        // divide & conquer) stays as the comparison path, BBGPU_EIG_Q=ormtr.
        static const bool use_ormtr = [] { const char *e = getenv("BBGPU_EIG_Q"); return e && e[0] == 'o'; }();
        const size_t zt = nn * (size_t) h->ldn;
        if (h->own_ws == nullptr) {
            int lw = 0;
            if (cusolverDnDormtr_bufferSize(sv, CUBLAS_SIDE_LEFT, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, n, n, Vt, n, Wout, Vt,
                    (int) h->ldn, &lw) != CUSOLVER_STATUS_SUCCESS) return fail(BBGPU_ERR_SOLVER, "ormtr_bufferSize failed");
            h->own_lwork = lw;
            BB_TRY(dev_alloc(&h->own_ws, 3 * nn + nn * nn + sytrd_workspace_doubles(n) + (size_t) lw + 2 * zt + 48));
            BB_TRY(dev_alloc(&h->own_info, (size_t) 4));
            BB_CUDA(cudaEventCreateWithFlags(&h->ev_eig_fork, cudaEventDisableTiming));
            BB_CUDA(cudaEventCreateWithFlags(&h->ev_eig_join, cudaEventDisableTiming));
        }
        double *dd = h->own_ws, *de = dd + nn, *dtau = de + nn, *Vcm = dtau + nn, *ws = Vcm + nn * nn,
                *owork = ws + sytrd_workspace_doubles(n);
        double *Ztmp = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(owork + h->own_lwork) + 127) & ~(uintptr_t) 127); // TMA operands
        double *Qt = Ztmp + ((zt + 15) & ~(size_t) 15);
        BB_CUDA(cudaMemsetAsync(de, 0, 2 * nn * 8, st));
        BB_TRY(sytrd_device(ctx, n, Cin, (int) h->ldn, dd, de, dtau, Vcm, ws));
        if (!use_ormtr && !ctx->aux) BB_CUDA(cudaStreamCreateWithFlags(&ctx->aux, cudaStreamNonBlocking));
        if (use_ormtr) {
            BB_TRY(stedc_device(ctx, n, dd, de, Wout, Vt, (int) h->ldn, h->dev_info));
            const cusolverStatus_t rc = cusolverDnDormtr(sv, CUBLAS_SIDE_LEFT, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, n, n, Vcm, n, dtau,
                    Vt, (int) h->ldn, owork, h->own_lwork, h->own_info);
            if (rc != CUSOLVER_STATUS_SUCCESS) return fail(BBGPU_ERR_SOLVER, "cusolverDnDormtr failed with status %d", (int) rc);
        } else {
            BB_CUDA(cudaEventRecord(h->ev_eig_fork, st));
            BB_CUDA(cudaStreamWaitEvent(ctx->aux, h->ev_eig_fork, 0));
            BB_TRY(q_rows_device(ctx, ctx->aux, n, Vcm, dtau, Qt, (int) h->ldn));
            BB_CUDA(cudaEventRecord(h->ev_eig_join, ctx->aux));
            BB_TRY(stedc_device(ctx, n, dd, de, Wout, Ztmp, (int) h->ldn, h->dev_info));
            BB_CUDA(cudaStreamWaitEvent(st, h->ev_eig_join, 0));
            // rows of Ztmp are the eigenvectors of T: B^T (rows = eigenvectors of C) = Z^T Q^T
            GemmArgs g {};
            g.A = Ztmp; g.lda = h->ldn; g.B = Qt; g.ldb = h->ldn; g.M = n; g.N = n; g.K = n; g.C = Vt; g.ldc = h->ldn;
            BB_TRY((launch_gemm<1, EPI_STORE>(ctx, g)));
        }
    } else {
        dim3 grid(ceil_div(n, 256), n);
        symmetrize_kernel<<<grid, 256, 0, st>>>(Cin, Vt, n, h->ldn);
        BB_TRY(launched(ctx, "symmetrize_kernel"));
        if (h->solver_work == nullptr) {
            int lwork = 0;
            if (cusolverDnDsyevd_bufferSize(sv, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n, Vt, (int) h->ldn, Wout,
                    &lwork) != CUSOLVER_STATUS_SUCCESS) return fail(BBGPU_ERR_SOLVER, "syevd_bufferSize failed");
            BB_TRY(dev_alloc(&h->solver_work, (size_t) lwork));
            h->solver_lwork = lwork;
        }
        const cusolverStatus_t rc = cusolverDnDsyevd(sv, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n, Vt, (int) h->ldn,
                Wout, h->solver_work, h->solver_lwork, h->dev_info);
        if (rc != CUSOLVER_STATUS_SUCCESS) return fail(BBGPU_ERR_SOLVER, "cusolverDnDsyevd failed with status %d", (int) rc);
    }
    return BBGPU_OK;
}

static int do_eig(bbgpu_cma *h, const double *h_B, const double *h_D) {
    bbgpu_ctx *ctx = h->ctx;
    ProfScope prof(ctx, BBGPU_PROF_EIG);
    const int n = h->n;
    cudaStream_t st = ctx->stream;
    if (h_B != nullptr) {
        // eigenbasis replay: adopt the oracle's (B, D); Bt = B^T
        void *hp = nullptr;
        BB_TRY(ctx->pinned((size_t) n * n * 8, &hp));
        double *hb = (double*) hp;
        BB_CUDA(cudaStreamSynchronize(st));
        for (int i = 0; i < n; i++)
            for (int j = 0; j < n; j++) hb[(size_t) j * n + i] = h_B[(size_t) i * n + j];
        BB_CUDA(cudaMemcpy2DAsync(h->Bt, (size_t) h->ldn * 8, hb, (size_t) n * 8, (size_t) n * 8, n, cudaMemcpyHostToDevice, st));
        BB_CUDA(cudaMemcpyAsync(h->D, h_D, (size_t) n * 8, cudaMemcpyHostToDevice, st));
        BB_CUDA(cudaStreamSynchronize(st));
        ends_report_kernel<<<1, 32, 0, st>>>(h->D, n, h->d_report);
        BB_TRY(launched(ctx, "ends_report_kernel"));
        return refresh_basis_products(h);
    }
    BB_TRY(eig_core(h, h->C, h->Bt, h->W));
    eig_post_kernel<<<1, 1024, 0, st>>>(h->W, h->D, h->C, n, h->ldn, h->dev_info, h->d_report);
    BB_TRY(launched(ctx, "eig_post_kernel"));
    return refresh_basis_products(h);
}

static int field_info(bbgpu_cma *h, int field, double **ptr, int64_t *rows, int64_t *cols, int64_t *ld) {
    const int n = h->n;
    *ld = 0;
    switch (field) {
    case BBGPU_CMA_XMEAN: *ptr = h->xmean; *rows = 1; *cols = n; break;
    case BBGPU_CMA_XOLD: *ptr = h->xold; *rows = 1; *cols = n; break;
    case BBGPU_CMA_PC: *ptr = h->pc; *rows = 1; *cols = n; break;
    case BBGPU_CMA_PS: *ptr = h->ps; *rows = 1; *cols = n; break;
    case BBGPU_CMA_D:
        if (h->chol || h->lm || h->xnes) return fail(BBGPU_ERR_ARG, "this kind has no D");
        *ptr = h->D; *rows = 1; *cols = n; break;
    case BBGPU_CMA_PCMAT: case BBGPU_CMA_VMAT:
        if (!h->lm) return fail(BBGPU_ERR_ARG, "only LM-CMA-ES has direction vectors");
        *ptr = field == BBGPU_CMA_PCMAT ? h->PCmat : h->Vmat; *rows = h->lm_mem; *cols = n; *ld = h->ldn; break;
    case BBGPU_CMA_LMB: case BBGPU_CMA_LMD:
        if (!h->lm) return fail(BBGPU_ERR_ARG, "only LM-CMA-ES has direction vectors");
        *ptr = field == BBGPU_CMA_LMB ? h->lmb : h->lmd; *rows = 1; *cols = h->lm_mem; break;
    case BBGPU_CMA_A:
        if (!h->chol) return fail(BBGPU_ERR_ARG, "only Cholesky-CMA-ES has a factor A");
        *ptr = h->Ystk; *rows = n; *cols = n; *ld = h->ldn; break;
    case BBGPU_CMA_C:
        if (h->chol || h->lm || h->xnes) return fail(BBGPU_ERR_ARG, "this kind keeps no covariance matrix");
        if (h->full) { *ptr = h->C; *rows = n; *cols = n; *ld = h->ldn; }
        else { *ptr = h->C; *rows = 1; *cols = n; }
        break;
    case BBGPU_CMA_INVSQRTC:
        if (!h->full) return fail(BBGPU_ERR_ARG, "separable CMA-ES has no C^-1/2");
        *ptr = h->isC; *rows = n; *cols = n; *ld = h->ldn; break;
    case BBGPU_CMA_B:
        if (!h->full && !h->xnes) return fail(BBGPU_ERR_ARG, "this kind has no matrix B");
        *ptr = h->Bt; *rows = n; *cols = n; *ld = h->ldn; break;
    case BBGPU_CMA_ARX: *ptr = h->X; *rows = h->lam_loc; *cols = n; *ld = h->ldn; break;
    case BBGPU_CMA_Z:
        if (h->lm) { *ptr = h->Zp; *rows = h->lm_pairs; *cols = n; *ld = h->ldn; break; } // one base vector per pair
        *ptr = h->Z; *rows = h->lam_loc; *cols = n; *ld = h->ldn; break;
    case BBGPU_CMA_NORMS:
        if (!(h->active && h->shards > 1)) return fail(BBGPU_ERR_ARG, "norms exist only for a sharded active CMA-ES");
        *ptr = h->norm_all; *rows = 1; *cols = h->lambda; break;
    case BBGPU_CMA_SHARD_RED:
        if (h->shards <= 1) return fail(BBGPU_ERR_ARG, "not a sharded handle");
        *ptr = h->red; *rows = 1; *cols = (int64_t) n * h->ldn + h->ldn; break;
    case BBGPU_CMA_FITNESS: *ptr = h->fit; *rows = 1; *cols = h->lambda; break;
    case BBGPU_CMA_RANKVAL: *ptr = h->rank_val; *rows = 1; *cols = h->lambda; break;
    case BBGPU_CMA_YCOEFF:
        if (!h->active) return fail(BBGPU_ERR_ARG, "ycoeff exists only for active CMA-ES");
        *ptr = h->ycoeff; *rows = 1; *cols = h->mu; break;
    case BBGPU_CMA_WEIGHTS: *ptr = h->weights; *rows = 1; *cols = h->mu; break;
    default: return fail(BBGPU_ERR_ARG, "unknown CMA field %d", field);
    }
    if (*ld == 0) *ld = *cols;
    return BBGPU_OK;
}

} // namespace bb

using namespace bb;

extern "C" {

int bbgpu_cma_create(bbgpu_ctx *ctx, const bbgpu_cma_config *cfg, const double *h_weights, const double *h_lower,
        const double *h_upper, bbgpu_cma **out) {
    BB_ARG(ctx && cfg && h_weights && out);
    BB_ARG(cfg->n >= 1 && cfg->lambda >= 2 && cfg->mu >= 1 && cfg->mu <= cfg->lambda);
    BB_ARG(cfg->kind >= BBGPU_CMA_FULL && cfg->kind <= BBGPU_CMA_XNES);
    if (cfg->kind == BBGPU_CMA_XNES) BB_ARG(cfg->lambda <= XN_MAXNP && cfg->mu == cfg->lambda && !cfg->bound);
    if (cfg->kind == BBGPU_CMA_LM) BB_ARG(cfg->lm_memory >= 1 && cfg->lm_memory <= LM_MAXMEM);
    BB_ARG(!cfg->bound || (h_lower && h_upper));
    if (cfg->shard_count > 1) {
        BB_ARG(cfg->shard_rank >= 0 && cfg->shard_rank < cfg->shard_count);
        if (cfg->kind != BBGPU_CMA_FULL && cfg->kind != BBGPU_CMA_ACTIVE)
            return fail(BBGPU_ERR_ARG, "candidate sharding covers the full-covariance kinds");
        if (cfg->lambda % cfg->shard_count != 0 || cfg->lambda / cfg->shard_count < 2)
            return fail(BBGPU_ERR_ARG, "lambda = %d does not split into %d equal shards of at least 2", cfg->lambda, cfg->shard_count);
    }
    BB_CUDA(cudaSetDevice(ctx->device));
    bbgpu_cma *h = new bbgpu_cma();
    h->ctx = ctx;
    h->cfg = *cfg;
    h->n = cfg->n;
    h->lambda = cfg->lambda;
    h->mu = cfg->mu;
    h->full = cfg->kind == BBGPU_CMA_FULL || cfg->kind == BBGPU_CMA_ACTIVE;
    h->active = cfg->kind == BBGPU_CMA_ACTIVE;
    h->chol = cfg->kind == BBGPU_CMA_CHOLESKY;
    h->lm = cfg->kind == BBGPU_CMA_LM;
    h->xnes = cfg->kind == BBGPU_CMA_XNES;
    h->lm_mem = h->lm ? cfg->lm_memory : 0;
    h->lm_pairs = (cfg->lambda + 1) / 2;
    h->nsel = h->active ? 2 * h->mu : h->mu;
    if (h->active && 2 * h->mu > h->lambda) { delete h; return fail(BBGPU_ERR_ARG, "active CMA-ES needs 2 mu <= lambda"); }
    h->ldn = round_up(h->n, 8);
    h->shards = std::max(1, cfg->shard_count);
    h->shard = h->shards > 1 ? cfg->shard_rank : 0;
    h->lam_loc = h->lambda / h->shards;
    h->row0 = h->shard * h->lam_loc;
    const int n = h->n;
    const size_t ldn = (size_t) h->ldn, lam = (size_t) h->lambda, laml = (size_t) h->lam_loc;
    h->splits = std::max(1, std::min(128, std::min(h->nsel / 16, ceil_div(16 * ctx->sm_count, ceil_div(n, 256)))));
    int rc = BBGPU_OK;
    auto A = [&](double **p, size_t count) {
        if (rc != BBGPU_OK) return;
        rc = dev_alloc(p, count);
        if (rc == BBGPU_OK && count) {
            if (cudaMemsetAsync(*p, 0, count * 8, ctx->stream) != cudaSuccess) rc = fail(BBGPU_ERR_CUDA, "memset failed");
        }
    };
    A(&h->xmean, ldn); A(&h->xold, ldn); A(&h->pc, ldn); A(&h->ps, ldn); A(&h->D, ldn); A(&h->dm, ldn);
    A(&h->tmpn, ldn); A(&h->W, ldn); A(&h->weights, h->mu); A(&h->lower, ldn); A(&h->upper, ldn);
    A(&h->coef, (size_t) round_up(h->chol ? h->n + h->mu : h->nsel, 32) + 32); A(&h->ycoeff, (size_t) h->mu); A(&h->dscal, 8);
    A(&h->X, laml * ldn); A(&h->fit, lam); A(&h->rank_val, lam);
    if (!h->lm) { A(&h->Zbuf[0], laml * ldn); A(&h->Zbuf[1], laml * ldn); } // the LM kind draws lambda/2 base vectors (Zp)
    A(&h->part, (size_t) h->splits * 2 * n);
    if (h->full || h->xnes) {
        A(&h->C, (size_t) n * ldn); A(&h->Bt, (size_t) n * ldn); A(&h->BDt, (size_t) n * ldn);
        A(&h->BDit, (size_t) n * ldn); A(&h->isC, (size_t) n * ldn);
        if (h->xnes) {
            // nothing else: the utilities live in `weights`
        } else if (h->shards > 1) {
            A(&h->Yloc, laml * ldn); A(&h->coef_loc, (size_t) round_up(h->lam_loc, 32) + 32); A(&h->wm_loc, laml);
            A(&h->red, (size_t) n * ldn + ldn);
            static const bool no_compact = [] { const char *e = getenv("BBGPU_SHARD_COMPACT"); return e && e[0] == '0'; }();
            if (!h->active && !no_compact) {
                A(&h->Ycomp, laml * ldn); A(&h->coef_comp, (size_t) round_up(h->lam_loc, 32) + 32);
                if (rc == BBGPU_OK) rc = dev_alloc(&h->sel_idx, laml);
                if (rc == BBGPU_OK) rc = dev_alloc(&h->ksel, (size_t) 1);
            }
            if (h->active) {
                h->rowsq_loc_tiles = ceil_div(n, 64);
                A(&h->norm_all, lam); A(&h->rowsq_loc, (size_t) h->rowsq_loc_tiles * laml);
            }
        } else {
            A(&h->Ysel, (size_t) h->nsel * ldn);
            if (h->active) {
                h->rowsq_tiles_cap = ceil_div(n, 64);
                A(&h->rowsq_part, (size_t) h->rowsq_tiles_cap * h->mu);
            }
        }
    } else if (h->chol) {
        A(&h->Ystk, (size_t) (n + h->mu) * ldn); A(&h->Mch, (size_t) n * ldn); A(&h->Lt, (size_t) CH_NB * ldn);
        A(&h->negone, (size_t) CH_NB + 32); A(&h->rnorm, lam); A(&h->rstat, 8);
        if (rc == BBGPU_OK && (cudaEventCreateWithFlags(&h->ev_ch_fork, cudaEventDisableTiming) != cudaSuccess ||
                cudaEventCreateWithFlags(&h->ev_ch_join, cudaEventDisableTiming) != cudaSuccess))
            rc = fail(BBGPU_ERR_CUDA, "cudaEventCreate failed");
        if (rc == BBGPU_OK) {
            fill_kernel<<<1, 64, 0, ctx->stream>>>(h->negone, CH_NB, -1.0);
            rc = launched(ctx, "fill_kernel");
        }
    } else if (h->lm) {
        A(&h->PCmat, (size_t) h->lm_mem * ldn); A(&h->Vmat, (size_t) h->lm_mem * ldn); A(&h->lmb, (size_t) h->lm_mem);
        A(&h->lmd, (size_t) h->lm_mem); A(&h->Zp, (size_t) h->lm_pairs * ldn);
        if (rc == BBGPU_OK) rc = dev_alloc(&h->lm_jarr, (size_t) h->lm_mem);
        if (rc == BBGPU_OK) rc = dev_alloc(&h->lm_i0, (size_t) h->lm_pairs);
        if (rc == BBGPU_OK) rc = dev_alloc(&h->lm_flags, (size_t) h->lm_mem + 8);
        if (n > LM_TP * 1024) A(&h->lm_tglob, (size_t) ctx->sm_count * ldn);
        // slabs: enough CTAs for two per SM, at least 512 columns each
        h->lm_slabs = std::max(1, std::min(std::min((int) LM_MAXSLABS, (int) ceil_div(2 * ctx->sm_count, h->lm_pairs)), n / 512));
        h->lm_part_ld = (int) round_up(std::max(h->lm_mem, 1), 8);
        A(&h->lm_part, (size_t) h->lm_pairs * h->lm_slabs * h->lm_part_ld);
    } else {
        A(&h->C, ldn);
    }
    h->Z = h->Zbuf[0];
    for (int b = 0; b < 2 && rc == BBGPU_OK; b++) {
        if (cudaEventCreateWithFlags(&h->ev_z[b], cudaEventDisableTiming) != cudaSuccess ||
                cudaEventCreateWithFlags(&h->ev_used[b], cudaEventDisableTiming) != cudaSuccess)
            rc = fail(BBGPU_ERR_CUDA, "cudaEventCreate failed");
    }
    if (rc == BBGPU_OK) rc = dev_alloc(&h->rank_idx, lam);
    if (rc == BBGPU_OK) rc = dev_alloc(&h->dev_info, 1);
    if (rc == BBGPU_OK) rc = dev_alloc(&h->d_report, 1);
    if (rc == BBGPU_OK && cudaMallocHost((void**) &h->h_report, sizeof(bbgpu_cma_report)) != cudaSuccess)
        rc = fail(BBGPU_ERR_CUDA, "cudaMallocHost(report) failed");
    if (rc != BBGPU_OK) { bbgpu_cma_destroy(h); return rc; }
    memset(h->h_report, 0, sizeof(bbgpu_cma_report));
    cudaMemsetAsync(h->d_report, 0, sizeof(bbgpu_cma_report), ctx->stream);
    cudaMemsetAsync(h->dev_info, 0, sizeof(int), ctx->stream);
    cudaMemcpyAsync(h->weights, h_weights, (size_t) h->mu * 8, cudaMemcpyHostToDevice, ctx->stream);
    if (h_lower && h_upper) {
        cudaMemcpyAsync(h->lower, h_lower, (size_t) n * 8, cudaMemcpyHostToDevice, ctx->stream);
        cudaMemcpyAsync(h->upper, h_upper, (size_t) n * 8, cudaMemcpyHostToDevice, ctx->stream);
    }
    BB_CUDA(cudaStreamSynchronize(ctx->stream));
    *out = h;
    return BBGPU_OK;
}

int bbgpu_cma_destroy(bbgpu_cma *h) {
    if (!h) return BBGPU_OK;
    cudaStreamSynchronize(h->ctx->stream);
    if (h->ctx->aux) cudaStreamSynchronize(h->ctx->aux);
    if (h->peer) { // collective on a sharded handle: the peers unmap this rank's buffer before it is freed
        comm_peer_region_destroy(h->peer);
        h->peer = nullptr;
    }
    for (int b = 0; b < 2; b++) {
        if (h->ev_z[b]) cudaEventDestroy(h->ev_z[b]);
        if (h->ev_used[b]) cudaEventDestroy(h->ev_used[b]);
    }
    double *bufs[] = { h->xmean, h->xold, h->pc, h->ps, h->D, h->dm, h->tmpn, h->W, h->weights, h->lower, h->upper, h->coef,
            h->ycoeff, h->dscal, h->X, h->Zbuf[0], h->Zbuf[1], h->fit, h->rank_val, h->part, h->C, h->Bt, h->BDt, h->BDit, h->isC, h->Ysel,
            h->rowsq_part, h->solver_work, h->ssq, h->Yloc, h->coef_loc, h->wm_loc, h->norm_all, h->red, h->rowsq_loc, h->own_ws,
            h->Ystk, h->Mch, h->Lt, h->negone, h->rnorm, h->rstat, h->PCmat, h->Vmat, h->lmb, h->lmd, h->Zp, h->Ycomp, h->coef_comp, h->lm_part, h->lm_tglob };
    if (h->sel_idx) cudaFree(h->sel_idx);
    if (h->ksel) cudaFree(h->ksel);
    if (h->lm_jarr) cudaFree(h->lm_jarr);
    if (h->lm_i0) cudaFree(h->lm_i0);
    if (h->lm_flags) cudaFree(h->lm_flags);
    if (h->ev_ch_fork) cudaEventDestroy(h->ev_ch_fork);
    if (h->ev_ch_join) cudaEventDestroy(h->ev_ch_join);
    if (h->own_info) cudaFree(h->own_info);
    if (h->ev_eig_fork) cudaEventDestroy(h->ev_eig_fork);
    if (h->ev_eig_join) cudaEventDestroy(h->ev_eig_join);
    for (double *b : bufs) if (b) cudaFree(b);
    if (h->rank_idx) cudaFree(h->rank_idx);
    if (h->dev_info) cudaFree(h->dev_info);
    if (h->d_report) cudaFree(h->d_report);
    if (h->h_report) cudaFreeHost(h->h_report);
    delete h;
    return BBGPU_OK;
}

int bbgpu_cma_reset(bbgpu_cma *h, const double *h_guess, int keep_offdiag) {
    BB_ARG(h && h_guess);
    bbgpu_ctx *ctx = h->ctx;
    const int n = h->n;
    cudaStream_t st = ctx->stream;
    BB_CUDA(cudaMemcpyAsync(h->xmean, h_guess, (size_t) n * 8, cudaMemcpyHostToDevice, st));
    BB_CUDA(cudaMemsetAsync(h->xold, 0, (size_t) n * 8, st));
    BB_CUDA(cudaMemsetAsync(h->pc, 0, (size_t) n * 8, st));
    BB_CUDA(cudaMemsetAsync(h->ps, 0, (size_t) n * 8, st));
    fill_kernel<<<ceil_div(n, 256), 256, 0, st>>>(h->D, n, 1.0);
    BB_TRY(launched(ctx, "fill_kernel"));
    if (h->full) {
        dim3 grid(ceil_div(n, 256), n);
        // B and C keep their off-diagonal entries on a re-init of the same object in the
        // reference (cmaes.cpp:53-60 resize without clear); C^-1/2 is cleared (:55-56)
        fill_identity_kernel<<<grid, 256, 0, st>>>(h->C, n, h->ldn, keep_offdiag != 0);
        BB_TRY(launched(ctx, "fill_identity_kernel"));
        // Bt is the transpose of _b: resetting the diagonal commutes with the transpose
        fill_identity_kernel<<<grid, 256, 0, st>>>(h->Bt, n, h->ldn, keep_offdiag != 0);
        BB_TRY(launched(ctx, "fill_identity_kernel"));
        fill_identity_kernel<<<grid, 256, 0, st>>>(h->isC, n, h->ldn, false);
        BB_TRY(launched(ctx, "fill_identity_kernel"));
        scale_basis_kernel<<<n, 256, 0, st>>>(h->Bt, h->D, h->BDt, h->BDit, n, h->ldn);
        BB_TRY(launched(ctx, "scale_basis_kernel"));
    } else if (h->chol) {
        dim3 grid(ceil_div(n, 256), n);
        fill_identity_kernel<<<grid, 256, 0, st>>>(h->Ystk, n, h->ldn, false); // A = I (cholesky_cmaes.cpp:44-51 clears it)
        BB_TRY(launched(ctx, "fill_identity_kernel"));
    } else if (h->xnes) { // B = (a0 / sigma) I = I (xnes.cpp:83-88)
        dim3 grid(ceil_div(n, 256), n);
        fill_identity_kernel<<<grid, 256, 0, st>>>(h->Bt, n, h->ldn, false);
        BB_TRY(launched(ctx, "fill_identity_kernel"));
    } else if (h->lm) { // memlen = 0: no row is live; b = d = 0 (lm_cmaes.cpp:75-78)
        BB_CUDA(cudaMemsetAsync(h->lmb, 0, (size_t) h->lm_mem * 8, st));
        BB_CUDA(cudaMemsetAsync(h->lmd, 0, (size_t) h->lm_mem * 8, st));
    } else {
        fill_kernel<<<ceil_div(n, 256), 256, 0, st>>>(h->C, n, 1.0);
        BB_TRY(launched(ctx, "fill_kernel"));
    }
    BB_CUDA(cudaStreamSynchronize(st));
    return BBGPU_OK;
}

int bbgpu_cma_set(bbgpu_cma *h, int field, const double *h_src, int64_t count) {
    BB_ARG(h && h_src);
    double *ptr;
    int64_t rows, cols, ld;
    BB_TRY(field_info(h, field, &ptr, &rows, &cols, &ld));
    BB_ARG(count == rows * cols);
    cudaStream_t st = h->ctx->stream;
    if (field == BBGPU_CMA_B || field == BBGPU_CMA_A) { // stored transposed (k-major operand)
        std::vector<double> t((size_t) count);
        const int n = h->n;
        for (int i = 0; i < n; i++)
            for (int j = 0; j < n; j++) t[(size_t) j * n + i] = h_src[(size_t) i * n + j];
        BB_CUDA(cudaMemcpy2DAsync(ptr, (size_t) ld * 8, t.data(), (size_t) cols * 8, (size_t) cols * 8, rows, cudaMemcpyHostToDevice, st));
        BB_CUDA(cudaStreamSynchronize(st));
        if (field == BBGPU_CMA_A || h->xnes) return BBGPU_OK;
        scale_basis_kernel<<<n, 256, 0, st>>>(h->Bt, h->D, h->BDt, h->BDit, n, h->ldn);
        BB_TRY(launched(h->ctx, "scale_basis_kernel"));
        BB_CUDA(cudaStreamSynchronize(st));
        return BBGPU_OK;
    }
    BB_CUDA(cudaMemcpy2DAsync(ptr, (size_t) ld * 8, h_src, (size_t) cols * 8, (size_t) cols * 8, rows, cudaMemcpyHostToDevice, st));
    BB_CUDA(cudaStreamSynchronize(st));
    if (field == BBGPU_CMA_D && h->full) {
        scale_basis_kernel<<<h->n, 256, 0, st>>>(h->Bt, h->D, h->BDt, h->BDit, h->n, h->ldn);
        BB_TRY(launched(h->ctx, "scale_basis_kernel"));
        BB_CUDA(cudaStreamSynchronize(st));
    }
    return BBGPU_OK;
}

int bbgpu_cma_get(bbgpu_cma *h, int field, double *h_dst, int64_t count) {
    BB_ARG(h && h_dst);
    double *ptr;
    int64_t rows, cols, ld;
    BB_TRY(field_info(h, field, &ptr, &rows, &cols, &ld));
    BB_ARG(count == rows * cols);
    cudaStream_t st = h->ctx->stream;
    BB_CUDA(cudaMemcpy2DAsync(h_dst, (size_t) cols * 8, ptr, (size_t) ld * 8, (size_t) cols * 8, rows, cudaMemcpyDeviceToHost, st));
    BB_CUDA(cudaStreamSynchronize(st));
    if (field == BBGPU_CMA_B || field == BBGPU_CMA_A) {
        const int n = h->n;
        for (int i = 0; i < n; i++)
            for (int j = i + 1; j < n; j++) std::swap(h_dst[(size_t) i * n + j], h_dst[(size_t) j * n + i]);
    }
    return BBGPU_OK;
}

int bbgpu_cma_get_rank(bbgpu_cma *h, int32_t *h_idx, int count) {
    BB_ARG(h && h_idx && count == h->lambda);
    BB_CUDA(cudaMemcpyAsync(h_idx, h->rank_idx, (size_t) count * 4, cudaMemcpyDeviceToHost, h->ctx->stream));
    BB_CUDA(cudaStreamSynchronize(h->ctx->stream));
    return BBGPU_OK;
}

int bbgpu_cma_get_rows(bbgpu_cma *h, int field, int64_t row0, int64_t rows, double *h_dst) {
    BB_ARG(h && h_dst && field != BBGPU_CMA_B && field != BBGPU_CMA_A);
    double *ptr;
    int64_t nrows, cols, ld;
    BB_TRY(field_info(h, field, &ptr, &nrows, &cols, &ld));
    BB_ARG(row0 >= 0 && rows >= 0 && row0 + rows <= nrows);
    cudaStream_t st = h->ctx->stream;
    BB_CUDA(cudaMemcpy2DAsync(h_dst, (size_t) cols * 8, ptr + row0 * ld, (size_t) ld * 8, (size_t) cols * 8, rows, cudaMemcpyDeviceToHost, st));
    BB_CUDA(cudaStreamSynchronize(st));
    return BBGPU_OK;
}

int bbgpu_cma_device_ptr(bbgpu_cma *h, int field, void **d_ptr, int64_t *ld) {
    BB_ARG(h && d_ptr && ld);
    double *ptr;
    int64_t rows, cols;
    BB_TRY(field_info(h, field, &ptr, &rows, &cols, ld));
    *d_ptr = ptr;
    return BBGPU_OK;
}

int bbgpu_cma_sample(bbgpu_cma *h, double sigma, uint64_t generation, const double *h_z) {
    BB_ARG(h && !h->lm);
    if (h_z != nullptr) { // lambda*n draws; a sharded handle consumes its own candidates' rows
        BB_CUDA(cudaMemcpy2DAsync(h->Z, (size_t) h->ldn * 8, h_z + (size_t) h->row0 * h->n, (size_t) h->n * 8, (size_t) h->n * 8,
                h->lam_loc, cudaMemcpyHostToDevice, h->ctx->stream));
        h->zgen[h->zcur] = -1;
        return do_sample(h, sigma);
    }
    BB_TRY(do_philox(h, generation));
    BB_TRY(do_sample(h, sigma));
    return prefetch_philox(h, generation);
}

int bbgpu_cma_sample_dev(bbgpu_cma *h, double sigma, const double *d_z) {
    BB_ARG(h && d_z && !h->lm);
    BB_CUDA(cudaMemcpy2DAsync(h->Z, (size_t) h->ldn * 8, d_z + (size_t) h->row0 * h->n, (size_t) h->n * 8, (size_t) h->n * 8,
            h->lam_loc, cudaMemcpyDeviceToDevice, h->ctx->stream));
    h->zgen[h->zcur] = -1;
    return do_sample(h, sigma);
}

int bbgpu_cma_evaluate(bbgpu_cma *h, bbgpu_objective *obj) {
    BB_ARG(h && obj && obj->n == h->n);
    if (h->shards > 1) { // local slice, then (with a communicator) the fitness all-gather
        BB_TRY(shard_evaluate(h, obj));
        return h->comm ? shard_exchange_fitness(h) : BBGPU_OK;
    }
    ProfScope prof(h->ctx, BBGPU_PROF_EVAL);
    return evaluate(h->ctx, obj, h->X, h->ldn, h->lambda, h->fit);
}

int bbgpu_cma_rank(bbgpu_cma *h, int ik) {
    BB_ARG(h && ik >= 0);
    return do_rank(h, ik);
}

int bbgpu_cma_update(bbgpu_cma *h, double sigma, int fev) {
    BB_ARG(h && sigma > 0.0 && !h->lm && !h->xnes);
    if (h->shards > 1) {
        if (!h->comm) return fail(BBGPU_ERR_ARG, "sharded handle without a communicator: use bbgpu_cma_shard_partial / _finish");
        BB_TRY(shard_partial(h));
        BB_TRY(shard_exchange_partials(h));
        return shard_finish(h, sigma, fev);
    }
    return do_update(h, sigma, fev);
}

int bbgpu_cma_attach_comm(bbgpu_cma *h, bbgpu_comm *comm) {
    BB_ARG(h);
    if (comm) {
        if (h->shards <= 1) return fail(BBGPU_ERR_ARG, "the handle was not created with shard_count > 1");
        BB_ARG(comm_ctx(comm) == h->ctx && comm_world(comm) == h->shards && comm_rank(comm) == h->shard);
    }
    if (h->peer) {
        comm_peer_region_destroy(h->peer);
        h->peer = nullptr;
    }
    h->comm = comm;
    // exchange step 2 over peer memory when every rank can map every rank's buffer (BBGPU_SHARD_P2P=0: always NCCL)
    const char *env = getenv("BBGPU_SHARD_P2P");
    if (comm && !(env && env[0] == '0'))
        BB_TRY(comm_peer_region_create(comm, h->red, ((size_t) h->n * h->ldn + h->ldn) * 8, &h->peer));
    return BBGPU_OK;
}

int bbgpu_cma_shard_uses_peer_memory(bbgpu_cma *h) { return h && h->peer ? 1 : 0; }

int bbgpu_cma_shard_partial(bbgpu_cma *h) {
    BB_ARG(h && h->shards > 1);
    return shard_partial(h);
}

int bbgpu_cma_shard_finish(bbgpu_cma *h, double sigma, int fev) {
    BB_ARG(h && h->shards > 1 && sigma > 0.0);
    return shard_finish(h, sigma, fev);
}

// one candidate by GLOBAL index; on a sharded handle with a communicator every rank must call this
// (the owner broadcasts the row)
int bbgpu_cma_get_candidate(bbgpu_cma *h, int index, double *h_dst) {
    BB_ARG(h && h_dst && index >= 0 && index < h->lambda);
    cudaStream_t st = h->ctx->stream;
    const int owner = index / h->lam_loc, loc = index - owner * h->lam_loc;
    if (h->shards > 1 && h->comm) {
        if (owner == h->shard) BB_CUDA(cudaMemcpyAsync(h->tmpn, h->X + (int64_t) loc * h->ldn, (size_t) h->n * 8, cudaMemcpyDeviceToDevice, st));
        BB_TRY(comm_broadcast(h->comm, h->tmpn, (size_t) h->n, owner));
        BB_CUDA(cudaMemcpyAsync(h_dst, h->tmpn, (size_t) h->n * 8, cudaMemcpyDeviceToHost, st));
    } else {
        if (owner != h->shard) return fail(BBGPU_ERR_ARG, "candidate %d lives on shard %d", index, owner);
        BB_CUDA(cudaMemcpyAsync(h_dst, h->X + (int64_t) loc * h->ldn, (size_t) h->n * 8, cudaMemcpyDeviceToHost, st));
    }
    BB_CUDA(cudaStreamSynchronize(st));
    return BBGPU_OK;
}

int bbgpu_cma_eig(bbgpu_cma *h, const double *h_B, const double *h_D) {
    BB_ARG(h && h->full);
    BB_ARG((h_B == nullptr) == (h_D == nullptr));
    return do_eig(h, h_B, h_D);
}

// xNES: everything after the ranking (xnes.cpp:130-190); sigma is the value the points were sampled with
int bbgpu_cma_xnes_update(bbgpu_cma *h, double sigma, double etamu, double etab, double *h_gsigma) {
    BB_ARG(h && h->xnes && sigma > 0.0 && h_gsigma);
    bbgpu_ctx *ctx = h->ctx;
    cudaStream_t st = ctx->stream;
    const int n = h->n, np = h->lambda;
    {
        ProfScope prof(ctx, BBGPU_PROF_UPDATE_MISC);
        xnes_gdelta_kernel<<<ceil_div(n, 256), 256, 0, st>>>(h->dm, h->Z, h->ldn, n, np, h->rank_idx, h->weights);
        BB_TRY(launched(ctx, "xnes_gdelta_kernel"));
        dim3 tgrid(ceil_div(n, 32), ceil_div(n, 32));
        xnes_gm_kernel<<<tgrid, 256, 0, st>>>(h->C, h->Z, h->ldn, n, np, h->rank_idx, h->weights);
        BB_TRY(launched(ctx, "xnes_gm_kernel"));
        xnes_trace_kernel<<<1, 1024, 0, st>>>(h->C, h->ldn, n, h->dscal);
        BB_TRY(launched(ctx, "xnes_trace_kernel"));
        dim3 sgrid(ceil_div(n, 256), n);
        xnes_scale_kernel<<<sgrid, 256, 0, st>>>(h->C, h->ldn, n, h->dscal, 0.5 * etab);
        BB_TRY(launched(ctx, "xnes_scale_kernel"));
        xnes_mean_kernel<<<ceil_div(n, 256), 256, 0, st>>>(h->xmean, h->Bt, h->ldn, n, h->dm, etamu * sigma);
        BB_TRY(launched(ctx, "xnes_mean_kernel"));
        BB_CUDA(cudaMemcpyAsync(h_gsigma, h->dscal, 8, cudaMemcpyDeviceToHost, st));
    }
    {
        ProfScope prof(ctx, BBGPU_PROF_EIG);
        BB_TRY(eig_core(h, h->C, h->BDt, h->W)); // rows of BDt: eigenvectors of C
        xnes_info_kernel<<<1, 32, 0, st>>>(h->dev_info, h->d_report);
        BB_TRY(launched(ctx, "xnes_info_kernel"));
        xnes_expscale_kernel<<<n, 256, 0, st>>>(h->BDt, h->W, h->BDit, n, h->ldn);
        BB_TRY(launched(ctx, "xnes_expscale_kernel"));
        // expm(C)[i][l] = sum_j V[j][i] (exp(w_j) V[j][l])
        GemmArgs g {};
        g.A = h->BDt; g.lda = h->ldn; g.B = h->BDit; g.ldb = h->ldn; g.M = n; g.N = n; g.K = n; g.C = h->isC; g.ldc = h->ldn;
        BB_TRY((launch_gemm<2, EPI_STORE>(ctx, g)));
        // B <- B expm(C)  <=>  B^T <- expm(C) B^T (expm(C) is symmetric)
        GemmArgs q {};
        q.A = h->isC; q.lda = h->ldn; q.B = h->Bt; q.ldb = h->ldn; q.M = n; q.N = n; q.K = n; q.C = h->BDt; q.ldc = h->ldn;
        BB_TRY((launch_gemm<1, EPI_STORE>(ctx, q)));
        std::swap(h->Bt, h->BDt);
    }
    BB_CUDA(cudaStreamSynchronize(st));
    return BBGPU_OK;
}

// LM kind: samplePopulation (lm_cmaes.cpp:91-127)
int bbgpu_cma_lm_sample(bbgpu_cma *h, double sigma, uint64_t generation, int memlen, const int32_t *h_jarr, const double *h_z,
        const int32_t *h_i0) {
    BB_ARG(h && h->lm && memlen >= 0 && memlen <= h->lm_mem && (memlen == 0 || h_jarr != nullptr));
    BB_ARG((h_z == nullptr) == (h_i0 == nullptr));
    bbgpu_ctx *ctx = h->ctx;
    cudaStream_t st = ctx->stream;
    h->last_sigma = sigma;
    if (memlen > 0) BB_CUDA(cudaMemcpyAsync(h->lm_jarr, h_jarr, (size_t) memlen * 4, cudaMemcpyHostToDevice, st));
    if (h_z != nullptr) {
        BB_CUDA(cudaMemcpy2DAsync(h->Zp, (size_t) h->ldn * 8, h_z, (size_t) h->n * 8, (size_t) h->n * 8, h->lm_pairs, cudaMemcpyHostToDevice, st));
        BB_CUDA(cudaMemcpyAsync(h->lm_i0, h_i0, (size_t) h->lm_pairs * 4, cudaMemcpyHostToDevice, st));
    }
    ProfScope prof(ctx, BBGPU_PROF_SEP_SAMPLE);
    static const bool one_cta = [] { const char *e = getenv("BBGPU_LM_SAMPLE"); return e && e[0] == '1'; }(); // 1: a pair per CTA
    if (h->lm_slabs == 1 || one_cta) {
        lm_sample_kernel<<<h->lm_pairs, 256, 0, st>>>(h->X, h->Zp, h->ldn, h->n, h->lambda, h->xmean, sigma, h->PCmat, h->Vmat, h->lmb,
                h->lm_jarr, memlen, h->lm_i0, h_z != nullptr, h->cfg.lm_rademacher, h->cfg.lm_usenew, std::sqrt(1.0 - h->cfg.c1),
                h->cfg.seed, generation);
        return launched(ctx, "lm_sample_kernel");
    }
    const int slab_len = (int) round_up(ceil_div(h->n, h->lm_slabs), 2);
    const dim3 grid(h->lm_pairs, h->lm_slabs);
    lm_dots_kernel<<<grid, 256, 0, st>>>(h->Zp, h->ldn, h->n, slab_len, h->Vmat, h->lm_jarr, memlen, h->lm_i0, h_z != nullptr,
            h->cfg.lm_rademacher, h->cfg.lm_usenew, h->cfg.seed, generation, h->lm_part, h->lm_part_ld);
    BB_TRY(launched(ctx, "lm_dots_kernel"));
    lm_apply_kernel<<<grid, 256, 0, st>>>(h->X, h->Zp, h->ldn, h->n, h->lambda, slab_len, h->xmean, sigma, h->PCmat, h->lmb, h->lm_jarr,
            memlen, h->lm_i0, std::sqrt(1.0 - h->cfg.c1), h->lm_part, h->lm_part_ld);
    return launched(ctx, "lm_apply_kernel");
}

// LM kind: updateDistribution up to updateSigma (lm_cmaes.cpp:129-178)
int bbgpu_cma_lm_update(bbgpu_cma *h, double sigma, int do_memory, int imin, int memlen, const int32_t *h_jarr) {
    BB_ARG(h && h->lm && sigma > 0.0 && memlen >= 0 && memlen <= h->lm_mem);
    BB_ARG(!do_memory || (memlen >= 1 && imin >= 0 && imin <= memlen && h_jarr != nullptr));
    bbgpu_ctx *ctx = h->ctx;
    const bbgpu_cma_config &c = h->cfg;
    const int n = h->n;
    cudaStream_t st = ctx->stream;
    const double *lo = c.bound ? h->lower : nullptr, *hi = c.bound ? h->upper : nullptr;
    ProfScope prof(ctx, BBGPU_PROF_UPDATE_MISC);
    BB_CUDA(cudaMemcpyAsync(h->xold, h->xmean, (size_t) n * 8, cudaMemcpyDeviceToDevice, st));
    if (do_memory) BB_CUDA(cudaMemcpyAsync(h->lm_jarr, h_jarr, (size_t) memlen * 4, cudaMemcpyHostToDevice, st));
    dim3 pgrid(ceil_div(n, 256), h->splits);
    recombine_partial_kernel<false><<<pgrid, 256, 0, st>>>(h->part, h->X, h->ldn, n, h->mu, h->splits, h->rank_idx, h->weights, h->xold,
            sigma);
    BB_TRY(launched(ctx, "recombine_partial_kernel"));
    recombine_final_kernel<false><<<ceil_div(n, 64), 256, 0, st>>>(h->part, n, h->splits, h->xmean, h->xold, h->dm, 1.0, sigma, lo, hi,
            false, h->ps, nullptr, 0.0, 0.0, nullptr);
    BB_TRY(launched(ctx, "recombine_final_kernel"));
    const double ccc = std::sqrt(c.cc * (2.0 - c.cc) * c.mueff);
    // the rebuild of rows imin.. as a wavefront over the SMs (lm_rebuild_kernel) unless BBGPU_LM_REBUILD=1 keeps it in the one CTA
    static const bool one_cta = [] { const char *e = getenv("BBGPU_LM_REBUILD"); return e && e[0] == '1'; }();
    const int rows = do_memory ? memlen - imin : 0;
    const bool wave = do_memory && !one_cta && rows >= 2;
    lm_memory_kernel<<<1, 1024, 0, st>>>(n, h->ldn, h->pc, h->xmean, h->xold, sigma, c.cc, ccc, c.c1, std::sqrt(1.0 - c.c1),
            wave ? 2 : do_memory, imin, memlen, h->lm_jarr, h->PCmat, h->Vmat, h->lmb, h->lmd, h->tmpn);
    BB_TRY(launched(ctx, "lm_memory_kernel"));
    if (wave) {
        BB_CUDA(cudaMemsetAsync(h->lm_flags, 0, ((size_t) memlen + 1) * sizeof(int), st));
        int nn = n, im = imin, ml = memlen;
        int64_t ldn = h->ldn;
        double c1 = c.c1, s1 = std::sqrt(1.0 - c.c1);
        const int *jarr = h->lm_jarr;
        const double *PC = h->PCmat;
        double *V = h->Vmat, *bb_ = h->lmb, *dd_ = h->lmd, *tg = h->lm_tglob;
        int *flags = h->lm_flags;
        void *args[] = { &nn, &ldn, &c1, &s1, &im, &ml, &jarr, &PC, &V, &bb_, &dd_, &flags, &tg };
        BB_CUDA(cudaLaunchCooperativeKernel((const void*) lm_rebuild_kernel, dim3(std::min(rows, ctx->sm_count)), dim3(1024), args, 0, st));
        BB_TRY(launched(ctx, "lm_rebuild_kernel"));
    }
    return BBGPU_OK;
}

// CholeskyCmaes::converged(), the O(lambda n) part (cholesky_cmaes.cpp:146-160): mean and M2 = sum (r_k - mean)^2 of
// the candidates' Euclidean norms
int bbgpu_cma_swarm_radius(bbgpu_cma *h, double *h_mean, double *h_m2) {
    BB_ARG(h && h->chol && h_mean && h_m2);
    cudaStream_t st = h->ctx->stream;
    row_norms_kernel<<<ceil_div(h->lambda, 8), 256, 0, st>>>(h->X, h->ldn, h->n, h->lambda, h->rnorm);
    BB_TRY(launched(h->ctx, "row_norms_kernel"));
    radius_stats_kernel<<<1, 1024, 0, st>>>(h->rnorm, h->lambda, h->rstat);
    BB_TRY(launched(h->ctx, "radius_stats_kernel"));
    double out[2];
    BB_CUDA(cudaMemcpyAsync(out, h->rstat, 16, cudaMemcpyDeviceToHost, st));
    BB_CUDA(cudaStreamSynchronize(st));
    *h_mean = out[0];
    *h_m2 = out[1];
    return BBGPU_OK;
}

int bbgpu_cma_probe(bbgpu_cma *h, double sigma, int iaxis) {
    BB_ARG(h && iaxis >= 0 && iaxis < h->n && !h->chol && !h->lm && !h->xnes);
    if (h->full) probe_kernel<false><<<1, 1024, 0, h->ctx->stream>>>(h->n, h->ldn, h->xmean, h->pc, h->C, h->D, h->Bt, sigma, iaxis, h->d_report);
    else probe_kernel<true><<<1, 1024, 0, h->ctx->stream>>>(h->n, h->ldn, h->xmean, h->pc, h->C, h->D, nullptr, sigma, iaxis, h->d_report);
    return launched(h->ctx, "probe_kernel");
}

int bbgpu_cma_fetch(bbgpu_cma *h, bbgpu_cma_report *out) {
    BB_ARG(h && out);
    BB_CUDA(cudaMemcpyAsync(h->h_report, h->d_report, sizeof(bbgpu_cma_report), cudaMemcpyDeviceToHost, h->ctx->stream));
    BB_CUDA(cudaStreamSynchronize(h->ctx->stream));
    *out = *h->h_report;
    if (out->eig_info != 0) return fail(BBGPU_ERR_SOLVER, "syevd reported info = %d", out->eig_info);
    return BBGPU_OK;
}

int bbgpu_cma_generation(bbgpu_cma *h, bbgpu_objective *obj, double sigma, uint64_t generation, int ik, int fev_after) {
    BB_ARG(h && obj && obj->n == h->n && !h->lm && !h->xnes);
    if (!h->full && !h->chol && obj->mode == 0 && obj->id != BBGPU_OBJ_ROSENBROCK) {
        // separable kind, registry objective: draws, sampling and evaluation fused into one pass over X
        bbgpu_ctx *ctx = h->ctx;
        const double *lo = h->cfg.bound ? h->lower : nullptr, *hi = h->cfg.bound ? h->upper : nullptr;
        const int grid = h->lam_loc;
        h->last_sigma = sigma;
        h->zgen[0] = h->zgen[1] = -1; // Z is not materialised on this path
        obj->evals += h->lam_loc;
        {
            ProfScope prof(ctx, BBGPU_PROF_SEP_SAMPLE);
#define BB_SEP_CASE(ID) case ID: sep_sample_eval_kernel<ID><<<grid, 256, 0, ctx->stream>>>(h->X, h->ldn, h->lam_loc, h->n, h->row0, \
            h->xmean, h->D, sigma, lo, hi, obj->d_aux, h->cfg.seed, generation, h->fit); break;
            switch (obj->id) {
            BB_SEP_CASE(BBGPU_OBJ_SPHERE) BB_SEP_CASE(BBGPU_OBJ_ELLIPSOID) BB_SEP_CASE(BBGPU_OBJ_RASTRIGIN) BB_SEP_CASE(BBGPU_OBJ_ACKLEY)
            default: return fail(BBGPU_ERR_ARG, "unknown registry objective %d", obj->id);
            }
#undef BB_SEP_CASE
            BB_TRY(launched(ctx, "sep_sample_eval_kernel"));
        }
        BB_TRY(do_rank(h, ik));
        return do_update(h, sigma, fev_after);
    }
    BB_TRY(do_philox(h, generation));
    BB_TRY(do_sample(h, sigma));
    BB_TRY(prefetch_philox(h, generation));
    if (h->shards > 1) {
        // sample/evaluate the local candidates -> fitness all-gather -> replicated ranking -> local rank-mu
        // partial -> all-reduce -> replicated paths and covariance; no host round trip
        if (!h->comm) return fail(BBGPU_ERR_ARG, "sharded handle without a communicator");
        BB_TRY(shard_evaluate(h, obj));
        BB_TRY(shard_exchange_fitness(h));
        BB_TRY(do_rank(h, ik));
        BB_TRY(shard_partial(h));
        BB_TRY(shard_exchange_partials(h));
        return shard_finish(h, sigma, fev_after);
    }
    {
        ProfScope prof(h->ctx, BBGPU_PROF_EVAL);
        BB_TRY(evaluate(h->ctx, obj, h->X, h->ldn, h->lambda, h->fit));
    }
    BB_TRY(do_rank(h, ik));
    return do_update(h, sigma, fev_after);
}

// ---- standalone kernels, host-buffer form -----------------------------------------------
static int upload_padded(bbgpu_ctx *ctx, const double *h, int rows, int cols, double **d, int64_t *ld) {
    *ld = round_up(cols, 8);
    BB_TRY(dev_alloc(d, (size_t) rows * *ld));
    BB_CUDA(cudaMemsetAsync(*d, 0, (size_t) rows * *ld * 8, ctx->stream));
    if (h) BB_CUDA(cudaMemcpy2DAsync(*d, (size_t) *ld * 8, h, (size_t) cols * 8, (size_t) cols * 8, rows, cudaMemcpyHostToDevice, ctx->stream));
    return BBGPU_OK;
}

int bbgpu_dgemm_host(bbgpu_ctx *ctx, int M, int N, int K, const double *h_A, const double *h_B, double *h_C) {
    BB_ARG(ctx && h_A && h_B && h_C && M > 0 && N > 0 && K > 0);
    double *dA = nullptr, *dB = nullptr, *dC = nullptr;
    int64_t lda, ldb, ldc;
    int rc = upload_padded(ctx, h_A, M, K, &dA, &lda);
    if (rc == BBGPU_OK) rc = upload_padded(ctx, h_B, K, N, &dB, &ldb);
    if (rc == BBGPU_OK) rc = upload_padded(ctx, nullptr, M, N, &dC, &ldc);
    if (rc == BBGPU_OK) {
        GemmArgs g {};
        g.A = dA; g.lda = lda; g.B = dB; g.ldb = ldb; g.M = M; g.N = N; g.K = K; g.C = dC; g.ldc = ldc;
        rc = launch_gemm<1, EPI_STORE>(ctx, g);
    }
    if (rc == BBGPU_OK && cudaMemcpy2DAsync(h_C, (size_t) N * 8, dC, (size_t) ldc * 8, (size_t) N * 8, M, cudaMemcpyDeviceToHost,
            ctx->stream) != cudaSuccess) rc = fail(BBGPU_ERR_CUDA, "copy back failed");
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess && rc == BBGPU_OK) rc = fail(BBGPU_ERR_CUDA, "sync failed: %s", cudaGetErrorString(cudaGetLastError()));
    cudaFree(dA); cudaFree(dB); cudaFree(dC);
    return rc;
}

int bbgpu_dsyrk_host(bbgpu_ctx *ctx, int n, int K, double alpha, const double *h_Y, const double *h_coef, double *h_C) {
    BB_ARG(ctx && h_Y && h_coef && h_C && n > 0 && K > 0);
    double *dY = nullptr, *dC = nullptr, *dcoef = nullptr, *dzero = nullptr;
    int64_t ldy, ldc;
    int rc = upload_padded(ctx, h_Y, K, n, &dY, &ldy);
    if (rc == BBGPU_OK) rc = upload_padded(ctx, h_C, n, n, &dC, &ldc);
    if (rc == BBGPU_OK) rc = dev_alloc(&dcoef, (size_t) round_up(K, 32) + 32);
    if (rc == BBGPU_OK) rc = dev_alloc(&dzero, (size_t) round_up(n, 8) + 8);
    if (rc == BBGPU_OK) {
        cudaMemsetAsync(dcoef, 0, ((size_t) round_up(K, 32) + 32) * 8, ctx->stream);
        cudaMemcpyAsync(dcoef, h_coef, (size_t) K * 8, cudaMemcpyHostToDevice, ctx->stream);
        cudaMemsetAsync(dzero, 0, ((size_t) round_up(n, 8) + 8) * 8, ctx->stream);
        GemmArgs g {};
        g.A = dY; g.lda = ldy; g.B = dY; g.ldb = ldy; g.M = n; g.N = n; g.K = K; g.kscale = dcoef; g.kscale_padded = true;
        g.C = dC; g.ldc = ldc; g.vec = dzero; g.s0 = alpha; g.s1 = 0.0; g.dscal = dzero;
        rc = launch_gemm<2, EPI_SYRK>(ctx, g);
    }
    if (rc == BBGPU_OK && cudaMemcpy2DAsync(h_C, (size_t) n * 8, dC, (size_t) ldc * 8, (size_t) n * 8, n, cudaMemcpyDeviceToHost,
            ctx->stream) != cudaSuccess) rc = fail(BBGPU_ERR_CUDA, "copy back failed");
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess && rc == BBGPU_OK) rc = fail(BBGPU_ERR_CUDA, "sync failed: %s", cudaGetErrorString(cudaGetLastError()));
    cudaFree(dY); cudaFree(dC); cudaFree(dcoef); cudaFree(dzero);
    return rc;
}

int bbgpu_rank_host(bbgpu_ctx *ctx, int count, const double *h_f, int32_t *h_idx) {
    BB_ARG(ctx && h_f && h_idx && count > 0);
    double *df = nullptr, *dv = nullptr;
    int *di = nullptr;
    int rc = dev_alloc(&df, (size_t) count);
    if (rc == BBGPU_OK) rc = dev_alloc(&dv, (size_t) count);
    if (rc == BBGPU_OK) rc = dev_alloc(&di, (size_t) count);
    if (rc == BBGPU_OK) {
        cudaMemcpyAsync(df, h_f, (size_t) count * 8, cudaMemcpyHostToDevice, ctx->stream);
        rc = rank_fitness(ctx, df, count, di, dv);
    }
    if (rc == BBGPU_OK) cudaMemcpyAsync(h_idx, di, (size_t) count * 4, cudaMemcpyDeviceToHost, ctx->stream);
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess && rc == BBGPU_OK) rc = fail(BBGPU_ERR_CUDA, "sync failed: %s", cudaGetErrorString(cudaGetLastError()));
    cudaFree(df); cudaFree(dv); cudaFree(di);
    return rc;
}

int bbgpu_philox_normal_host(bbgpu_ctx *ctx, uint64_t seed, uint64_t generation, int rows, int cols, double *h_z) {
    BB_ARG(ctx && h_z && rows > 0 && cols > 0);
    double *dz = nullptr;
    int64_t ld;
    int rc = upload_padded(ctx, nullptr, rows, cols, &dz, &ld);
    if (rc == BBGPU_OK) {
        const int64_t total = (int64_t) rows * ((cols + 1) / 2);
        philox_normal_kernel<<<ceil_div(total, 256), 256, 0, ctx->stream>>>(dz, ld, rows, cols, 0, seed, generation, 0u);
        rc = launched(ctx, "philox_normal_kernel");
    }
    if (rc == BBGPU_OK) cudaMemcpy2DAsync(h_z, (size_t) cols * 8, dz, (size_t) ld * 8, (size_t) cols * 8, rows, cudaMemcpyDeviceToHost, ctx->stream);
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess && rc == BBGPU_OK) rc = fail(BBGPU_ERR_CUDA, "sync failed: %s", cudaGetErrorString(cudaGetLastError()));
    cudaFree(dz);
    return rc;
}

int bbgpu_bench_dgemm(bbgpu_ctx *ctx, int M, int N, int K, int reps, float *ms_per_launch) {
    BB_ARG(ctx && ms_per_launch && M > 0 && N > 0 && K > 0 && reps > 0);
    double *dA = nullptr, *dB = nullptr, *dC = nullptr;
    int64_t lda, ldb, ldc;
    int rc = upload_padded(ctx, nullptr, M, K, &dA, &lda);
    if (rc == BBGPU_OK) rc = upload_padded(ctx, nullptr, K, N, &dB, &ldb);
    if (rc == BBGPU_OK) rc = upload_padded(ctx, nullptr, M, N, &dC, &ldc);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    if (rc == BBGPU_OK) {
        fill_kernel<<<ceil_div((int64_t) M * lda, 256), 256, 0, ctx->stream>>>(dA, (int) ((int64_t) M * lda), 1.0);
        fill_kernel<<<ceil_div((int64_t) K * ldb, 256), 256, 0, ctx->stream>>>(dB, (int) ((int64_t) K * ldb), 0.5);
        GemmArgs g {};
        g.A = dA; g.lda = lda; g.B = dB; g.ldb = ldb; g.M = M; g.N = N; g.K = K; g.C = dC; g.ldc = ldc;
        for (int i = 0; i < 3 && rc == BBGPU_OK; i++) rc = launch_gemm<1, EPI_STORE>(ctx, g);
        cudaEventRecord(e0, ctx->stream);
        for (int i = 0; i < reps && rc == BBGPU_OK; i++) rc = launch_gemm<1, EPI_STORE>(ctx, g);
        cudaEventRecord(e1, ctx->stream);
    }
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess && rc == BBGPU_OK) rc = fail(BBGPU_ERR_CUDA, "sync failed: %s", cudaGetErrorString(cudaGetLastError()));
    float ms = 0.f;
    if (rc == BBGPU_OK) cudaEventElapsedTime(&ms, e0, e1);
    *ms_per_launch = ms / reps;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(dA); cudaFree(dB); cudaFree(dC);
    return rc;
}

} // extern "C"
