/*
 * bbgpu.h -- C ABI of the B200 population engine (libbbgpu.so).
 *
 * This is the drop-in boundary for the reference's data-parallel hot path: the
 * per-generation sample -> evaluate -> rank -> update loop of the CMA-ES family (Vanilla,
 * Active, Separable, Cholesky, LM; xNES shares the handle) and the generation of the DE family
 * (JADE / L-SHADE: bbgpu_de_*, SaNSDE: bbgpu_sde_*, SSDE: bbgpu_ssde_*).  The reference has no FFI for this path (it is straight C++
 * behind pybind11), so each entry point below names the reference code it replaces
 * (paths relative to /root/reference).  The host classes in bboptpy_b200/host/ keep the
 * reference's MultivariateOptimizer interface (src/multivariate/multivariate.h:132-146) and
 * call only these functions; INTEGRATION.md shows the binding a maintainer would add.
 *
 * Conventions: plain pointers and sizes, no C++ or torch types; every function returns an
 * int status (0 = BBGPU_OK) and bbgpu_last_error() describes the last failure on the calling
 * thread; handles own their device memory; host pointers are borrowed for the call only;
 * one ctx <-> one device + one stream; not thread-safe per handle (like the reference,
 * src/random.hpp:168).  There is NO CPU fallback: without a CUDA device bbgpu_ctx_create
 * fails with BBGPU_ERR_NODEVICE and nothing else can be constructed.
 *
 * All matrices crossing this ABI are dense row-major fp64 with the reference's index
 * meaning (e.g. B[i*n+j] = component i of eigenvector j, cma/cmaes.cpp:65-80).
 */
#ifndef BBGPU_H_
#define BBGPU_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default) /* the library itself is built with -fvisibility=hidden */
#endif

#define BBGPU_VERSION 100

enum bbgpu_status {
    BBGPU_OK = 0,
    BBGPU_ERR_CUDA = 1,     /* a CUDA runtime / driver call failed */
    BBGPU_ERR_ARG = 2,      /* invalid argument */
    BBGPU_ERR_NODEVICE = 3, /* no usable CUDA device (no CPU fallback exists) */
    BBGPU_ERR_SOLVER = 4,   /* cuSOLVER failure / eigensolver did not converge */
    BBGPU_ERR_OBJECTIVE = 5 /* a user objective returned non-zero */
};

typedef struct bbgpu_ctx bbgpu_ctx;
typedef struct bbgpu_objective bbgpu_objective;
typedef struct bbgpu_cma bbgpu_cma;
typedef struct bbgpu_de bbgpu_de;
typedef struct bbgpu_cmab bbgpu_cmab;

int bbgpu_version(void);
const char *bbgpu_last_error(void);

/* ---- context: device + stream.  stream == NULL -> the ctx creates and owns a stream. */
int bbgpu_ctx_create(int device, void *cuda_stream, bbgpu_ctx **out);
int bbgpu_ctx_destroy(bbgpu_ctx *ctx);
int bbgpu_ctx_sync(bbgpu_ctx *ctx);
void *bbgpu_ctx_stream(bbgpu_ctx *ctx);
int bbgpu_ctx_device(bbgpu_ctx *ctx);
/* number of kernels this library has launched on ctx since creation (bench.py gpu_launches) */
int64_t bbgpu_ctx_launch_count(bbgpu_ctx *ctx);

/* ---- measurement helpers (bench.py): CUDA-event timers on the ctx stream, per-kernel device-time
 * accounting, and an L2 flush (writes a buffer larger than the 126 MB L2). */
enum bbgpu_profile_tag {
    BBGPU_PROF_SAMPLE_GEMM = 0, /* X = m + sigma Z (B D)^T                 */
    BBGPU_PROF_COV_SYRK = 1,    /* rank-mu / active covariance update      */
    BBGPU_PROF_NORM_GEMM = 2,   /* active: Mahalanobis norms               */
    BBGPU_PROF_EVAL = 3,
    BBGPU_PROF_RANK = 4,
    BBGPU_PROF_PHILOX = 5,
    BBGPU_PROF_EIG = 6,         /* syevd + C^-1/2 rebuild                  */
    BBGPU_PROF_UPDATE_MISC = 7, /* gather, recombination, paths            */
    BBGPU_PROF_SEP_SAMPLE = 8,
    BBGPU_PROF_DE_ROUND = 9,
    BBGPU_PROF_DE_MISC = 10,
    BBGPU_PROF_COMM = 11,       /* NCCL exchange steps of a sharded run    */
    BBGPU_PROF_CMAB_STEP = 12,  /* batched small-n engine: whole generations */
    BBGPU_PROF_TAGS = 13
};
int bbgpu_ctx_profile_enable(bbgpu_ctx *ctx, int on);
/* synchronises, then returns summed device milliseconds and number of timed launches for tag */
int bbgpu_ctx_profile_read(bbgpu_ctx *ctx, int tag, double *total_ms, int64_t *count);
int bbgpu_ctx_profile_reset(bbgpu_ctx *ctx);
int bbgpu_ctx_timer_start(bbgpu_ctx *ctx);
int bbgpu_ctx_timer_stop(bbgpu_ctx *ctx, float *ms); /* records, synchronises, returns elapsed */
int bbgpu_ctx_flush_l2(bbgpu_ctx *ctx);

/* ---- objectives.  Replaces the per-candidate callback `_f._f(const double*)`
 * (cma/base_cmaes.cpp:214-217, de/jade.cpp:159) by batched device evaluation. */
enum bbgpu_builtin_objective {
    BBGPU_OBJ_SPHERE = 0,
    BBGPU_OBJ_ELLIPSOID = 1,
    BBGPU_OBJ_ROSENBROCK = 2, /* README form, examples/multivariate.py:6-10 */
    BBGPU_OBJ_RASTRIGIN = 3,
    BBGPU_OBJ_ACKLEY = 4
};
/* user CUDA objective: evaluate `count` candidates stored row-major with leading dimension
 * ld (doubles) at d_X, write d_f[count]; must enqueue on `cuda_stream`; return 0 on success */
typedef int (*bbgpu_device_objective_fn)(const double *d_X, int64_t ld, int n, int count,
                                         double *d_f, void *cuda_stream, void *user);
/* batched host callback bridge (for callers that only have a CPU function, e.g. the Python
 * callable of py/multivariate_py.cpp:385-388): X and f are pinned host buffers */
typedef int (*bbgpu_host_objective_fn)(const double *h_X, int64_t ld, int n, int count,
                                       double *h_f, void *user);
int bbgpu_objective_builtin(bbgpu_ctx *ctx, int id, int n, bbgpu_objective **out);
int bbgpu_objective_device(bbgpu_ctx *ctx, bbgpu_device_objective_fn fn, void *user, int n,
                           bbgpu_objective **out);
int bbgpu_objective_host(bbgpu_ctx *ctx, bbgpu_host_objective_fn fn, void *user, int n,
                         bbgpu_objective **out);
int bbgpu_objective_destroy(bbgpu_objective *obj);
/* device-pointer evaluation (async on the ctx stream) and host-buffer convenience form */
int bbgpu_eval(bbgpu_ctx *ctx, bbgpu_objective *obj, const double *d_X, int64_t ld, int count,
               double *d_f);
int bbgpu_eval_host(bbgpu_ctx *ctx, bbgpu_objective *obj, const double *h_X, int count,
                    double *h_f);

/* ---- CMA-ES engine (Cmaes / ActiveCmaes / SepCmaes device state and phases) */
/* BBGPU_CMA_CHOLESKY: cma/cholesky_cmaes.cpp -- the factor A (C = A A^T) instead of C / B / D; sampling x = m + sigma A z on
 * the same tensor-core GEMM, the mu + 1 sequential rank-one updates of the factor (:85-90,163-177) as one SYRK + one blocked
 * Cholesky factorisation (the factor is unique), no eigendecomposition at all */
/* BBGPU_CMA_LM: cma/lm_cmaes.cpp (Loshchilov 2014/2015) -- the covariance is implicit in `memory` direction vectors: rows of
 * PCMAT / VMAT with the coefficients LMB / LMD; mirrored sampling of lambda/2 base vectors (bbgpu_cma_lm_sample), the memory
 * bookkeeping (which rows are live, in which order: `jarr`) stays with the host class */
/* BBGPU_CMA_XNES: nes/xnes.cpp -- the same handle with lambda = mu = np points, `weights` = the np rank utilities (they sum to
 * zero), BBGPU_CMA_B = the transform B (x = mu + sigma B z) and BBGPU_CMA_XMEAN = mu; sample / evaluate / rank as for the
 * other kinds, then bbgpu_cma_xnes_update */
enum bbgpu_cma_kind { BBGPU_CMA_FULL = 0, BBGPU_CMA_ACTIVE = 1, BBGPU_CMA_SEP = 2, BBGPU_CMA_CHOLESKY = 3, BBGPU_CMA_LM = 4,
                      BBGPU_CMA_XNES = 5 };

typedef struct bbgpu_cma_config {
    int kind;       /* bbgpu_cma_kind */
    int n;          /* dimension */
    int lambda;     /* population size */
    int mu;         /* parents = lambda/2 (cma/base_cmaes.cpp:71) */
    int bound;      /* clip samples and mean to [lower,upper] (cma/cmaes.cpp:74-77,92-94) */
    int shard_rank; /* candidate sharding for one large-lambda run: this GPU owns candidates */
    int shard_count;/*   [rank*lambda/count, (rank+1)*lambda/count); 0/1 = unsharded */
    uint64_t seed;  /* Philox key for free-running sampling */
    /* strategy constants computed by the host class exactly as the reference does
     * (cma/base_cmaes.cpp:94-117, cma/active_cmaes.cpp:48-60, cma/sep_cmaes.cpp:44-58) */
    double cc, cs, c1, cmu, mueff, chi;
    double cneg, alphaold, cm; /* active only */
    double ccov;               /* sep only */
    /* LM only (cma/lm_cmaes.cpp:42-66): number of direction vectors, Rademacher (1) or Gaussian (0) base draws, the 2015
     * variant with a random subset of the vectors per sample (algorithm 6) */
    int lm_memory, lm_rademacher, lm_usenew;
} bbgpu_cma_config;

enum bbgpu_cma_field {
    BBGPU_CMA_XMEAN = 0, /* n */
    BBGPU_CMA_XOLD = 1,  /* n */
    BBGPU_CMA_PC = 2,    /* n */
    BBGPU_CMA_PS = 3,    /* n */
    BBGPU_CMA_C = 4,     /* full/active: n*n (lower triangle authoritative, cma/cmaes.cpp:127);
                            sep: n */
    BBGPU_CMA_B = 5,     /* n*n, B[i*n+j] = component i of eigenvector j */
    BBGPU_CMA_D = 6,     /* n, sqrt of eigenvalues ascending (sep: sqrt(c_i)) */
    BBGPU_CMA_INVSQRTC = 7, /* n*n */
    BBGPU_CMA_ARX = 8,   /* lambda*n, candidate-major (cma/base_cmaes.h:61) */
    BBGPU_CMA_Z = 9,     /* lambda*n, the standard-normal draws of the last sample() */
    BBGPU_CMA_FITNESS = 10, /* lambda, fitness by candidate index */
    BBGPU_CMA_RANKVAL = 11, /* lambda, fitness ascending */
    BBGPU_CMA_YCOEFF = 12,  /* mu (active) */
    BBGPU_CMA_WEIGHTS = 13, /* mu */
    /* candidate-sharded handles only (the two buffers the exchange steps move) */
    BBGPU_CMA_NORMS = 14,     /* lambda, ||C^-1/2 (x_k - m)||^2 by candidate index (active) */
    BBGPU_CMA_SHARD_RED = 15, /* n*ld + ld: this GPU's rank-mu partial (lower triangle, row pitch
                                 ld = ceil(n/8)*8) followed by its partial sum_k w_k x_k */
    BBGPU_CMA_A = 16,         /* Cholesky kind: n*n, the lower-triangular factor `_a` (cma/cholesky_cmaes.h:36) */
    BBGPU_CMA_PCMAT = 17,     /* LM kind: memory*n, `_pcmat` (cma/lm_cmaes.h:44) */
    BBGPU_CMA_VMAT = 18,      /* LM kind: memory*n, `_vmat` */
    BBGPU_CMA_LMB = 19,       /* LM kind: memory, `_b` */
    BBGPU_CMA_LMD = 20        /* LM kind: memory, `_d` */
};

/* everything the host class needs back per generation; filled by bbgpu_cma_fetch */
typedef struct bbgpu_cma_report {
    int32_t ibw[4]; /* indices of best, 2nd best, 2nd worst, worst (cma/base_cmaes.cpp:222-229) */
    double ybw[4];
    double f_kth;   /* fitness[_ik] for the flat-fitness test and the history ring (:183,:198) */
    double pslen;   /* ||p_sigma||_2 after the path update (dnrm2 in cma/base_cmaes.cpp:179) */
    int32_t hsig;
    int32_t eig_info;     /* 0 = ok (Cholesky kind: j > 0 = pivot j of the factorisation was not positive) */
    double d_min, d_max;  /* D[0], D[n-1] after the last decomposition (sep: c-ordered ends) */
    double tolx_max;      /* max_i max(pc_i, sqrt(C_ii))        (cma/cmaes.cpp:184-194) */
    int32_t noeffect_axis;/* 1 if NoEffectAxis holds            (cma/cmaes.cpp:208-221) */
    int32_t noeffect_coor;/* 1 if NoEffectCoor holds            (cma/cmaes.cpp:223-229) */
    double eig_resid;     /* max |B D^2 B^T - C| / max|C| when a replayed basis is validated */
} bbgpu_cma_report;

/* weights[mu]; lower/upper[n] may be NULL when bound == 0 */
int bbgpu_cma_create(bbgpu_ctx *ctx, const bbgpu_cma_config *cfg, const double *h_weights,
                     const double *h_lower, const double *h_upper, bbgpu_cma **out);
int bbgpu_cma_destroy(bbgpu_cma *h);
/* reset to the reference's init state: m = guess, C = B = C^-1/2 = I, D = 1, paths = 0
 * (cma/base_cmaes.cpp:119-124, cma/cmaes.cpp:52-60).  keep_offdiag != 0 reproduces the
 * reference's re-init quirk (only the diagonals of B and C are reset; cma/cmaes.cpp:53-60) */
int bbgpu_cma_reset(bbgpu_cma *h, const double *h_guess, int keep_offdiag);
int bbgpu_cma_set(bbgpu_cma *h, int field, const double *h_src, int64_t count);
int bbgpu_cma_get(bbgpu_cma *h, int field, double *h_dst, int64_t count);
int bbgpu_cma_get_rank(bbgpu_cma *h, int32_t *h_idx, int count);
/* rows [row0, row0+rows) of a matrix field (e.g. one candidate of ARX: bestSolution(),
 * cma/base_cmaes.cpp:232-238) */
int bbgpu_cma_get_rows(bbgpu_cma *h, int field, int64_t row0, int64_t rows, double *h_dst);
/* device pointer + leading dimension of a field (for user kernels / collectives) */
int bbgpu_cma_device_ptr(bbgpu_cma *h, int field, void **d_ptr, int64_t *ld);

/* sample(): cma/cmaes.cpp:65-80, cma/sep_cmaes.cpp:69-79.
 * h_z == NULL: draws come from the device Philox stream (key = cfg.seed, counter =
 * (generation, candidate, coordinate)).  h_z != NULL: DRAW REPLAY -- lambda*n standard
 * normals in the reference's draw order (candidate-major) are consumed instead. */
int bbgpu_cma_sample(bbgpu_cma *h, double sigma, uint64_t generation, const double *h_z);
/* same with the draws already on the device (ld = n) */
int bbgpu_cma_sample_dev(bbgpu_cma *h, double sigma, const double *d_z);
/* evaluateAndSortPopulation(): cma/base_cmaes.cpp:211-230.  Ties are ordered by index. */
int bbgpu_cma_evaluate(bbgpu_cma *h, bbgpu_objective *obj);
int bbgpu_cma_rank(bbgpu_cma *h, int ik);
/* updateDistribution() up to and excluding updateSigma()/eigenDecomposition():
 * cma/cmaes.cpp:82-142, cma/active_cmaes.cpp:71-161, cma/sep_cmaes.cpp:81-131.
 * fev = evaluation count AFTER this generation (the hsig exponent, cma/cmaes.cpp:109). */
int bbgpu_cma_update(bbgpu_cma *h, double sigma, int fev);
/* eigenDecomposition() body, cma/cmaes.cpp:236-282 (the host applies the gate :233).
 * h_B/h_D == NULL: cuSOLVER syevd + clamps + D = sqrt + C^-1/2 = B D^-1 B^T on the device.
 * h_B/h_D != NULL: EIGENBASIS REPLAY -- adopt the given basis (validated against C, see
 * report.eig_resid) and rebuild C^-1/2 from it. */
int bbgpu_cma_eig(bbgpu_cma *h, const double *h_B, const double *h_D);
/* quantities for converged(): cma/cmaes.cpp:184-229, cma/sep_cmaes.cpp:170-205 */
int bbgpu_cma_probe(bbgpu_cma *h, double sigma, int iaxis);
/* xNES, everything after the ranking (nes/xnes.cpp:130-190): natural-gradient terms G_delta / G_M / G_sigma from the ranked
 * draws, mu += eta_mu sigma B G_delta, B <- B expm(eta_B (G_M - G_sigma I) / 2) with the reference's eigendecomposition-based
 * matrix exponential (:227-253) as two tensor-core GEMMs around the eigensolver.  Returns G_sigma (the host applies
 * sigma *= exp(eta_sigma G_sigma / 2), :171); synchronises. */
int bbgpu_cma_xnes_update(bbgpu_cma *h, double sigma, double etamu, double etab, double *h_gsigma);
/* LM kind, samplePopulation() (cma/lm_cmaes.cpp:91-127): candidates 2p and 2p+1 are m +- sigma Az_p, Az_p = the base vector
 * z_p pushed through live memory rows jarr[i0_p .. memlen-1] (algorithm 3; the inner products use the ORIGINAL z_p, :117).
 * h_z / h_i0 == NULL: base vectors (Rademacher or Gaussian) and subset starts from the Philox stream; otherwise DRAW REPLAY:
 * ceil(lambda/2)*n base values and ceil(lambda/2) subset starts in the reference's order.  No box clipping (:124). */
int bbgpu_cma_lm_sample(bbgpu_cma *h, double sigma, uint64_t generation, int memlen, const int32_t *h_jarr,
                        const double *h_z, const int32_t *h_i0);
/* LM kind, updateDistribution() up to updateSigma() (cma/lm_cmaes.cpp:129-178): mean, p_c, and -- when do_memory
 * (it % t == 0) -- p_c stored as row jarr[memlen-1] of PCMAT and rows jarr[imin .. memlen-1] of VMAT / LMB / LMD rebuilt
 * (algorithm 4).  jarr / memlen / imin AFTER the host's updateSet() (algorithm 5, :274-306). */
int bbgpu_cma_lm_update(bbgpu_cma *h, double sigma, int do_memory, int imin, int memlen, const int32_t *h_jarr);
/* Cholesky kind, converged() (cma/cholesky_cmaes.cpp:146-160): mean and M2 = sum_k (r_k - mean)^2 of the candidates'
 * Euclidean norms r_k (what the reference's Welford recurrence over dnrm2(_arx[k]) accumulates); synchronises */
int bbgpu_cma_swarm_radius(bbgpu_cma *h, double *h_mean, double *h_m2);
/* synchronise the stream and return the report of everything enqueued so far */
int bbgpu_cma_fetch(bbgpu_cma *h, bbgpu_cma_report *out);
/* one whole generation without host round trips in between (sample -> evaluate -> rank ->
 * update); used by bench.py and by the host classes' fast path */
int bbgpu_cma_generation(bbgpu_cma *h, bbgpu_objective *obj, double sigma, uint64_t generation,
                         int ik, int fev_after);

/* ---- one large-lambda run sharded by candidates over several GPUs.
 * Create every rank's handle with cfg.shard_count = world, cfg.shard_rank = rank: state is
 * replicated, the handle samples and evaluates candidates [rank*lambda/world, (rank+1)*lambda/world)
 * from Philox counters keyed by the GLOBAL candidate id, and BBGPU_CMA_ARX / _Z hold only those
 * rows.  With a communicator attached, bbgpu_cma_evaluate ends with the fitness all-gather,
 * bbgpu_cma_update = local partial -> all-reduce -> replicated finish, and bbgpu_cma_generation
 * runs the whole generation without a host round trip.  Without one (single-GPU tests, custom
 * transports) the caller exchanges BBGPU_CMA_FITNESS (+ _NORMS) after evaluate and sums
 * BBGPU_CMA_SHARD_RED between bbgpu_cma_shard_partial and bbgpu_cma_shard_finish. */
#define BBGPU_COMM_ID_BYTES 128
typedef struct bbgpu_comm bbgpu_comm;
/* rank 0 creates the id and hands the 128 bytes to the other ranks (any side channel) */
int bbgpu_comm_unique_id(void *out128);
int bbgpu_comm_create(bbgpu_ctx *ctx, int world, int rank, const void *id128, bbgpu_comm **out);
int bbgpu_comm_destroy(bbgpu_comm *comm);
int bbgpu_comm_rank(bbgpu_comm *comm);
int bbgpu_comm_world(bbgpu_comm *comm);
int bbgpu_comm_allreduce_sum(bbgpu_comm *comm, double *d_buf, int64_t count);
/* best-of reduce over independent runs (ipop_cmaes.cpp:150-158 keeps the best run): d_f = device
 * scalar with this rank's best fitness, d_x = its n-vector (overwritten with the winner's) */
int bbgpu_comm_best_of(bbgpu_comm *comm, const double *d_f, double *d_x, int n, double *h_fbest,
                       int *h_owner);
/* the same with host buffers: f by value, h_x[n] in/out */
int bbgpu_comm_best_of_host(bbgpu_comm *comm, double f, double *h_x, int n, double *h_fbest,
                            int *h_owner);
int bbgpu_cma_attach_comm(bbgpu_cma *h, bbgpu_comm *comm);
/* 1 when the rank-mu exchange of this sharded handle runs as ONE kernel over NVLink peer memory (the
 * partial buffers of all ranks mapped into every rank: arrive barrier -> each rank sums its slice from
 * every peer in rank order and stores it into every peer -> done barrier), 0 when it is ncclAllReduce
 * (peer mapping unavailable, or BBGPU_SHARD_P2P=0).  A handle with a peer mapping must be destroyed
 * collectively (every rank calls bbgpu_cma_destroy). */
int bbgpu_cma_shard_uses_peer_memory(bbgpu_cma *h);
int bbgpu_cma_shard_partial(bbgpu_cma *h);
int bbgpu_cma_shard_finish(bbgpu_cma *h, double sigma, int fev);
/* candidate by GLOBAL index (bestSolution(), cma/base_cmaes.cpp:232-238); collective on a sharded
 * handle with a communicator */
int bbgpu_cma_get_candidate(bbgpu_cma *h, int index, double *h_dst);

/* ---- standalone fp64 tensor-core kernels, host-buffer form (unit tests, peak probes) */
/* C[M,N] = A[M,K] * B[K,N], all row-major */
int bbgpu_dgemm_host(bbgpu_ctx *ctx, int M, int N, int K, const double *h_A, const double *h_B,
                     double *h_C);
/* C[n,n] (lower triangle) = alpha*C + sum_k coef[k] * Y[k,:]^T Y[k,:], Y is K x n row-major */
int bbgpu_dsyrk_host(bbgpu_ctx *ctx, int n, int K, double alpha, const double *h_Y,
                     const double *h_coef, double *h_C);
/* sort fitness ascending, ties by index; writes the permutation */
int bbgpu_rank_host(bbgpu_ctx *ctx, int count, const double *h_f, int32_t *h_idx);
/* fill count standard normals from the Philox stream (key seed, counter (generation, i)) */
int bbgpu_philox_normal_host(bbgpu_ctx *ctx, uint64_t seed, uint64_t generation, int rows,
                             int cols, double *h_z);
/* time `reps` launches of the sampling GEMM shape (M=lambda,N=n,K=n) with CUDA events; ms */
int bbgpu_bench_dgemm(bbgpu_ctx *ctx, int M, int N, int K, int reps, float *ms_per_launch);

/* ---- DE engine: JADE (de/jade.cpp:98-206) and L-SHADE (de/shade.cpp:96-236) generations */
enum bbgpu_de_kind { BBGPU_DE_JADE = 0, BBGPU_DE_SHADE = 1 };

typedef struct bbgpu_de_config {
    int kind;
    int n;
    int np;        /* (initial) population size */
    int archive;   /* use the external archive */
    int repaircr;  /* record the realised crossover rate (de/jade.cpp:259-263) */
    int h;         /* L-SHADE memory size */
    int npmin;     /* L-SHADE minimum population */
    int mfev;      /* budget (drives L-SHADE's population reduction, de/shade.cpp:218-225) */
    double pelite; /* JADE */
    double cdamp;  /* JADE */
    double sigma;  /* JADE: CR power-mean switch (de/jade.cpp:188) */
    uint64_t seed;
} bbgpu_de_config;

/* tagged draws of one generation, host arrays of length np unless noted (DRAW REPLAY).
 * The reference consumes a data-dependent number of engine draws per candidate
 * (de/jade.cpp:105-183); the oracle's recorder tags the accepted ones. */
typedef struct bbgpu_de_draws {
    const double *zcr;    /* N(0,1) variate behind CR (de/jade.cpp:108) */
    const double *cauchy; /* accepted tan(pi(u-1/2)) variate behind F (de/jade.cpp:112-116) */
    const int32_t *ibest; /* de/jade.cpp:119-120 */
    const int32_t *r1;    /* de/jade.cpp:123-126 */
    const int32_t *r2;    /* de/jade.cpp:127-131 (>= np: archive row r2-np) */
    const int32_t *jrand; /* de/jade.cpp:249 */
    const int32_t *slot;  /* archive slot replaced on success with a full archive (:165-166) */
    const double *u;      /* np*n crossover uniforms (entry jrand unused) (de/jade.cpp:252) */
    const int32_t *ri;    /* L-SHADE memory slot (de/shade.cpp:105) */
    const double *pi;     /* L-SHADE greediness draw (de/shade.cpp:116) -- informational */
    const int32_t *trim;  /* L-SHADE archive trimming draws, ntrim entries (de/shade.cpp:228-235) */
    int ntrim;
} bbgpu_de_draws;

typedef struct bbgpu_de_report {
    int32_t np;       /* population size after the generation */
    int32_t larch;    /* archive size after the generation */
    int32_t nsucc;    /* trials accepted (f_trial <= f_parent) */
    int32_t fev;      /* evaluations so far */
    int32_t rounds;   /* dependency-resolution rounds the generation needed */
    int32_t kslot;    /* L-SHADE memory cursor */
    double mucr, muf; /* JADE adaptive means */
    double fbest;
    double radius_m2; /* Welford M2 of ||x_i||_2 over the swarm (de/jade.cpp:228-245) */
} bbgpu_de_report;

int bbgpu_de_create(bbgpu_ctx *ctx, const bbgpu_de_config *cfg, const double *h_lower,
                    const double *h_upper, bbgpu_de **out);
int bbgpu_de_destroy(bbgpu_de *h);
/* init(): de/jade.cpp:64-96, de/shade.cpp:56-94.  h_x == NULL: uniform Philox points;
 * h_x != NULL: np*n replayed initial points.  Evaluates the swarm. */
int bbgpu_de_init(bbgpu_de *h, bbgpu_objective *obj, const double *h_x);
/* one generation with the reference's sequential in-place semantics, resolved in parallel.
 * draws == NULL -> Philox. */
int bbgpu_de_generation(bbgpu_de *h, bbgpu_objective *obj, const bbgpu_de_draws *draws,
                        bbgpu_de_report *out);
/* One generation of `count` independent runs in ONE launch: same configuration (kind, n, np, budget),
 * their own seeds and state; Philox draws only.  Every run's result is bit-identical to
 * bbgpu_de_generation on its handle.  out: NULL or `count` reports. */
int bbgpu_de_generation_batch(bbgpu_de *const *hs, int count, bbgpu_objective *obj,
                              bbgpu_de_report *out);
/* swarm in the reference's order after the generation (sorted for L-SHADE; JADE sorts at the
 * start of the next generation, de/jade.cpp:101) */
int bbgpu_de_get_swarm(bbgpu_de *h, double *h_x, double *h_f);
int bbgpu_de_get_archive(bbgpu_de *h, double *h_arch, int *rows);
int bbgpu_de_get_memory(bbgpu_de *h, double *h_mcr, double *h_mf);
int bbgpu_de_get_trace(bbgpu_de *h, int32_t *h_succ, double *h_ftrial, double *h_F, double *h_CR);
/* diagnostics: device timestamps (ns) at the phase boundaries of the last generation kernel:
 * [0] start [1] sorted + parameters [2..13] rounds [16] finalized [17] L-SHADE sort [18] end;
 * [32 + 4r + k], r < 8: round r after scan / signatures / builds / grid.sync.  64 entries. */
int bbgpu_de_phase_times(bbgpu_de *h, uint64_t *out64);

/* ---- SaNSDE engine (de/sansde.cpp:94-212): self-adaptive DE with neighbourhood search.  Same in-place sequential semantics
 * as JADE, resolved by parallel fixed-point rounds; the draws of a candidate do not depend on the history of the
 * generation, so they are generated (Philox) or replayed up front. */
typedef struct bbgpu_sde bbgpu_sde;
/* tagged draws of one generation, host arrays of length np unless noted (DRAW REPLAY; the oracle's recorder tags them) */
typedef struct bbgpu_sde_draws {
    const double *zcr;      /* N(0,1) behind a refreshed CR (de/sansde.cpp:108-111); may be NULL when CR is not refreshed */
    const int32_t *ifstrat; /* which F distribution: 0 normal, 1 Cauchy (:114-119) */
    const double *F;        /* the accepted F (:120-128) */
    const int32_t *r1, *r2, *r3; /* :131-142 */
    const int32_t *istrat;  /* mutation strategy: 0 rand/1, 1 current-to-best/2 (:145-150) */
    const int32_t *jrand;   /* :269 */
    const double *u;        /* np*n crossover uniforms (:271-272; drawn for every coordinate) */
} bbgpu_sde_draws;
/* the adaptation sums the candidate loop accumulates (:160-176), carried by the host between generations */
typedef struct bbgpu_sde_accum {
    int32_t pns[2], pnf[2];
    double crrec, crdeltaf, fpns[2], fpnf[2];
    int32_t nsucc, pad; /* out: trials accepted in this generation */
    double rmean, rm2;  /* out: converged() (:235-253) on the swarm this generation leaves: mean radius, sum of squared deviations */
} bbgpu_sde_accum;
int bbgpu_sde_create(bbgpu_ctx *ctx, int n, int np, int repaircr, uint64_t seed, const double *h_lower, const double *h_upper,
                     bbgpu_sde **out);
int bbgpu_sde_destroy(bbgpu_sde *h);
/* init(): de/sansde.cpp:54-92.  h_x == NULL: uniform Philox points; else np*n replayed points.  Evaluates the swarm. */
int bbgpu_sde_init(bbgpu_sde *h, bbgpu_objective *obj, const double *h_x);
/* iterate() up to the periodic adaptation (:94-177): sort, CR refresh (refresh_cr = it % crref == 0, around crm), F / parents /
 * strategy draws (fp, p), mutation + crossover + bound control, greedy selection in place, the adaptation sums (acc in/out).
 * draws == NULL -> Philox.  *rounds_out (may be NULL): dependency-resolution rounds the generation needed. */
int bbgpu_sde_generation(bbgpu_sde *h, bbgpu_objective *obj, int refresh_cr, double crm, double fp, double p,
                         const bbgpu_sde_draws *draws, bbgpu_sde_accum *acc, int *rounds_out);
int bbgpu_sde_sort(bbgpu_sde *h); /* std::sort(_swarm) as solution() does (:214-217) */
int bbgpu_sde_get_swarm(bbgpu_sde *h, double *h_x, double *h_f, double *h_cr);
int bbgpu_sde_get_trace(bbgpu_sde *h, int32_t *h_succ, double *h_ftrial, double *h_cr1);
int bbgpu_sde_radius(bbgpu_sde *h, double *h_mean, double *h_m2); /* converged(), :235-253 */
int bbgpu_sde_evaluations(bbgpu_sde *h);

/* ---- SSDE engine (de/ssde.cpp:134-339): spherical search with an optional DE stage and linear population reduction.  The
 * generation depends on its own history through the live rows and -- with usede -- through R = fev / mfev, which advances with
 * the second-stage evaluations of the earlier candidates; both are resolved by parallel fixed-point rounds.  The random
 * rotation A (randA, :410-432) is never formed: it has at most two non-zeros per row. */
typedef struct bbgpu_ssde bbgpu_ssde;
/* tagged draws of one generation, host arrays of length np unless noted (DRAW REPLAY; the oracle's recorder tags them) */
typedef struct bbgpu_ssde_draws {
    const int32_t *iL;      /* memory slot (:151) */
    const double *zrank;    /* N(0,1) behind the mask probability (:154-155) */
    const double *ub;       /* np*n mask uniforms (:156-162) */
    const int32_t *p, *q, *r, *pbest; /* :165-170 */
    const double *cauchy;   /* the accepted Cauchy value behind c_i (:225-229) */
    /* second stage, usede only; entries of candidates that did not reach it are ignored */
    const double *zcr;      /* N(0,1) behind CR_i (:253-254) */
    const int32_t *j0;      /* :258 */
    const double *ucr;      /* np*n crossover uniforms, NaN where the reference drew none (:260) */
    const double *redraw;   /* np*n uniform replacements of coordinates that left the box, NaN elsewhere (:264-266) */
} bbgpu_ssde_draws;
typedef struct bbgpu_ssde_state {
    int32_t k, kcr;   /* cursors of the success memories (:309-314, :335-340), 1-based as in the reference */
    int32_t evals;    /* evaluations this generation spent (np + second-stage trials) */
    int32_t nsucc;    /* trials accepted */
    int32_t np_next;  /* population after the linear reduction (:354-360); the handle has already dropped the others */
    int32_t pad;
    double fbest;     /* best fitness after the generation (the patience counter, :347-351) */
    double rmean, rm2;/* converged() (:383-408) over the members that stay: mean radius and sum of squared deviations */
} bbgpu_ssde_state;
int bbgpu_ssde_create(bbgpu_ctx *ctx, int n, int npinit, int h_mem, int usede, int repaircr, uint64_t seed, const double *h_lower,
                      const double *h_upper, bbgpu_ssde **out);
int bbgpu_ssde_destroy(bbgpu_ssde *h);
/* init(): de/ssde.cpp:59-132.  h_x == NULL: uniform Philox points; else npinit*n replayed points.  With usede the opposite
 * points are added, everything is evaluated and sorted and the best npinit stay.  *fev_out: evaluations spent. */
int bbgpu_ssde_init(bbgpu_ssde *h, bbgpu_objective *obj, const double *h_x, int *fev_out);
/* iterate() (:134-361) including the population reduction (npmin: its floor).  h_iwork: the shuffled index list behind A (n entries; consecutive pairs
 * rotate).  fev / mfev: evaluations before the generation and the budget.  draws == NULL -> Philox. */
int bbgpu_ssde_generation(bbgpu_ssde *h, bbgpu_objective *obj, const int32_t *h_iwork, int fev, int mfev, double ptop, int npmin,
                          const bbgpu_ssde_draws *draws, bbgpu_ssde_state *state, int *rounds_out);
int bbgpu_ssde_set_population(bbgpu_ssde *h, int np); /* keep the best np (the generation applies the schedule itself) */
int bbgpu_ssde_get_swarm(bbgpu_ssde *h, double *h_x, double *h_f);
int bbgpu_ssde_get_memory(bbgpu_ssde *h, double *h_L1, double *h_L2, double *h_LCR); /* H doubles each; any may be NULL */
/* outcome of the last generation by candidate (before the sort): 0 rejected, 1 first stage, 2 second stage; the fitness of the
 * last trial the candidate evaluated */
int bbgpu_ssde_get_trace(bbgpu_ssde *h, int32_t *h_succ, double *h_ftrial);
int bbgpu_ssde_radius(bbgpu_ssde *h, double *h_mean, double *h_m2); /* converged(), :383-408 */

/* ---- batched small-n CMA-ES engine (BASELINE config 5: thousands of independent
 * IPOP/BIPOP runs at n <= 64).  One CTA per run; sample -> evaluate -> rank -> update -> updateSigma ->
 * lazy eigendecomposition (the reference's tred2/tql2, inside the CTA) -> updateHistory -> converged()
 * all run on the device, whole generations per launch (cma/base_cmaes.cpp:150-174 per run).  Registry
 * objectives are evaluated inside the kernel; a user CUDA objective or a batched host callback is called
 * ONCE per generation with the candidates of all listed runs (sample launch -> evaluate -> update
 * launch, one status read per generation).  The restart schedule (cma/ipop_cmaes.cpp,
 * cma/bipop_cmaes.cpp) stays with the caller: it starts base runs, steps the batch, reads the status
 * words and restarts the runs that stopped. */
typedef struct bbgpu_cmab_run_params { /* one base run == BaseCmaes::init + Cmaes::init */
    int32_t lambda, mu, mfev, mit, hlen, ik; /* cma/base_cmaes.cpp:71,88-92,128-129 */
    double sigma0, tol;
    double cc, cs, c1, cmu, mueff, damps, chi; /* cma/base_cmaes.cpp:94-117 */
    double eigenfreq;                          /* cma/cmaes.cpp:48 */
    uint64_t seed;                             /* Philox key of this base run */
    const double *weights;                     /* mu recombination weights (host; a pure function of
                                                  lambda, cached on the device per lambda) */
} bbgpu_cmab_run_params;

typedef struct bbgpu_cmab_run_status {
    int32_t status;  /* 0 running, 1 converged() returned true, 2 fev >= mfev, 3 decomposition failed */
    int32_t flag;    /* Cmaes::_flag 1..9 (cma/cmaes.cpp:151-227) */
    int32_t it, fev, lambda;
    int32_t eig_count; /* decompositions so far */
    int32_t eig_due;   /* replay mode: the gate fired in the last step, the caller adopts a basis */
    int32_t hsig;
    int32_t ql_iters;  /* QL iterations spent by the in-CTA solver so far */
    int32_t generations_this_step;
    int32_t ibw[4];
    double ybw[4];
    double sigma, fx_final /* f(bestSolution()) */, pslen, fkth, hist_best, hist_worst;
} bbgpu_cmab_run_status;

int bbgpu_cmab_create(bbgpu_ctx *ctx, int n, int runs, int bound, const double *h_lower,
                      const double *h_upper, bbgpu_cmab **out);
int bbgpu_cmab_destroy(bbgpu_cmab *h);
/* engine limits at dimension n, so that a caller can validate a restart schedule BEFORE it spends compute: the largest
 * dimension, the largest lambda whose ranking fits the CTA's shared memory, the capacity of the history rings
 * (hlen = 10 + ceil(30 n / lambda), cma/base_cmaes.cpp:128).  Any out pointer may be NULL. */
int bbgpu_cmab_limits(int n, int *max_n, int *max_lambda, int *max_hlen);
/* (re)start `count` runs: m = guess (count*n, host), paths = 0, D = 1, C = B = C^-1/2 = I.
 * keep_offdiag != 0: only the diagonals of B and C are reset (the reference re-inits the same object
 * on a restart, cma/cmaes.cpp:53-60) */
int bbgpu_cmab_start(bbgpu_cmab *h, int count, const int32_t *run_ids,
                     const bbgpu_cmab_run_params *params, const double *h_guess, int keep_offdiag);
/* advance the listed runs (async on the ctx stream): each run does whole generations until it stops,
 * has spent `quantum` evaluations in this call, or has done max_gens generations.
 * h_z != NULL (needs max_gens == 1): DRAW REPLAY, sum(lambda_r)*n standard normals, runs in list
 * order, candidate-major.  eig_mode 0: in-CTA decomposition; 1: EIGENBASIS REPLAY -- the gate only
 * raises status.eig_due and the caller calls bbgpu_cmab_adopt_basis. */
int bbgpu_cmab_step(bbgpu_cmab *h, bbgpu_objective *obj, int count, const int32_t *run_ids,
                    int quantum, int max_gens, const double *h_z, int eig_mode);
/* synchronise; status words of ALL runs (h_out[runs]) */
int bbgpu_cmab_fetch(bbgpu_cmab *h, bbgpu_cmab_run_status *h_out);
/* bestSolution() of runs [run0, run0+count) (cma/base_cmaes.cpp:232-238), count*n doubles */
int bbgpu_cmab_get_best(bbgpu_cmab *h, int run0, int count, double *h_x);
/* per-run state, fields BBGPU_CMA_XMEAN, _XOLD, _PC, _PS, _C, _B, _D, _INVSQRTC */
int bbgpu_cmab_get(bbgpu_cmab *h, int run, int field, double *h_dst, int64_t count);
int bbgpu_cmab_set(bbgpu_cmab *h, int run, int field, const double *h_src, int64_t count);
int bbgpu_cmab_set_sigma(bbgpu_cmab *h, int run, double sigma);
/* candidates, fitness (by candidate) and ranking of the run's last generation of the last step */
int bbgpu_cmab_get_population(bbgpu_cmab *h, int run, double *h_X, double *h_f, int32_t *h_rank);
int bbgpu_cmab_adopt_basis(bbgpu_cmab *h, int run, const double *h_B, const double *h_D);
/* Householder tridiagonalisation of ONE dense symmetric matrix as one persistent cooperative kernel (the
 * matrix distributed over the SMs' shared memory, one grid barrier per column): the building block of
 * a replacement for cuSOLVER's sytrd, 80 % of the n = 1000 decomposition (cma/cmaes.cpp:285-381 is the
 * reference's tred2).  Host buffers; h_A n*n row-major (lower triangle read); outputs in LAPACK
 * dsytrd('L') convention: d[n], e[n-1], tau[n-1], h_V (may be NULL) n*n COLUMN-major with reflector i in
 * rows i+2.. of column i.  *ms (may be NULL): best kernel time of `reps` launches. */
int bbgpu_sytrd_host(bbgpu_ctx *ctx, int n, const double *h_A, double *h_d, double *h_e, double *h_tau,
                     double *h_V, int reps, float *ms);
/* Divide & conquer eigensolver for ONE symmetric tridiagonal matrix, entirely on the device (leaves by the
 * one-warp QL solver, merges = deflation scan + secular equation + Loewner vectors + one GEMM): the second
 * half of the replacement for cuSOLVER's syevd.  Host buffers: d[n], e[n-1] -> w[n] ascending, Zt[n*n]
 * with ROW j = eigenvector of w[j].  *ms (may be NULL): device time of one decomposition. */
int bbgpu_stedc_host(bbgpu_ctx *ctx, int n, const double *h_d, const double *h_e, double *h_w,
                     double *h_Zt, float *ms);
/* the in-CTA solver alone (unit tests): `count` symmetric n x n matrices (lower triangles read) ->
 * B (B[i*n+j] = component i of eigenvector j), eigenvalues ascending, info = QL iterations or -1 */
int bbgpu_eig_small_host(bbgpu_ctx *ctx, int n, int count, const double *h_A, double *h_B,
                         double *h_D, int32_t *h_info);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* BBGPU_H_ */
