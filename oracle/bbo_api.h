/*
 * ORACLE (test infrastructure).  One C API, exported twice with identical symbols:
 *
 *   oracle/_ref/libbbo_ref.so   the UNMODIFIED reference sources (/root/reference/src/...)
 *                               compiled in place behind oracle/ref_harness.cpp
 *   oracle/libbbo_port.so       oracle/port/bbo_port.cpp, our CPU restatement of the same
 *                               algorithms (pinned bit-for-bit against the library above by
 *                               tests/test_oracle_pinning.py and tests/golden/)
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg
 * may load either library.
 */
#ifndef BBO_API_H_
#define BBO_API_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* which implementation is behind this library: "reference" or "port" */
const char *bbo_kind(void);

/* ---- process-global random engine (reference: src/random.hpp:677-680, one static mt19937) */
void bbo_seed(uint32_t s);
/* kind 0: count draws of N(0,1) from ONE std::normal_distribution object (keeps its cache)
 * kind 1: uniform_real(a,b)   kind 2: uniform_int(a,b) (closed)   kind 3: raw engine u32 */
void bbo_rng_draw(int kind, int count, double a, double b, double *out);

/* ---- objectives (spec: oracle/objectives_ref.h) */
double bbo_eval(int obj, int n, const double *x);

/* ---- CMA-ES family.  kind: 0 Cmaes, 1 ActiveCmaes, 2 SepCmaes, 3 CholeskyCmaes (its `stol` travels in
 * the eigenrate argument; cma/cholesky_cmaes.h:44-48), 4 LmCmaes (cma/lm_cmaes.h:62-70: `memory` travels in the eigenrate
 * argument, `rademacher` in alphacov (0 / 1), `usenew` in adjustlr) */
void *bbo_cma_new(int kind, int mfev, double tol, int np, double sigma0, int bound,
                  double eigenrate, double alphacov, int adjustlr);
void bbo_cma_free(void *h);
void bbo_cma_init(void *h, int obj, int n, const double *lower, const double *upper,
                  const double *guess);
void bbo_cma_set_params(void *h, int np, double sigma, int mfev);
void bbo_cma_sample(void *h);
void bbo_cma_evalsort(void *h);
void bbo_cma_update(void *h);
void bbo_cma_history(void *h); /* updateHistory(); _it++  (base_cmaes.cpp:150-156) */
void bbo_cma_iterate(void *h);
int bbo_cma_converged(void *h);
/* full optimize(); writes best x (n) and returns converged; *fev receives evaluations */
int bbo_cma_optimize(void *h, int obj, int n, const double *lower, const double *upper,
                     const double *guess, double *xout, int *fev);
int bbo_cma_geti(void *h, const char *name);
double bbo_cma_getd(void *h, const char *name);
/* copies up to cap doubles, returns the field length (call with cap=0 to size) */
int bbo_cma_getv(void *h, const char *name, double *out, int cap);
/* CPU-baseline helper: force one eigenDecomposition() of the current C (cma/cmaes.cpp:229-283)
 * past the lazy gate and return its wall time in seconds (0 for the separable variant) */
double bbo_cma_time_eig(void *h);
/* reset the optimizer's own normal_distribution cache (public _Z member) */
void bbo_cma_reset_z(void *h);
/* eigenbasis-replay feed for whole optimize() / restart-driver runs: with the log enabled, every
 * eigenDecomposition() that fires (cma/cmaes.cpp:233-283) appends (fev, B, D) -- B n*n with
 * B[i*n+j] = component i of eigenvector j, D n.  enable != 0 clears the log and starts recording. */
void bbo_cma_eig_log(void *h, int enable);
int bbo_cma_eig_log_count(void *h);
/* event k: returns its n (0 when k is out of range); B / D may be NULL */
int bbo_cma_eig_log_get(void *h, int k, int *fev, double *B, double *D);

/* ---- restart drivers.  kind: 0 IPOP, 1 BIPOP; base = handle from bbo_cma_new.
 * trace receives up to cap rows of {lambda, sigma, fx, fev_total, regime}; returns rows */
int bbo_restart_optimize(int kind, void *base, int mfev, double sigma0, int flag /*nipop|nbipop*/,
                         double ksigmadec, int boundlambda, int maxlargeruns, double kbudget,
                         int obj, int n, const double *lower, const double *upper,
                         const double *guess, double *xout, int *fev, double *trace, int cap);

/* ---- DE family.  kind: 0 JADE, 1 L-SHADE */
void *bbo_de_new(int kind, int mfev, int np, double tol, int archive, int repaircr,
                 double pelite, double cdamp, double sigma, int h, int npmin);
void bbo_de_free(void *h);
void bbo_de_init(void *h, int obj, int n, const double *lower, const double *upper,
                 const double *guess);
void bbo_de_iterate(void *h);
int bbo_de_converged(void *h);
int bbo_de_optimize(void *h, int obj, int n, const double *lower, const double *upper,
                    const double *guess, double *xout, int *fev);
int bbo_de_geti(void *h, const char *name);
double bbo_de_getd(void *h, const char *name);
int bbo_de_getv(void *h, const char *name, double *out, int cap);

/* ---- xNES (nes/xnes.cpp): exponential natural evolution strategy; np = 4 + floor(3 ln n) points, the mean starts at
 * the origin (the guess is ignored, :82).  getv: "u", "mu", "B" (n*n), "Gdelta", and -- in sorted order after iterate() --
 * "x" (np*n), "z" (np*n), "f" (np) */
void *bbo_xnes_new(int mfev, double tol, double a0, double etamu);
void bbo_xnes_free(void *h);
void bbo_xnes_init(void *h, int obj, int n, const double *lower, const double *upper, const double *guess);
void bbo_xnes_iterate(void *h);
int bbo_xnes_converged(void *h);
int bbo_xnes_optimize(void *h, int obj, int n, const double *lower, const double *upper, const double *guess,
                      double *xout, int *fev);
int bbo_xnes_geti(void *h, const char *name);
double bbo_xnes_getd(void *h, const char *name);
int bbo_xnes_getv(void *h, const char *name, double *out, int cap);

/* ---- SaNSDE (de/sansde.cpp): self-adaptive DE with neighbourhood search, in-place (asynchronous) like JADE.
 * geti: "n", "np", "fev", "it", "pns0", "pns1", "pnf0", "pnf1";  getd: "p", "fp", "crm", "crrec", "crdeltaf", "fpns0", ...;
 * getv: "swarm_x" (np*n), "swarm_f", "swarm_cr" (np) in the order the reference holds them, and -- port only -- the tagged
 * draws of the generation just run: "t_zcr" (NaN when CR was not refreshed), "t_ifstrat", "t_F", "t_r1", "t_r2", "t_r3",
 * "t_istrat", "t_jrand", "t_u" (np*n), plus its outcome "t_succ", "t_ftrial", "t_cr1" */
void *bbo_sansde_new(int mfev, int np, double tol, int repaircr, int crref, int pupdate, int crupdate);
void bbo_sansde_free(void *h);
void bbo_sansde_init(void *h, int obj, int n, const double *lower, const double *upper, const double *guess);
void bbo_sansde_iterate(void *h);
int bbo_sansde_converged(void *h);
int bbo_sansde_optimize(void *h, int obj, int n, const double *lower, const double *upper, const double *guess,
                        double *xout, int *fev);
int bbo_sansde_geti(void *h, const char *name);
double bbo_sansde_getd(void *h, const char *name);
int bbo_sansde_getv(void *h, const char *name, double *out, int cap);

/* ---- SSDE (de/ssde.cpp): spherical search (+ optionally a DE stage for the failures), success-history memory, linear
 * population reduction.  geti: "n", "np", "fev", "convcount", "k", "kcr";  getv: "swarm_x", "swarm_f", "L1", "L2", "LCR", "iwork"
 * and -- port only -- the tagged draws of the generation just run: "t_iwork" (n, the permutation behind A), per candidate
 * "t_iL", "t_zrank", "t_p", "t_q", "t_r", "t_pbest", "t_cauchy" (the accepted variate), "t_ub" (np*n), outcome "t_succ"
 * (0 fail, 1 first stage, 2 DE stage), "t_f" (fitness of the accepted or last trial); with usede also "t_zcr", "t_j0",
 * "t_ucr" (np*n), "t_redraw" (np*n, NaN where no coordinate was redrawn) */
void *bbo_ssde_new(int mfev, int npinit, double tol, int patience, int npmin, double ptop, int h, int usede, int repaircr);
void bbo_ssde_free(void *h);
void bbo_ssde_init(void *h, int obj, int n, const double *lower, const double *upper, const double *guess);
void bbo_ssde_iterate(void *h);
int bbo_ssde_converged(void *h);
int bbo_ssde_optimize(void *h, int obj, int n, const double *lower, const double *upper, const double *guess,
                      double *xout, int *fev);
int bbo_ssde_geti(void *h, const char *name);
double bbo_ssde_getd(void *h, const char *name);
int bbo_ssde_getv(void *h, const char *name, double *out, int cap);

#ifdef __cplusplus
}
#endif
#endif
