// TEST INFRASTRUCTURE: a CPU stand-in for libbbgpu.so's C ABI, just enough to drive the HOST logic
// (restart schedules, budgets, history rings, printing) under sanitizers without a GPU.  It does no
// optimisation: fitness improves geometrically with the generation count.  Never shipped.
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "bbgpu.h"

struct bbgpu_ctx { int launches = 0; };
struct bbgpu_objective { int n; bbgpu_host_objective_fn host; void *user; };
struct bbgpu_cma { bbgpu_cma_config cfg; std::vector<double> mean; int gen = 0; double f = 100.0; };
struct bbgpu_de { bbgpu_de_config cfg; int fev = 0; int gen = 0; };

extern "C" {
int bbgpu_version(void) { return BBGPU_VERSION; }
const char *bbgpu_last_error(void) { return "mock"; }
int bbgpu_ctx_create(int, void *, bbgpu_ctx **out) { *out = new bbgpu_ctx(); return 0; }
int bbgpu_ctx_destroy(bbgpu_ctx *c) { delete c; return 0; }
int bbgpu_ctx_sync(bbgpu_ctx *) { return 0; }
void *bbgpu_ctx_stream(bbgpu_ctx *) { return nullptr; }
int bbgpu_ctx_device(bbgpu_ctx *) { return 0; }
int64_t bbgpu_ctx_launch_count(bbgpu_ctx *c) { return c->launches; }
int bbgpu_ctx_profile_enable(bbgpu_ctx *, int) { return 0; }
int bbgpu_ctx_profile_read(bbgpu_ctx *, int, double *t, int64_t *c) { *t = 0; *c = 0; return 0; }
int bbgpu_ctx_profile_reset(bbgpu_ctx *) { return 0; }
int bbgpu_ctx_timer_start(bbgpu_ctx *) { return 0; }
int bbgpu_ctx_timer_stop(bbgpu_ctx *, float *ms) { *ms = 1.f; return 0; }
int bbgpu_ctx_flush_l2(bbgpu_ctx *) { return 0; }
int bbgpu_objective_builtin(bbgpu_ctx *, int, int n, bbgpu_objective **out) { *out = new bbgpu_objective{n, nullptr, nullptr}; return 0; }
int bbgpu_objective_device(bbgpu_ctx *, bbgpu_device_objective_fn, void *, int n, bbgpu_objective **out) { *out = new bbgpu_objective{n, nullptr, nullptr}; return 0; }
int bbgpu_objective_host(bbgpu_ctx *, bbgpu_host_objective_fn fn, void *user, int n, bbgpu_objective **out) { *out = new bbgpu_objective{n, fn, user}; return 0; }
int bbgpu_objective_destroy(bbgpu_objective *o) { delete o; return 0; }
int bbgpu_eval(bbgpu_ctx *, bbgpu_objective *, const double *, int64_t, int, double *) { return 0; }
int bbgpu_eval_host(bbgpu_ctx *, bbgpu_objective *o, const double *x, int count, double *f) {
    if (o->host) return o->host(x, o->n, o->n, count, f, o->user);
    for (int k = 0; k < count; k++) { double s = 0; for (int i = 0; i < o->n; i++) s += x[k * o->n + i] * x[k * o->n + i]; f[k] = s; }
    return 0;
}
int bbgpu_cma_create(bbgpu_ctx *, const bbgpu_cma_config *cfg, const double *, const double *, const double *, bbgpu_cma **out) {
    bbgpu_cma *h = new bbgpu_cma(); h->cfg = *cfg; h->mean.assign(cfg->n, 0.0); *out = h; return 0;
}
int bbgpu_cma_destroy(bbgpu_cma *h) { delete h; return 0; }
int bbgpu_cma_reset(bbgpu_cma *h, const double *g, int) { h->mean.assign(g, g + h->cfg.n); h->gen = 0; h->f = 100.0; return 0; }
int bbgpu_cma_set(bbgpu_cma *, int, const double *, int64_t) { return 0; }
int bbgpu_cma_get(bbgpu_cma *h, int field, double *dst, int64_t count) {
    for (int64_t i = 0; i < count; i++) dst[i] = (field == BBGPU_CMA_XMEAN && i < (int64_t) h->mean.size()) ? h->mean[i] : 0.0;
    return 0;
}
int bbgpu_cma_get_rank(bbgpu_cma *, int32_t *idx, int count) { for (int i = 0; i < count; i++) idx[i] = i; return 0; }
int bbgpu_cma_get_rows(bbgpu_cma *h, int, int64_t, int64_t rows, double *dst) {
    for (int64_t r = 0; r < rows; r++) for (int i = 0; i < h->cfg.n; i++) dst[r * h->cfg.n + i] = h->mean[i] * std::pow(0.9, h->gen);
    return 0;
}
int bbgpu_cma_device_ptr(bbgpu_cma *, int, void **p, int64_t *ld) { *p = nullptr; *ld = 0; return 0; }
int bbgpu_cma_sample(bbgpu_cma *, double, uint64_t, const double *) { return 0; }
int bbgpu_cma_sample_dev(bbgpu_cma *, double, const double *) { return 0; }
int bbgpu_cma_evaluate(bbgpu_cma *, bbgpu_objective *) { return 0; }
int bbgpu_cma_rank(bbgpu_cma *, int) { return 0; }
int bbgpu_cma_update(bbgpu_cma *h, double, int) { h->gen++; h->f *= 0.8; return 0; }
int bbgpu_cma_eig(bbgpu_cma *, const double *, const double *) { return 0; }
int bbgpu_cma_probe(bbgpu_cma *, double, int) { return 0; }
int bbgpu_cma_swarm_radius(bbgpu_cma *, double *mean, double *m2) { *mean = 1.0; *m2 = 1.0; return 0; }
int bbgpu_cma_xnes_update(bbgpu_cma *h, double, double, double, double *g) { h->gen++; *g = 0.0; return 0; }
int bbgpu_cma_lm_sample(bbgpu_cma *, double, uint64_t, int, const int32_t *, const double *, const int32_t *) { return 0; }
int bbgpu_cma_lm_update(bbgpu_cma *h, double, int, int, int, const int32_t *) { h->gen++; h->f *= 0.8; return 0; }
int bbgpu_cma_fetch(bbgpu_cma *h, bbgpu_cma_report *r) {
    std::memset(r, 0, sizeof(*r));
    for (int q = 0; q < 4; q++) { r->ibw[q] = q; r->ybw[q] = h->f * (1 + q); }
    r->f_kth = h->f * 1.5; r->pslen = h->cfg.chi; r->hsig = 1; r->d_min = 1; r->d_max = 1; r->tolx_max = std::pow(0.8, h->gen);
    return 0;
}
int bbgpu_cma_generation(bbgpu_cma *h, bbgpu_objective *, double, uint64_t, int, int) { h->gen++; return 0; }
int bbgpu_dgemm_host(bbgpu_ctx *, int, int, int, const double *, const double *, double *) { return 0; }
int bbgpu_dsyrk_host(bbgpu_ctx *, int, int, double, const double *, const double *, double *) { return 0; }
int bbgpu_rank_host(bbgpu_ctx *, int, const double *, int32_t *) { return 0; }
int bbgpu_philox_normal_host(bbgpu_ctx *, uint64_t, uint64_t, int, int, double *) { return 0; }
int bbgpu_bench_dgemm(bbgpu_ctx *, int, int, int, int, float *) { return 0; }
int bbgpu_de_create(bbgpu_ctx *, const bbgpu_de_config *cfg, const double *, const double *, bbgpu_de **out) { bbgpu_de *h = new bbgpu_de(); h->cfg = *cfg; *out = h; return 0; }
int bbgpu_de_destroy(bbgpu_de *h) { delete h; return 0; }
int bbgpu_de_init(bbgpu_de *h, bbgpu_objective *, const double *) { h->fev = h->cfg.np; h->gen = 0; return 0; }
int bbgpu_de_generation(bbgpu_de *h, bbgpu_objective *, const bbgpu_de_draws *, bbgpu_de_report *r) {
    h->fev += h->cfg.np; h->gen++;
    if (r) { std::memset(r, 0, sizeof(*r)); r->np = h->cfg.np; r->fev = h->fev; r->rounds = 3; r->radius_m2 = std::pow(0.5, h->gen); }
    return 0;
}
int bbgpu_de_get_swarm(bbgpu_de *h, double *x, double *f) {
    if (x) for (int i = 0; i < h->cfg.np * h->cfg.n; i++) x[i] = 0.01 * i;
    if (f) for (int i = 0; i < h->cfg.np; i++) f[i] = 10.0 - 0.001 * i;
    return 0;
}
int bbgpu_de_get_archive(bbgpu_de *, double *, int *rows) { *rows = 0; return 0; }
int bbgpu_de_get_memory(bbgpu_de *, double *, double *) { return 0; }
int bbgpu_de_get_trace(bbgpu_de *, int32_t *, double *, double *, double *) { return 0; }
int bbgpu_de_phase_times(bbgpu_de *, uint64_t *) { return 0; }
int bbgpu_comm_unique_id(void *) { return 0; }
int bbgpu_comm_create(bbgpu_ctx *, int, int, const void *, bbgpu_comm **) { return 0; }
int bbgpu_comm_destroy(bbgpu_comm *) { return 0; }
int bbgpu_comm_rank(bbgpu_comm *) { return 0; }
int bbgpu_comm_world(bbgpu_comm *) { return 1; }
int bbgpu_comm_allreduce_sum(bbgpu_comm *, double *, int64_t) { return 0; }
int bbgpu_comm_best_of(bbgpu_comm *, const double *, double *, int, double *, int *) { return 0; }
int bbgpu_cma_attach_comm(bbgpu_cma *, bbgpu_comm *) { return 0; }
int bbgpu_cma_shard_uses_peer_memory(bbgpu_cma *) { return 0; }
int bbgpu_cma_shard_partial(bbgpu_cma *) { return 0; }
int bbgpu_cma_shard_finish(bbgpu_cma *, double, int) { return 0; }
int bbgpu_cma_get_candidate(bbgpu_cma *, int, double *) { return 0; }
int bbgpu_de_generation_batch(bbgpu_de *const *, int, bbgpu_objective *, bbgpu_de_report *) { return 0; }
int bbgpu_comm_best_of_host(bbgpu_comm *, double f, double *, int, double *fb, int *owner) { *fb = f; *owner = 0; return 0; }
// batched engine: a base run "converges" after 40 generations or when its budget is spent
struct bbgpu_cmab { int n, runs; std::vector<bbgpu_cmab_run_params> p; std::vector<bbgpu_cmab_run_status> st; std::vector<double> x; };
int bbgpu_cmab_create(bbgpu_ctx *, int n, int runs, int, const double *, const double *, bbgpu_cmab **out) {
    bbgpu_cmab *h = new bbgpu_cmab(); h->n = n; h->runs = runs; h->p.resize(runs); h->st.resize(runs); h->x.assign((size_t) runs * n, 0.0); *out = h; return 0;
}
int bbgpu_cmab_destroy(bbgpu_cmab *h) { delete h; return 0; }
int bbgpu_cmab_limits(int, int *max_n, int *max_lambda, int *max_hlen) {
    if (max_n) *max_n = 64;
    if (max_lambda) *max_lambda = 16384;
    if (max_hlen) *max_hlen = 512;
    return 0;
}
int bbgpu_cmab_start(bbgpu_cmab *h, int count, const int32_t *ids, const bbgpu_cmab_run_params *p, const double *g, int) {
    for (int q = 0; q < count; q++) {
        h->p[ids[q]] = p[q];
        std::memset(&h->st[ids[q]], 0, sizeof(bbgpu_cmab_run_status));
        h->st[ids[q]].lambda = p[q].lambda; h->st[ids[q]].sigma = p[q].sigma0;
        for (int i = 0; i < h->n; i++) h->x[(size_t) ids[q] * h->n + i] = g[(size_t) q * h->n + i];
    }
    return 0;
}
int bbgpu_cmab_step(bbgpu_cmab *h, bbgpu_objective *, int count, const int32_t *ids, int quantum, int max_gens, const double *, int) {
    for (int q = 0; q < count; q++) {
        bbgpu_cmab_run_status &s = h->st[ids[q]];
        const bbgpu_cmab_run_params &p = h->p[ids[q]];
        int spent = 0, g = 0;
        s.generations_this_step = 0;
        while (s.status == 0 && g < max_gens && (g == 0 || spent < quantum)) {
            if (s.fev >= p.mfev) { s.status = 2; break; }
            s.fev += p.lambda; s.it++; g++; spent += p.lambda; s.generations_this_step++;
            s.fx_final = 100.0 * std::pow(0.9, s.it) / p.lambda;
            if (s.it >= 40 || s.it >= p.mit) { s.status = 1; s.flag = s.it >= p.mit ? 1 : 2; }
        }
    }
    return 0;
}
int bbgpu_cmab_fetch(bbgpu_cmab *h, bbgpu_cmab_run_status *out) { std::memcpy(out, h->st.data(), sizeof(bbgpu_cmab_run_status) * h->runs); return 0; }
int bbgpu_cmab_get_best(bbgpu_cmab *h, int r0, int count, double *x) { std::memcpy(x, &h->x[(size_t) r0 * h->n], sizeof(double) * count * h->n); return 0; }
int bbgpu_cmab_get(bbgpu_cmab *, int, int, double *, int64_t) { return 0; }
int bbgpu_cmab_set(bbgpu_cmab *, int, int, const double *, int64_t) { return 0; }
int bbgpu_cmab_set_sigma(bbgpu_cmab *, int, double) { return 0; }
int bbgpu_cmab_get_population(bbgpu_cmab *, int, double *, double *, int32_t *) { return 0; }
int bbgpu_cmab_adopt_basis(bbgpu_cmab *, int, const double *, const double *) { return 0; }
int bbgpu_eig_small_host(bbgpu_ctx *, int, int, const double *, double *, double *, int32_t *) { return 0; }
int bbgpu_stedc_host(bbgpu_ctx *, int, const double *, const double *, double *, double *, float *) { return 0; }
int bbgpu_sytrd_host(bbgpu_ctx *, int, const double *, double *, double *, double *, double *, int, float *) { return 0; }
}

/* SaNSDE / SSDE entry points: np evaluations per generation, nothing accepted, radius statistic 1 */
struct bbgpu_sde { int n, np; };
int bbgpu_sde_create(bbgpu_ctx *, int n, int np, int, uint64_t, const double *, const double *, bbgpu_sde **out) { *out = new bbgpu_sde { n, np }; return 0; }
int bbgpu_sde_destroy(bbgpu_sde *h) { delete h; return 0; }
int bbgpu_sde_init(bbgpu_sde *, bbgpu_objective *, const double *) { return 0; }
int bbgpu_sde_generation(bbgpu_sde *, bbgpu_objective *, int, double, double, double, const bbgpu_sde_draws *, bbgpu_sde_accum *a, int *r) { a->rmean = 1; a->rm2 = 1; if (r) *r = 1; return 0; }
int bbgpu_sde_sort(bbgpu_sde *) { return 0; }
int bbgpu_sde_get_swarm(bbgpu_sde *, double *, double *, double *) { return 0; }
int bbgpu_sde_get_trace(bbgpu_sde *, int32_t *, double *, double *) { return 0; }
int bbgpu_sde_radius(bbgpu_sde *, double *m, double *m2) { *m = 1; *m2 = 1; return 0; }
int bbgpu_sde_evaluations(bbgpu_sde *) { return 0; }
struct bbgpu_ssde { int n, np, np0; };
int bbgpu_ssde_create(bbgpu_ctx *, int n, int np, int, int, int, uint64_t, const double *, const double *, bbgpu_ssde **out) { *out = new bbgpu_ssde { n, np, np }; return 0; }
int bbgpu_ssde_destroy(bbgpu_ssde *h) { delete h; return 0; }
int bbgpu_ssde_init(bbgpu_ssde *h, bbgpu_objective *, const double *, int *fev) { *fev = h->np; return 0; }
int bbgpu_ssde_generation(bbgpu_ssde *h, bbgpu_objective *, const int32_t *, int fev, int mfev, double, int npmin, const bbgpu_ssde_draws *, bbgpu_ssde_state *s, int *r) {
    s->k = s->kcr = 1; s->evals = h->np; s->nsucc = 0;
    const int npnew = static_cast<int>(h->np0 + (npmin - h->np0) * (1. * (fev + h->np)) / mfev);
    if (h->np > npnew) h->np = npnew < 1 ? 1 : npnew;
    s->np_next = h->np; s->fbest = 1.0; s->rmean = 1.0; s->rm2 = 1.0;
    if (r) *r = 1;
    return 0;
}
int bbgpu_ssde_set_population(bbgpu_ssde *h, int np) { h->np = np; return 0; }
int bbgpu_ssde_get_swarm(bbgpu_ssde *, double *, double *) { return 0; }
int bbgpu_ssde_get_memory(bbgpu_ssde *, double *, double *, double *) { return 0; }
int bbgpu_ssde_get_trace(bbgpu_ssde *, int32_t *, double *) { return 0; }
int bbgpu_ssde_radius(bbgpu_ssde *, double *m, double *m2) { *m = 1; *m2 = 1; return 0; }
