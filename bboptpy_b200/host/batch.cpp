#include "batch.h"

#include <cmath>
#include <cstring>
#include <iostream>
#include <limits>
#include <memory>
#include <stdexcept>
#include <string>

#include "random.hpp"

using bbopt::gpu_check;

namespace {

// one run's restart bookkeeping, IPOP or BIPOP behind one face
struct RunSchedule {
    RestartBatch::Kind kind;
    IPopSchedule ipop;
    BiPopSchedule bipop;
    std::mt19937 rng;      // this run's host stream: x0 and the two small-regime draws (bipop_cmaes.cpp:112-114,241,244)
    int base_runs = 0;
    bool finished = false;
    base_run_request first(int n) { return kind == RestartBatch::IPOP ? ipop.first(n) : bipop.first(n); }
    base_run_request next(const double *lo, const double *hi, std::vector<double> &x0) {
        const uniform_fn u = [this](double a, double b) { return std::uniform_real_distribution<double>(a, b)(rng); };
        return kind == RestartBatch::IPOP ? ipop.next(lo, hi, x0, u) : bipop.next(lo, hi, x0, u);
    }
    bool finish(int fev, double fx) { return kind == RestartBatch::IPOP ? ipop.finish(fev, fx) : bipop.finish(fev, fx); }
    // the loop exits of optimize(): IPOP tests before every iterate() (ipop_cmaes.cpp:171), BIPOP iterates at least
    // once and tests after (bipop_cmaes.cpp:173-186)
    bool done() const {
        if (kind == RestartBatch::IPOP) return ipop.done();
        return base_runs >= 2 && bipop.done();
    }
    int fev() const { return kind == RestartBatch::IPOP ? ipop._fev : bipop._fev; }
};

// BaseCmaes::init + Cmaes::init constants of one base run (base_cmaes.cpp:71-129, cmaes.cpp:48); `weights` must
// outlive the bbgpu_cmab_start call
void base_run_params(int n, const base_run_request &req, double tol, double eigenrate, uint64_t seed,
        std::vector<double> &weights, bbgpu_cmab_run_params &p) {
    const int lambda = req.lambda, mu = lambda / 2;
    weights.assign(mu, 0.);
    double sum = 0.;
    for (int i = 0; i < mu; i++) {
        weights[i] = std::log(0.5 * (lambda + 1.)) - std::log(i + 1.);
        sum += weights[i];
    }
    const double inv = 1. / sum;
    for (int i = 0; i < mu; i++) weights[i] *= inv;
    double lenw = 0.;
    for (int i = 0; i < mu; i++) lenw = lenw + weights[i] * weights[i];
    const double mueff = 1. / lenw;
    std::memset(&p, 0, sizeof(p));
    p.lambda = lambda;
    p.mu = mu;
    p.mfev = req.mfev;
    p.mit = req.mfev / lambda;
    p.hlen = 10 + static_cast<int>(std::ceil((30. * n) / lambda));
    p.ik = static_cast<int>(std::ceil(0.1 + lambda / 4.));
    p.sigma0 = req.sigma;
    p.tol = tol;
    p.mueff = mueff;
    p.chi = std::sqrt(n) * (1. - 1. / (4. * n) + 1. / (21. * n * n));
    p.cc = (4. + mueff / n) / (n + 4. + 2. * mueff / n);
    p.cs = (mueff + 2.) / (5. + n + mueff);
    p.c1 = 2. / ((1.3 + n) * (1.3 + n) + mueff);
    p.cmu = std::min(1. - p.c1, 2. * (mueff - 2. + 1. / mueff) / ((2. + n) * (2. + n) + mueff));
    p.damps = 1. + p.cs + 2. * std::max(0., std::sqrt((mueff - 1.) / (n + 1.)) - 1.);
    p.eigenfreq = eigenrate * lambda / (p.c1 + p.cmu) / n;
    p.seed = seed;
    p.weights = weights.data();
}

} // namespace

// one stream's share of the local runs: its own context (stream), engine handle and staging
struct RestartBatch::Group {
    bbgpu_ctx *ctx = nullptr;
    bool own_ctx = false;
    bbgpu_cmab *dev = nullptr;
    std::vector<int> members;                 // local run index of in-group run q
    std::vector<int32_t> active, start_ids;
    std::vector<bbgpu_cmab_run_params> params;
    std::vector<double> start_x, xs;
    std::vector<bbgpu_cmab_run_status> status;
    bool in_flight = false, first_wave = true;
    ~Group() {
        if (dev) bbgpu_cmab_destroy(dev);
        if (own_ctx && ctx) bbgpu_ctx_destroy(ctx);
    }
};

RestartBatch::RestartBatch(Kind kind, int runs, int mfev, double tol, double sigma0, double eigenrate) :
        _kind(kind), _runs(runs), _mfev(mfev), _tol(tol), _sigma0(sigma0), _eigenrate(eigenrate) {
    if (runs < 1) throw bbopt::GpuError("RestartBatch: runs must be >= 1");
}

RestartBatch::~RestartBatch() {
}

void RestartBatch::configureIPop(bool nipop, double ksigmadec, bool boundlambda) {
    _nipop = nipop;
    _ksigmadec = ksigmadec;
    _boundlambda = boundlambda;
}

void RestartBatch::configureBiPop(int maxlargeruns, bool nbipop, double ksigmadec, double kbudget) {
    _maxlargeruns = maxlargeruns;
    _nbipop = nbipop;
    _ksigmadec = ksigmadec;
    _kbudget = kbudget;
}

// The batched engine keeps a run inside one CTA: n is capped by the CTA's thread layout and lambda by the shared memory of
// the in-CTA ranking (bbgpu_cmab_limits).  n beyond the cap is rejected here, BEFORE any compute.  The restart schedules
// grow lambda (IPOP: doubling up to 10 n^2, ipop_cmaes.cpp:81,120-131; BIPOP: lambdadef 2^largerestarts,
// bipop_cmaes.cpp:208), in theory past the cap; a sequence whose NEXT base run would exceed it ends there with what it
// has (one warning per batch) instead of failing the whole batch in the middle and dropping every run's result.
int RestartBatch::engineMaxLambda(int n) const {
    int max_n = 0, max_lambda = 0, max_hlen = 0;
    gpu_check(bbgpu_cmab_limits(n, &max_n, &max_lambda, &max_hlen));
    if (n > max_n)
        throw std::invalid_argument("restart batch: n = " + std::to_string(n) + " exceeds the batched engine's limit of " +
                std::to_string(max_n) + "; use IPopCMAES / BiPopCMAES (one-handle engine) for larger problems");
    const int lambdadef = 4 + static_cast<int>(3. * std::log(1. * n));
    const int hlen = 10 + static_cast<int>(std::ceil((30. * n) / lambdadef));
    if (hlen > max_hlen || lambdadef > max_lambda)
        throw std::invalid_argument("restart batch: the default run at n = " + std::to_string(n) + " does not fit the batched engine");
    return max_lambda;
}

std::vector<multivariate_solution> RestartBatch::optimize(const multivariate_problem &f, const double *guesses) {
    if (f._hasc || f._hasbbc) {
        std::cerr << "Warning [CMA batch]: problem constraints will be ignored." << std::endl;
    }
    _n = f._n;
    const int n = _n;
    bbgpu_ctx *ctx0 = bbopt::device_context();
    const int max_lambda = engineMaxLambda(n);
    bool warned_lambda = false;
    const int world = _comm ? bbgpu_comm_world(_comm) : 1, rank = _comm ? bbgpu_comm_rank(_comm) : 0;
    _ids.clear();
    for (int r = rank; r < _runs; r += world) _ids.push_back(r);
    const int L = (int) _ids.size();
    _stats = restart_batch_stats();
    _fbest.assign(L, std::numeric_limits<double>::infinity());
    _xbest.assign((size_t) L * n, 0.);
    _fev.assign(L, 0);
    if (L == 0) return {};
    bbopt::ObjectiveHandle obj;
    obj.bind(f);

    // Streams: the runs are dealt to G groups, each with its own stream and engine handle.  While the host reads one
    // group's status words and restarts its stopped runs, the other groups' kernels keep the GPU full -- and the tail
    // of one launch (its last CTAs draining) overlaps the head of the next group's.
    const int G = _streams > 0 ? std::min(_streams, L) : (L >= 1024 ? 2 : 1);
    std::vector<std::unique_ptr<Group>> groups;
    for (int g = 0; g < G; g++) {
        groups.emplace_back(new Group());
        Group &gr = *groups.back();
        if (g == 0) {
            gr.ctx = ctx0;
        } else {
            gpu_check(bbgpu_ctx_create(bbgpu_ctx_device(ctx0), nullptr, &gr.ctx));
            gr.own_ctx = true;
        }
        for (int l = g; l < L; l += G) gr.members.push_back(l);
        const int Lg = (int) gr.members.size();
        gpu_check(bbgpu_cmab_create(gr.ctx, n, Lg, 0, f._lower, f._upper, &gr.dev));
        gr.status.resize(Lg);
        gr.xs.resize((size_t) Lg * n);
        if (_timing) {
            gpu_check(bbgpu_ctx_profile_enable(gr.ctx, 1));
            gpu_check(bbgpu_ctx_profile_reset(gr.ctx));
        }
    }

    const uint64_t key = bbopt::Random::next_stream_key();
    std::vector<RunSchedule> sched(L);
    std::vector<std::vector<double>> weights(L);
    std::vector<double> x0;
    auto queue_start = [&](Group &gr, int q, const base_run_request &req, const double *x) {
        const int l = gr.members[q];
        RunSchedule &s = sched[l];
        bbgpu_cmab_run_params p;
        // a distinct Philox key per base run: (batch key, global run id, restart index)
        const uint64_t seed = key + 0x9E3779B97F4A7C15ull * (uint64_t) (_ids[l] + 1) + 0xD1B54A32D192ED03ull * (uint64_t) s.base_runs;
        base_run_params(n, req, _tol, _eigenrate, seed, weights[l], p);
        gr.params.push_back(p);
        gr.start_ids.push_back(q);
        gr.start_x.insert(gr.start_x.end(), x, x + n);
        s.base_runs++;
        _stats.base_runs++;
    };
    // start what is queued, then launch the group's next step (async)
    auto launch = [&](Group &gr) {
        if (!gr.start_ids.empty()) {
            // initial runs start from clean matrices; restarts keep the off-diagonals of B and C (cmaes.cpp:53-60)
            gpu_check(bbgpu_cmab_start(gr.dev, (int) gr.start_ids.size(), gr.start_ids.data(), gr.params.data(),
                    gr.start_x.data(), gr.first_wave ? 0 : 1));
            for (int32_t q : gr.start_ids) gr.active.push_back(q);
            gr.start_ids.clear();
            gr.params.clear();
            gr.start_x.clear();
            gr.first_wave = false;
        }
        gr.in_flight = !gr.active.empty();
        if (!gr.in_flight) return;
        gpu_check(bbgpu_cmab_step(gr.dev, obj.get(), (int) gr.active.size(), gr.active.data(), _quantum, 1 << 30, nullptr, 0));
        _stats.launches++;
    };
    // wait for the group's step, book the runs that stopped, queue their restarts
    auto collect = [&](Group &gr) {
        gpu_check(bbgpu_cmab_fetch(gr.dev, gr.status.data()));
        bool any_stopped = false;
        for (int32_t q : gr.active) {
            _stats.generations += gr.status[q].generations_this_step;
            if (gr.status[q].status != 0) any_stopped = true;
        }
        if (!any_stopped) return;
        gpu_check(bbgpu_cmab_get_best(gr.dev, 0, (int) gr.members.size(), gr.xs.data()));
        std::vector<int32_t> still;
        for (int32_t q : gr.active) {
            const bbgpu_cmab_run_status &st = gr.status[q];
            if (st.status == 0) {
                still.push_back(q);
                continue;
            }
            if (st.status == 3) throw bbopt::GpuError("RestartBatch: the eigendecomposition of a run failed (non-finite covariance)");
            const int l = gr.members[q];
            RunSchedule &s = sched[l];
            _stats.decompositions += st.eig_count;
            if (s.finish(st.fev, st.fx_final)) {
                _fbest[l] = st.fx_final;
                std::memcpy(&_xbest[(size_t) l * n], &gr.xs[(size_t) q * n], sizeof(double) * n);
            }
            if (s.done()) {
                s.finished = true;
                _fev[l] = s.fev();
                _stats.evaluations += s.fev();
            } else {
                const base_run_request req = s.next(f._lower, f._upper, x0); // fills x0
                if (req.lambda > max_lambda) { // see engineMaxLambda(): end this sequence with what it has
                    if (!warned_lambda) {
                        std::cerr << "Warning [CMA batch]: a restart asked for lambda = " << req.lambda << " > " << max_lambda
                                  << " (batched engine limit at n = " << n << "); such sequences stop early." << std::endl;
                        warned_lambda = true;
                    }
                    s.finished = true;
                    _fev[l] = s.fev();
                    _stats.evaluations += s.fev();
                    continue;
                }
                queue_start(gr, q, req, x0.data());
            }
        }
        gr.active.swap(still);
    };
    for (int l = 0; l < L; l++) {
        RunSchedule &s = sched[l];
        s.kind = _kind;
        s.ipop.configure(_mfev, _sigma0, _nipop, _ksigmadec, _boundlambda);
        s.bipop.configure(_mfev, _sigma0, _maxlargeruns, _nbipop, _ksigmadec, _kbudget);
        std::seed_seq seq { (uint32_t) key, (uint32_t) (key >> 32), (uint32_t) _ids[l], (uint32_t) 0x62617463u };
        s.rng.seed(seq);
    }
    for (auto &gp : groups) {
        Group &gr = *gp;
        for (int q = 0; q < (int) gr.members.size(); q++) {
            const int l = gr.members[q];
            queue_start(gr, q, sched[l].first(n), guesses + (size_t) _ids[l] * n);
        }
        launch(gr);
    }
    while (true) {
        bool any = false;
        for (auto &gp : groups) {
            Group &gr = *gp;
            if (!gr.in_flight) continue;
            any = true;
            collect(gr);
            launch(gr);
        }
        if (!any) break;
    }
    if (_timing) {
        for (auto &gp : groups) {
            double ms = 0.;
            int64_t cnt = 0;
            gpu_check(bbgpu_ctx_profile_read(gp->ctx, BBGPU_PROF_CMAB_STEP, &ms, &cnt));
            _stats.device_ms += ms;
            gpu_check(bbgpu_ctx_profile_enable(gp->ctx, 0));
        }
    }
    groups.clear();
    std::vector<multivariate_solution> out;
    out.reserve(L);
    for (int l = 0; l < L; l++)
        out.emplace_back(std::vector<double>(_xbest.begin() + (size_t) l * n, _xbest.begin() + (size_t) (l + 1) * n), _fev[l], false);
    return out;
}

double RestartBatch::bestOf(std::vector<double> &x, int *owner) {
    int arg = -1;
    double fb = std::numeric_limits<double>::infinity();
    for (size_t l = 0; l < _fbest.size(); l++)
        if (_fbest[l] < fb) { fb = _fbest[l]; arg = (int) l; }
    x.assign(_n, 0.);
    if (arg >= 0) std::memcpy(x.data(), &_xbest[(size_t) arg * _n], sizeof(double) * _n);
    int who = arg >= 0 ? _ids[arg] : -1;
    if (_comm && bbgpu_comm_world(_comm) > 1) {
        // the winner's global run id travels as one more coordinate of its solution vector
        std::vector<double> xe(x);
        xe.push_back((double) who);
        int rank_owner = 0;
        gpu_check(bbgpu_comm_best_of_host(_comm, fb, xe.data(), _n + 1, &fb, &rank_owner));
        x.assign(xe.begin(), xe.begin() + _n);
        who = (int) xe[_n];
    }
    if (owner) *owner = who;
    return fb;
}
