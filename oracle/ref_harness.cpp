/*
 * ORACLE (test infrastructure).  Thin C driver around the UNMODIFIED reference sources.
 *
 * Built by oracle/Makefile: this file + /root/reference/src/{blas.cpp,
 * multivariate/cma/{base_cmaes,cmaes,active_cmaes,sep_cmaes,cholesky_cmaes,lm_cmaes,ipop_cmaes,bipop_cmaes}.cpp,
 * multivariate/de/{jade,shade,sansde,ssde}.cpp, multivariate/nes/xnes.cpp} compiled where they lie -> oracle/_ref/libbbo_ref.so.
 * No reference source is copied into this repository; the headers are included by path.
 *
 * State is reached without patching the reference: every member is `protected`
 * (base_cmaes.h:52-63, cmaes.h:42-46, active_cmaes.h:42-44, sep_cmaes.h:39-43,
 * jade.h:51-58, shade.h:44-51), so the probes below are subclasses that only add getters.
 * Determinism: effolkronium::random_static::seed (random.hpp:238-241) on the one
 * process-global engine (random.hpp:677-680).
 */
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstring>
#include <functional>
#include <random>
#include <string>
#include <vector>

#include "multivariate/multivariate.h"
#include "multivariate/cma/base_cmaes.h"
#include "multivariate/cma/cmaes.h"
#include "multivariate/cma/active_cmaes.h"
#include "multivariate/cma/sep_cmaes.h"
#include "multivariate/cma/cholesky_cmaes.h"
#include "multivariate/cma/lm_cmaes.h"
#include "multivariate/cma/ipop_cmaes.h"
#include "multivariate/cma/bipop_cmaes.h"
#include "multivariate/de/jade.h"
#include "multivariate/de/shade.h"
#include "multivariate/de/sansde.h"
#include "multivariate/de/ssde.h"
#include "multivariate/nes/xnes.h"
#include "random.hpp"

#include "bbo_api.h"
#include "objectives_ref.h"

using Random = effolkronium::random_static;

namespace {

multivariate make_objective(int obj, int n) {
    std::vector<double> aux;
    if (obj == BBO_ELLIPSOID) {
        aux.resize(n);
        bbo_ellipsoid_coef(n, aux.data());
    }
    return [obj, n, aux](const double *x) -> double {
        return bbo_objective(obj, n, x, aux.empty() ? nullptr : aux.data());
    };
}

int copy_out(const std::vector<double> &v, double *out, int cap) {
    const int len = (int) v.size();
    for (int i = 0; i < len && i < cap; i++) out[i] = v[i];
    return len;
}

int copy_out2(const std::vector<std::vector<double>> &m, double *out, int cap) {
    int len = 0;
    for (const auto &row : m) {
        for (double v : row) {
            if (len < cap) out[len] = v;
            len++;
        }
    }
    return len;
}

/* one probe type per concrete class; a tiny virtual interface keeps the C API generic */
struct CmaProbe {
    virtual ~CmaProbe() {}
    virtual BaseCmaes *base() = 0;
    virtual int geti(const std::string &k) = 0;
    virtual double getd(const std::string &k) = 0;
    virtual int getv(const std::string &k, double *out, int cap) = 0;
    virtual void reset_z() = 0;
    virtual double time_eig() { return 0.0; }
    std::vector<double> lower, upper, guess;
    multivariate_problem prob;
    /* eigenbasis-replay feed (bbo_cma_eig_log): filled by FullProbe::iterate */
    struct EigEvent { int fev, n; std::vector<double> B, D; };
    bool eiglog = false;
    std::vector<EigEvent> eigevents;
};

#define BASE_GETI(k)                                                                       \
    if (k == "n") return this->_n;                                                         \
    if (k == "lambda") return this->_lambda;                                               \
    if (k == "mu") return this->_mu;                                                       \
    if (k == "it") return this->_it;                                                       \
    if (k == "fev") return this->_fev;                                                     \
    if (k == "mfev") return this->_mfev;                                                   \
    if (k == "mit") return this->_mit;                                                     \
    if (k == "hlen") return this->_hlen;                                                   \
    if (k == "ik") return this->_ik;                                                       \
    if (k == "bound") return this->_bound ? 1 : 0;                                         \
    if (k == "best_len") return this->_best._len;
#define FLAG_GETI(k)                                                                       \
    if (k == "flag") return this->_flag;

#define BASE_GETD(k)                                                                       \
    if (k == "sigma") return this->_sigma;                                                 \
    if (k == "sigma0") return this->_sigma0;                                               \
    if (k == "tol") return this->_tol;                                                     \
    if (k == "mueff") return this->_mueff;                                                 \
    if (k == "cc") return this->_cc;                                                       \
    if (k == "cs") return this->_cs;                                                       \
    if (k == "c1") return this->_c1;                                                       \
    if (k == "cmu") return this->_cmu;                                                     \
    if (k == "damps") return this->_damps;                                                 \
    if (k == "chi") return this->_chi;                                                     \
    if (k == "fbest") return this->_fbest;                                                 \
    if (k == "fworst") return this->_fworst;

#define BASE_GETV(k)                                                                       \
    if (k == "xmean") return copy_out(this->_xmean, out, cap);                             \
    if (k == "xold") return copy_out(this->_xold, out, cap);                               \
    if (k == "pc") return copy_out(this->_pc, out, cap);                                   \
    if (k == "ps") return copy_out(this->_ps, out, cap);                                   \
    if (k == "weights") return copy_out(this->_weights, out, cap);                         \
    if (k == "ybw") return copy_out(this->_ybw, out, cap);                                 \
    if (k == "arx") return copy_out2(this->_arx, out, cap);                                \
    if (k == "hist_best") return copy_out(this->_best._hist, out, cap);                    \
    if (k == "hist_kth") return copy_out(this->_kth._hist, out, cap);                      \
    if (k == "ibw") {                                                                      \
        for (int i = 0; i < 4 && i < cap; i++) out[i] = this->_ibw[i];                     \
        return 4;                                                                          \
    }                                                                                      \
    if (k == "fit_val") {                                                                  \
        for (int i = 0; i < this->_lambda && i < cap; i++) out[i] = this->_fitness[i]._value; \
        return this->_lambda;                                                              \
    }                                                                                      \
    if (k == "fit_idx") {                                                                  \
        for (int i = 0; i < this->_lambda && i < cap; i++) out[i] = this->_fitness[i]._index; \
        return this->_lambda;                                                              \
    }

template<class T> struct FullProbe: public T, public CmaProbe {
    using T::T;
    BaseCmaes *base() override { return this; }
    void reset_z() override { this->_Z.reset(); }
    /* init()/iterate() are virtual in MultivariateOptimizer (multivariate.h:132-146), so the reference's own
     * optimize() loop and its IPOP/BIPOP drivers dispatch here: record the basis whenever the decomposition fired */
    int _eiglast = 0;
    void init(const multivariate_problem &f, const double *guess) override {
        T::init(f, guess);
        _eiglast = this->_eigenlastev;
    }
    void iterate() override {
        T::iterate();
        if (this->_eigenlastev != _eiglast) {
            _eiglast = this->_eigenlastev;
            if (eiglog) {
                EigEvent e { this->_eigenlastev, this->_n, {}, this->_diagd };
                for (const auto &row : this->_b) e.B.insert(e.B.end(), row.begin(), row.end());
                eigevents.push_back(std::move(e));
            }
        }
    }
    double time_eig() override {
        this->_eigenlastev = this->_fev - static_cast<int>(this->_eigenfreq) - 1;
        const auto t0 = std::chrono::steady_clock::now();
        this->eigenDecomposition();
        return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    }
    int geti(const std::string &k) override {
        BASE_GETI(k)
        FLAG_GETI(k)
        if (k == "eigenlastev") return this->_eigenlastev;
        return -1;
    }
    double getd(const std::string &k) override {
        BASE_GETD(k)
        if (k == "eigenfreq") return this->_eigenfreq;
        return extra_d(k);
    }
    int getv(const std::string &k, double *out, int cap) override {
        BASE_GETV(k)
        if (k == "D") return copy_out(this->_diagd, out, cap);
        if (k == "B") return copy_out2(this->_b, out, cap);
        if (k == "C") return copy_out2(this->_c, out, cap);
        if (k == "invsqrtC") return copy_out2(this->_invsqrtc, out, cap);
        return extra_v(k, out, cap);
    }
    virtual double extra_d(const std::string &) { return NAN; }
    virtual int extra_v(const std::string &, double *, int) { return -1; }
};

struct VanillaProbe: public FullProbe<Cmaes> {
    using FullProbe<Cmaes>::FullProbe;
};

struct ActiveProbe: public FullProbe<ActiveCmaes> {
    using FullProbe<ActiveCmaes>::FullProbe;
    double extra_d(const std::string &k) override {
        if (k == "cneg") return _cneg;
        if (k == "cm") return _cm;
        if (k == "alphaold") return _alphaold;
        return NAN;
    }
    int extra_v(const std::string &k, double *out, int cap) override {
        if (k == "ycoeff") return copy_out(_ycoeff, out, cap);
        return -1;
    }
};

struct SepProbe: public SepCmaes, public CmaProbe {
    using SepCmaes::SepCmaes;
    BaseCmaes *base() override { return this; }
    void reset_z() override { _Z.reset(); }
    int geti(const std::string &k) override {
        BASE_GETI(k)
        FLAG_GETI(k)
        return -1;
    }
    double getd(const std::string &k) override {
        BASE_GETD(k)
        if (k == "ccov") return _ccov;
        return NAN;
    }
    int getv(const std::string &k, double *out, int cap) override {
        BASE_GETV(k)
        if (k == "D") return copy_out(_diagd, out, cap);
        if (k == "C") return copy_out(_c, out, cap);
        return -1;
    }
};

/* CholeskyCmaes (cholesky_cmaes.h:34-37: _stol, _dmean, _a, _mattmp are protected); it has no _flag */
struct CholProbe: public CholeskyCmaes, public CmaProbe {
    using CholeskyCmaes::CholeskyCmaes;
    BaseCmaes *base() override { return this; }
    void reset_z() override { _Z.reset(); }
    int geti(const std::string &k) override {
        BASE_GETI(k)
        if (k == "flag") return 0;
        return -1;
    }
    double getd(const std::string &k) override {
        BASE_GETD(k)
        if (k == "stol") return _stol;
        return NAN;
    }
    int getv(const std::string &k, double *out, int cap) override {
        BASE_GETV(k)
        if (k == "A") return copy_out2(_a, out, cap);
        if (k == "dmean") return copy_out(_dmean, out, cap);
        return -1;
    }
};

/* LmCmaes (lm_cmaes.h:38-46: everything protected) */
struct LmProbe: public LmCmaes, public CmaProbe {
    using LmCmaes::LmCmaes;
    BaseCmaes *base() override { return this; }
    void reset_z() override { _Z.reset(); }
    int geti(const std::string &k) override {
        BASE_GETI(k)
        if (k == "flag") return 0;
        if (k == "memlen") return _memlen;
        if (k == "memsize") return _memsize;
        if (k == "nsteps") return _nsteps;
        if (k == "t") return _t;
        return -1;
    }
    double getd(const std::string &k) override {
        BASE_GETD(k)
        if (k == "s") return _s;
        if (k == "ccc") return _ccc;
        if (k == "sqrt1mc1") return _sqrt1mc1;
        return NAN;
    }
    int getv(const std::string &k, double *out, int cap) override {
        BASE_GETV(k)
        if (k == "pcmat") return copy_out2(_pcmat, out, cap);
        if (k == "vmat") return copy_out2(_vmat, out, cap);
        if (k == "b") return copy_out(_b, out, cap);
        if (k == "d") return copy_out(_d, out, cap);
        if (k == "fp") return copy_out(_fp, out, cap);
        if (k == "jarr" || k == "larr") {
            const std::vector<int> &v = k == "jarr" ? _jarr : _larr;
            for (int i = 0; i < (int) v.size() && i < cap; i++) out[i] = v[i];
            return (int) v.size();
        }
        return -1;
    }
};

void bind_problem(CmaProbe *p, int obj, int n, const double *lower, const double *upper,
        const double *guess) {
    p->lower.assign(lower, lower + n);
    p->upper.assign(upper, upper + n);
    p->guess.assign(guess, guess + n);
    p->prob = multivariate_problem(make_objective(obj, n), n, p->lower.data(),
            p->upper.data());
}

/* ---- DE probes */
struct DeProbe {
    virtual ~DeProbe() {}
    virtual MultivariateOptimizer *opt() = 0;
    virtual int geti(const std::string &k) = 0;
    virtual double getd(const std::string &k) = 0;
    virtual int getv(const std::string &k, double *out, int cap) = 0;
    std::vector<double> lower, upper, guess;
    multivariate_problem prob;
};

#define DE_GETV(k)                                                                         \
    if (k == "swarm_f") {                                                                  \
        for (int i = 0; i < (int) this->_swarm.size() && i < cap; i++)                     \
            out[i] = this->_swarm[i]._f;                                                   \
        return (int) this->_swarm.size();                                                  \
    }                                                                                      \
    if (k == "swarm_x") {                                                                  \
        int len = 0;                                                                       \
        for (const auto &pt : this->_swarm)                                                \
            for (double v : pt._x) {                                                       \
                if (len < cap) out[len] = v;                                               \
                len++;                                                                     \
            }                                                                              \
        return len;                                                                        \
    }                                                                                      \
    if (k == "arch") return copy_out2(this->_arch, out, cap);

struct JadeProbe: public JadeSearch, public DeProbe {
    using JadeSearch::JadeSearch;
    MultivariateOptimizer *opt() override { return this; }
    int geti(const std::string &k) override {
        if (k == "n") return _n;
        if (k == "np") return _np;
        if (k == "fev") return _fev;
        if (k == "larch") return (int) _arch.size();
        return -1;
    }
    double getd(const std::string &k) override {
        if (k == "mucr") return _mucr;
        if (k == "muf") return _muf;
        return NAN;
    }
    int getv(const std::string &k, double *out, int cap) override {
        DE_GETV(k)
        return -1;
    }
};

struct ShadeProbe: public ShadeSearch, public DeProbe {
    using ShadeSearch::ShadeSearch;
    MultivariateOptimizer *opt() override { return this; }
    int geti(const std::string &k) override {
        if (k == "n") return _n;
        if (k == "np") return _np;
        if (k == "fev") return _fev;
        if (k == "larch") return (int) _arch.size();
        if (k == "k") return _k;
        return -1;
    }
    double getd(const std::string &) override { return NAN; }
    int getv(const std::string &k, double *out, int cap) override {
        DE_GETV(k)
        if (k == "MCR") return copy_out(_MCR, out, cap);
        if (k == "MF") return copy_out(_MF, out, cap);
        return -1;
    }
};

/* xNES (nes/xnes.h:50-57: state is protected) */
struct XnesProbe: public xNES {
    using xNES::xNES;
    std::vector<double> lower, upper, guess;
    multivariate_problem prob;
    int geti(const std::string &k) {
        if (k == "n") return _n;
        if (k == "np") return _np;
        if (k == "fev") return _fev;
        return -1;
    }
    double getd(const std::string &k) {
        if (k == "sigma") return _sigma;
        if (k == "etasigma") return _etasigma;
        if (k == "etab") return _etab;
        if (k == "etamu") return _etamu;
        if (k == "Gsigma") return _Gsigma;
        return NAN;
    }
    int getv(const std::string &k, double *out, int cap) {
        if (k == "u") return copy_out(_u, out, cap);
        if (k == "mu") return copy_out(_mu, out, cap);
        if (k == "B") return copy_out2(_B, out, cap);
        if (k == "Gdelta") return copy_out(_Gdelta, out, cap);
        if (k == "x" || k == "z") {
            int len = 0;
            for (const auto &pt : _points)
                for (double v : (k == "x" ? pt._x : pt._z)) {
                    if (len < cap) out[len] = v;
                    len++;
                }
            return len;
        }
        if (k == "f") {
            for (int i = 0; i < (int) _points.size() && i < cap; i++) out[i] = _points[i]._f;
            return (int) _points.size();
        }
        return -1;
    }
};

/* SaNSDE (de/sansde.h:58-66: state is protected; the point type is private but its members are reachable) */
struct SansdeProbe: public SaNSDESearch {
    using SaNSDESearch::SaNSDESearch;
    std::vector<double> lower, upper, guess;
    multivariate_problem prob;
    int geti(const std::string &k) {
        if (k == "n") return _n;
        if (k == "np") return _np;
        if (k == "fev") return _fev;
        if (k == "it") return _it;
        if (k == "pns0") return _pns[0];
        if (k == "pns1") return _pns[1];
        if (k == "pnf0") return _pnf[0];
        if (k == "pnf1") return _pnf[1];
        return -1;
    }
    double getd(const std::string &k) {
        if (k == "p") return _p;
        if (k == "fp") return _fp;
        if (k == "crm") return _crm;
        if (k == "crrec") return _crrec;
        if (k == "crdeltaf") return _crdeltaf;
        if (k == "fpns0") return _fpns[0];
        if (k == "fpns1") return _fpns[1];
        if (k == "fpnf0") return _fpnf[0];
        if (k == "fpnf1") return _fpnf[1];
        return NAN;
    }
    int getv(const std::string &k, double *out, int cap) {
        int len = 0;
        if (k == "swarm_x") {
            for (const auto &pt : _swarm)
                for (double v : pt._x) {
                    if (len < cap) out[len] = v;
                    len++;
                }
            return len;
        }
        if (k == "swarm_f" || k == "swarm_cr") {
            for (const auto &pt : _swarm) {
                if (len < cap) out[len] = k == "swarm_f" ? pt._f : pt._cr;
                len++;
            }
            return len;
        }
        return -1;
    }
};

/* SSDE (de/ssde.h:44-58: state is protected) */
struct SsdeProbe: public SSDESearch {
    using SSDESearch::SSDESearch;
    std::vector<double> lower, upper, guess;
    multivariate_problem prob;
    int geti(const std::string &k) {
        if (k == "n") return _n;
        if (k == "np") return _np;
        if (k == "fev") return _fev;
        if (k == "convcount") return _convcount;
        if (k == "k") return _k;
        if (k == "kcr") return _kCR;
        return -1;
    }
    double getd(const std::string &) { return NAN; }
    int getv(const std::string &k, double *out, int cap) {
        int len = 0;
        if (k == "swarm_x") {
            for (const auto &pt : _swarm)
                for (double v : pt._x) {
                    if (len < cap) out[len] = v;
                    len++;
                }
            return len;
        }
        if (k == "swarm_f") {
            for (const auto &pt : _swarm) {
                if (len < cap) out[len] = pt._f;
                len++;
            }
            return len;
        }
        if (k == "L1") return copy_out(_L1, out, cap);
        if (k == "L2") return copy_out(_L2, out, cap);
        if (k == "LCR") return copy_out(_LCR, out, cap);
        if (k == "iwork") {
            for (int i = 0; i < (int) _iwork.size() && i < cap; i++) out[i] = _iwork[i];
            return (int) _iwork.size();
        }
        return -1;
    }
};

std::normal_distribution<> g_normal { 0., 1. };

} // namespace

extern "C" {

const char *bbo_kind(void) { return "reference"; }

void bbo_seed(uint32_t s) {
    Random::seed(s);
    g_normal.reset();
}

void bbo_rng_draw(int kind, int count, double a, double b, double *out) {
    for (int i = 0; i < count; i++) {
        switch (kind) {
        case 0: out[i] = Random::get(g_normal); break;
        case 1: out[i] = Random::get(a, b); break;
        case 2: out[i] = (double) Random::get((int) a, (int) b); break;
        default: out[i] = (double) Random::get<unsigned int>(0u, 0xffffffffu); break;
        }
    }
}

double bbo_eval(int obj, int n, const double *x) {
    return make_objective(obj, n)(x);
}

void *bbo_cma_new(int kind, int mfev, double tol, int np, double sigma0, int bound,
        double eigenrate, double alphacov, int adjustlr) {
    CmaProbe *p = nullptr;
    if (kind == 0) p = new VanillaProbe(mfev, tol, np, sigma0, bound != 0, eigenrate);
    else if (kind == 1) p = new ActiveProbe(mfev, tol, np, sigma0, bound != 0, alphacov, eigenrate);
    else if (kind == 3) p = new CholProbe(mfev, tol, /*stol=*/eigenrate, np, sigma0, bound != 0);
    else if (kind == 4) p = new LmProbe(mfev, tol, np, /*memory=*/(int) eigenrate, sigma0, bound != 0, /*rademacher=*/alphacov != 0.,
            /*usenew=*/adjustlr != 0);
    else p = new SepProbe(mfev, tol, np, sigma0, bound != 0, adjustlr != 0);
    return p;
}

void bbo_cma_free(void *h) { delete static_cast<CmaProbe*>(h); }

void bbo_cma_init(void *h, int obj, int n, const double *lower, const double *upper,
        const double *guess) {
    CmaProbe *p = static_cast<CmaProbe*>(h);
    bind_problem(p, obj, n, lower, upper, guess);
    p->base()->init(p->prob, p->guess.data());
}

void bbo_cma_set_params(void *h, int np, double sigma, int mfev) {
    static_cast<CmaProbe*>(h)->base()->setParams(np, sigma, mfev);
}

void bbo_cma_sample(void *h) { static_cast<CmaProbe*>(h)->base()->samplePopulation(); }
void bbo_cma_evalsort(void *h) { static_cast<CmaProbe*>(h)->base()->evaluateAndSortPopulation(); }
void bbo_cma_update(void *h) { static_cast<CmaProbe*>(h)->base()->updateDistribution(); }

namespace {
/* updateHistory(); _it++ -- the tail of BaseCmaes::iterate (base_cmaes.cpp:154-155).
 * _it is protected, so we reach it through a member-pointer obtained in a subclass. */
struct ItAccess: public BaseCmaes {
    static int BaseCmaes::* ptr() { return &ItAccess::_it; }
};
}

void bbo_cma_history(void *h) {
    BaseCmaes *b = static_cast<CmaProbe*>(h)->base();
    b->updateHistory();
    (b->*ItAccess::ptr())++;
}

void bbo_cma_iterate(void *h) { static_cast<CmaProbe*>(h)->base()->iterate(); }
int bbo_cma_converged(void *h) { return static_cast<CmaProbe*>(h)->base()->converged() ? 1 : 0; }

int bbo_cma_optimize(void *h, int obj, int n, const double *lower, const double *upper,
        const double *guess, double *xout, int *fev) {
    CmaProbe *p = static_cast<CmaProbe*>(h);
    bind_problem(p, obj, n, lower, upper, guess);
    multivariate_solution sol = p->base()->optimize(p->prob, p->guess.data());
    for (int i = 0; i < n; i++) xout[i] = sol._sol[i];
    *fev = sol._fev;
    return sol._converged ? 1 : 0;
}

int bbo_cma_geti(void *h, const char *name) { return static_cast<CmaProbe*>(h)->geti(name); }
double bbo_cma_getd(void *h, const char *name) { return static_cast<CmaProbe*>(h)->getd(name); }
int bbo_cma_getv(void *h, const char *name, double *out, int cap) {
    return static_cast<CmaProbe*>(h)->getv(name, out, cap);
}
void bbo_cma_reset_z(void *h) { static_cast<CmaProbe*>(h)->reset_z(); }
double bbo_cma_time_eig(void *h) { return static_cast<CmaProbe*>(h)->time_eig(); }

void bbo_cma_eig_log(void *h, int enable) {
    CmaProbe *p = static_cast<CmaProbe*>(h);
    p->eigevents.clear();
    p->eiglog = enable != 0;
}

int bbo_cma_eig_log_count(void *h) { return (int) static_cast<CmaProbe*>(h)->eigevents.size(); }

int bbo_cma_eig_log_get(void *h, int k, int *fev, double *B, double *D) {
    CmaProbe *p = static_cast<CmaProbe*>(h);
    if (k < 0 || k >= (int) p->eigevents.size()) return 0;
    const CmaProbe::EigEvent &e = p->eigevents[k];
    if (fev) *fev = e.fev;
    if (B) std::copy(e.B.begin(), e.B.end(), B);
    if (D) std::copy(e.D.begin(), e.D.begin() + e.n, D);
    return e.n;
}

int bbo_restart_optimize(int kind, void *base, int mfev, double sigma0, int flag,
        double ksigmadec, int boundlambda, int maxlargeruns, double kbudget, int obj, int n,
        const double *lower, const double *upper, const double *guess, double *xout, int *fev,
        double *trace, int cap) {
    (void) trace;
    (void) cap;
    CmaProbe *p = static_cast<CmaProbe*>(base);
    bind_problem(p, obj, n, lower, upper, guess);
    if (kind == 0) {
        IPopCmaes drv(p->base(), mfev, false, sigma0, flag != 0, ksigmadec, boundlambda != 0);
        multivariate_solution sol = drv.optimize(p->prob, p->guess.data());
        for (int i = 0; i < n; i++) xout[i] = sol._sol[i];
        *fev = sol._fev;
    } else {
        BiPopCmaes drv(p->base(), mfev, false, sigma0, maxlargeruns, flag != 0, ksigmadec,
                kbudget);
        multivariate_solution sol = drv.optimize(p->prob, p->guess.data());
        for (int i = 0; i < n; i++) xout[i] = sol._sol[i];
        *fev = sol._fev;
    }
    return 0; /* the reference exposes no per-restart trace */
}

void *bbo_xnes_new(int mfev, double tol, double a0, double etamu) { return new XnesProbe(mfev, tol, a0, etamu); }
void bbo_xnes_free(void *h) { delete static_cast<XnesProbe*>(h); }

static void xnes_bind(XnesProbe *p, int obj, int n, const double *lower, const double *upper, const double *guess) {
    p->lower.assign(lower, lower + n);
    p->upper.assign(upper, upper + n);
    p->guess.assign(guess, guess + n);
    p->prob = multivariate_problem(make_objective(obj, n), n, p->lower.data(), p->upper.data());
}

void bbo_xnes_init(void *h, int obj, int n, const double *lower, const double *upper, const double *guess) {
    XnesProbe *p = static_cast<XnesProbe*>(h);
    xnes_bind(p, obj, n, lower, upper, guess);
    p->init(p->prob, p->guess.data());
}

void bbo_xnes_iterate(void *h) { static_cast<XnesProbe*>(h)->iterate(); }
/* converged() is private in the reference; solution() reports it (xnes.cpp:193-195) */
int bbo_xnes_converged(void *h) { return static_cast<XnesProbe*>(h)->solution()._converged ? 1 : 0; }

int bbo_xnes_optimize(void *h, int obj, int n, const double *lower, const double *upper, const double *guess, double *xout,
        int *fev) {
    XnesProbe *p = static_cast<XnesProbe*>(h);
    xnes_bind(p, obj, n, lower, upper, guess);
    multivariate_solution sol = p->optimize(p->prob, p->guess.data());
    for (int i = 0; i < n; i++) xout[i] = sol._sol[i];
    *fev = sol._fev;
    return sol._converged ? 1 : 0;
}

int bbo_xnes_geti(void *h, const char *name) { return static_cast<XnesProbe*>(h)->geti(name); }
double bbo_xnes_getd(void *h, const char *name) { return static_cast<XnesProbe*>(h)->getd(name); }
int bbo_xnes_getv(void *h, const char *name, double *out, int cap) { return static_cast<XnesProbe*>(h)->getv(name, out, cap); }

void *bbo_sansde_new(int mfev, int np, double tol, int repaircr, int crref, int pupdate, int crupdate) {
    return new SansdeProbe(mfev, np, tol, repaircr != 0, crref, pupdate, crupdate);
}
void bbo_sansde_free(void *h) { delete static_cast<SansdeProbe*>(h); }

static void sansde_bind(SansdeProbe *p, int obj, int n, const double *lower, const double *upper, const double *guess) {
    p->lower.assign(lower, lower + n);
    p->upper.assign(upper, upper + n);
    p->guess.assign(guess, guess + n);
    p->prob = multivariate_problem(make_objective(obj, n), n, p->lower.data(), p->upper.data());
}

void bbo_sansde_init(void *h, int obj, int n, const double *lower, const double *upper, const double *guess) {
    SansdeProbe *p = static_cast<SansdeProbe*>(h);
    sansde_bind(p, obj, n, lower, upper, guess);
    p->init(p->prob, p->guess.data());
}

void bbo_sansde_iterate(void *h) { static_cast<SansdeProbe*>(h)->iterate(); }
/* converged() is private; solution() reports it -- and sorts the swarm first (sansde.cpp:214-217), which iterate() would do
 * at its start anyway */
int bbo_sansde_converged(void *h) { return static_cast<SansdeProbe*>(h)->solution()._converged ? 1 : 0; }

int bbo_sansde_optimize(void *h, int obj, int n, const double *lower, const double *upper, const double *guess, double *xout,
        int *fev) {
    SansdeProbe *p = static_cast<SansdeProbe*>(h);
    sansde_bind(p, obj, n, lower, upper, guess);
    multivariate_solution sol = p->optimize(p->prob, p->guess.data());
    for (int i = 0; i < n; i++) xout[i] = sol._sol[i];
    *fev = sol._fev;
    return sol._converged ? 1 : 0;
}

int bbo_sansde_geti(void *h, const char *name) { return static_cast<SansdeProbe*>(h)->geti(name); }
double bbo_sansde_getd(void *h, const char *name) { return static_cast<SansdeProbe*>(h)->getd(name); }
int bbo_sansde_getv(void *h, const char *name, double *out, int cap) { return static_cast<SansdeProbe*>(h)->getv(name, out, cap); }

void *bbo_ssde_new(int mfev, int npinit, double tol, int patience, int npmin, double ptop, int h, int usede, int repaircr) {
    return new SsdeProbe(mfev, npinit, tol, patience, npmin, ptop, h, usede != 0, repaircr != 0);
}
void bbo_ssde_free(void *h) { delete static_cast<SsdeProbe*>(h); }

static void ssde_bind(SsdeProbe *p, int obj, int n, const double *lower, const double *upper, const double *guess) {
    p->lower.assign(lower, lower + n);
    p->upper.assign(upper, upper + n);
    p->guess.assign(guess, guess + n);
    p->prob = multivariate_problem(make_objective(obj, n), n, p->lower.data(), p->upper.data());
}

void bbo_ssde_init(void *h, int obj, int n, const double *lower, const double *upper, const double *guess) {
    SsdeProbe *p = static_cast<SsdeProbe*>(h);
    ssde_bind(p, obj, n, lower, upper, guess);
    p->init(p->prob, p->guess.data());
}

void bbo_ssde_iterate(void *h) { static_cast<SsdeProbe*>(h)->iterate(); }
/* converged() is private; solution() reports it (its sort is a no-op: iterate() leaves the swarm sorted, ssde.cpp:322) */
int bbo_ssde_converged(void *h) { return static_cast<SsdeProbe*>(h)->solution()._converged ? 1 : 0; }

int bbo_ssde_optimize(void *h, int obj, int n, const double *lower, const double *upper, const double *guess, double *xout,
        int *fev) {
    SsdeProbe *p = static_cast<SsdeProbe*>(h);
    ssde_bind(p, obj, n, lower, upper, guess);
    multivariate_solution sol = p->optimize(p->prob, p->guess.data());
    for (int i = 0; i < n; i++) xout[i] = sol._sol[i];
    *fev = sol._fev;
    return sol._converged ? 1 : 0;
}

int bbo_ssde_geti(void *h, const char *name) { return static_cast<SsdeProbe*>(h)->geti(name); }
double bbo_ssde_getd(void *h, const char *name) { return static_cast<SsdeProbe*>(h)->getd(name); }
int bbo_ssde_getv(void *h, const char *name, double *out, int cap) { return static_cast<SsdeProbe*>(h)->getv(name, out, cap); }

void *bbo_de_new(int kind, int mfev, int np, double tol, int archive, int repaircr,
        double pelite, double cdamp, double sigma, int h, int npmin) {
    DeProbe *p;
    if (kind == 0) p = new JadeProbe(mfev, np, tol, archive != 0, repaircr != 0, pelite, cdamp, sigma);
    else p = new ShadeProbe(mfev, np, tol, archive != 0, repaircr != 0, h, npmin);
    return p;
}

void bbo_de_free(void *h) { delete static_cast<DeProbe*>(h); }

static void de_bind(DeProbe *p, int obj, int n, const double *lower, const double *upper,
        const double *guess) {
    p->lower.assign(lower, lower + n);
    p->upper.assign(upper, upper + n);
    p->guess.assign(guess, guess + n);
    p->prob = multivariate_problem(make_objective(obj, n), n, p->lower.data(),
            p->upper.data());
}

void bbo_de_init(void *h, int obj, int n, const double *lower, const double *upper,
        const double *guess) {
    DeProbe *p = static_cast<DeProbe*>(h);
    de_bind(p, obj, n, lower, upper, guess);
    p->opt()->init(p->prob, p->guess.data());
}

void bbo_de_iterate(void *h) { static_cast<DeProbe*>(h)->opt()->iterate(); }

int bbo_de_converged(void *h) {
    /* converged() is private in the reference; solution() reports it (jade.cpp:208-211) */
    return static_cast<DeProbe*>(h)->opt()->solution()._converged ? 1 : 0;
}

int bbo_de_optimize(void *h, int obj, int n, const double *lower, const double *upper,
        const double *guess, double *xout, int *fev) {
    DeProbe *p = static_cast<DeProbe*>(h);
    de_bind(p, obj, n, lower, upper, guess);
    multivariate_solution sol = p->opt()->optimize(p->prob, p->guess.data());
    for (int i = 0; i < n; i++) xout[i] = sol._sol[i];
    *fev = sol._fev;
    return sol._converged ? 1 : 0;
}

int bbo_de_geti(void *h, const char *name) { return static_cast<DeProbe*>(h)->geti(name); }
double bbo_de_getd(void *h, const char *name) { return static_cast<DeProbe*>(h)->getd(name); }
int bbo_de_getv(void *h, const char *name, double *out, int cap) {
    return static_cast<DeProbe*>(h)->getv(name, out, cap);
}

} // extern "C"
