#include "de.h"

#include <cstring>
#include <iostream>

using bbopt::gpu_check;

DeSearchBase::DeSearchBase() {
    _archive = _repaircr = true;
    _n = _np = _npinit = _fev = _mfev = _gen = 0;
    _h = 1;
    _npmin = 4;
    _pelite = 0.05;
    _c = 0.1;
    _tol = 0.;
    _sigma = 0.07;
    _stream_key = bbopt::Random::next_stream_key();
    std::memset(&_rep, 0, sizeof(_rep));
}

DeSearchBase::~DeSearchBase() {
    if (_dev) bbgpu_de_destroy(_dev);
}

JadeSearch::JadeSearch(int mfev, int np, double tol, bool archive, bool repaircr, double pelite, double cdamp,
        double sigma) {
    _mfev = mfev;
    _npinit = _np = np;
    _tol = tol;
    _archive = archive;
    _repaircr = repaircr;
    _pelite = pelite;
    _c = cdamp;
    _sigma = sigma;
}

ShadeSearch::ShadeSearch(int mfev, int npinit, double tol, bool archive, bool repaircr, int h, int npmin) {
    _mfev = mfev;
    _tol = tol;
    _archive = archive;
    _repaircr = repaircr;
    _h = h;
    _npinit = _np = npinit;
    _npmin = npmin;
}

// de/jade.cpp:64-96, de/shade.cpp:56-94: uniform initial points, np evaluations
void DeSearchBase::init(const multivariate_problem &f, const double *guess) {
    (void) guess; // the reference ignores the guess as well
    if (f._hasc || f._hasbbc) {
        std::cerr << "Warning [" << label() << "]: problem constraints will be ignored." << std::endl;
    }
    _f = f;
    _n = f._n;
    _lower = std::vector<double>(f._lower, f._lower + _n);
    _upper = std::vector<double>(f._upper, f._upper + _n);
    _obj.bind(_f);
    if (_dev) {
        bbgpu_de_destroy(_dev);
        _dev = nullptr;
    }
    bbgpu_de_config cfg;
    std::memset(&cfg, 0, sizeof(cfg));
    cfg.kind = kind();
    cfg.n = _n;
    cfg.np = _npinit;
    cfg.archive = _archive ? 1 : 0;
    cfg.repaircr = _repaircr ? 1 : 0;
    cfg.h = _h;
    cfg.npmin = _npmin;
    cfg.mfev = _mfev;
    cfg.pelite = _pelite;
    cfg.cdamp = _c;
    cfg.sigma = _sigma;
    // a fresh Philox key per init(): a second optimize()/initialize() on the same object is an independent run, as in the
    // reference, whose process-global engine simply keeps advancing (src/random.hpp:677-680).  A key given explicitly
    // through setStreamKey() is kept (reproducible runs).
    if (!_stream_key_fixed) _stream_key = bbopt::Random::next_stream_key();
    cfg.seed = _stream_key;
    gpu_check(bbgpu_de_create(bbopt::device_context(), &cfg, _lower.data(), _upper.data(), &_dev));
    const bool replayInit = (int) _initial.size() == _npinit * _n;
    gpu_check(bbgpu_de_init(_dev, _obj.get(), replayInit ? _initial.data() : nullptr));
    _np = _npinit;
    _fev = _npinit;
    _gen = 0;
    std::memset(&_rep, 0, sizeof(_rep));
    _rep.np = _np;
}

// de/jade.cpp:98-206, de/shade.cpp:96-236: one generation on the device
void DeSearchBase::iterate() {
    de_generation_draws dr;
    if (_replay && _replay(_gen, dr)) {
        bbgpu_de_draws d;
        std::memset(&d, 0, sizeof(d));
        d.zcr = dr.zcr.data(); d.cauchy = dr.cauchy.data(); d.ibest = dr.ibest.data(); d.r1 = dr.r1.data();
        d.r2 = dr.r2.data(); d.jrand = dr.jrand.data(); d.slot = dr.slot.data(); d.u = dr.u.data();
        d.ri = dr.ri.empty() ? nullptr : dr.ri.data();
        d.pi = dr.pi.empty() ? nullptr : dr.pi.data();
        d.trim = dr.trim.empty() ? nullptr : dr.trim.data();
        d.ntrim = (int) dr.trim.size();
        gpu_check(bbgpu_de_generation(_dev, _obj.get(), &d, &_rep));
    } else {
        gpu_check(bbgpu_de_generation(_dev, _obj.get(), nullptr, &_rep));
    }
    _np = _rep.np;
    _fev = _rep.fev;
    _gen++;
}

// de/jade.cpp:228-245: variance of the member radii against (np-1) tol^2
bool DeSearchBase::converged() {
    if (_gen == 0) return false;
    return _rep.radius_m2 <= (_np - 1) * _tol * _tol;
}

std::vector<double> DeSearchBase::swarm() {
    std::vector<double> x((size_t) _np * _n);
    gpu_check(bbgpu_de_get_swarm(_dev, x.data(), nullptr));
    return x;
}

std::vector<double> DeSearchBase::fitness() {
    std::vector<double> f((size_t) _np);
    gpu_check(bbgpu_de_get_swarm(_dev, nullptr, f.data()));
    return f;
}

// de/jade.cpp:208-211: best member (the reference sorts the swarm here; we pick the minimum)
multivariate_solution DeSearchBase::solution() {
    std::vector<double> x = swarm(), f = fitness();
    int best = 0;
    for (int i = 1; i < _np; i++) if (f[i] < f[best]) best = i;
    return {std::vector<double>(x.begin() + (size_t) best * _n, x.begin() + (size_t) (best + 1) * _n), _fev, converged()};
}

// de/jade.cpp:213-226
multivariate_solution DeSearchBase::optimize(const multivariate_problem &f, const double *guess) {
    init(f, guess);
    bool converge = false;
    while (_fev < _mfev) {
        iterate();
        if (converged()) {
            converge = true;
            break;
        }
    }
    std::vector<double> x = swarm(), fit = fitness();
    int best = 0;
    for (int i = 1; i < _np; i++) if (fit[i] < fit[best]) best = i;
    return {std::vector<double>(x.begin() + (size_t) best * _n, x.begin() + (size_t) (best + 1) * _n), _fev, converge};
}
