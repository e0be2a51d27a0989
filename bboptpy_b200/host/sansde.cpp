// Host logic of SaNSDE (see sansde.h); citations: /root/reference/src/multivariate/de/sansde.cpp.
#include "sansde.h"

#include <cmath>
#include <iostream>

using bbopt::gpu_check;

SaNSDESearch::SaNSDESearch(int mfev, int np, double tol, bool repaircr, int crref, int pupdate, int crupdate) { // :44-52
    _mfev = mfev;
    _np = np;
    _tol = tol;
    _repaircr = repaircr;
    _ncrref = crref;
    _npup = pupdate;
    _ncrup = crupdate;
    _n = _it = _fev = 0;
    _p = _fp = _crm = 0.5;
    _acc = bbgpu_sde_accum();
    _stream_key = bbopt::Random::next_stream_key();
}

SaNSDESearch::~SaNSDESearch() {
    if (_dev) bbgpu_sde_destroy(_dev);
}

// :54-92: np uniform points in the box, np evaluations; every member starts with CR = 0.5.  The guess is not used.
void SaNSDESearch::init(const multivariate_problem &f, const double *guess) {
    (void) guess;
    if (f._hasc || f._hasbbc) {
        std::cerr << "Warning [SaNSDE]: problem constraints will be ignored." << std::endl;
    }
    _f = f;
    _n = f._n;
    _lower = std::vector<double>(f._lower, f._lower + _n);
    _upper = std::vector<double>(f._upper, f._upper + _n);
    _fev = 0;
    _it = 0;
    _p = 0.5;
    _fp = 0.5;
    _crm = 0.5;
    _acc = bbgpu_sde_accum();
    _obj.bind(_f);
    if (!_stream_key_fixed) _stream_key = bbopt::Random::next_stream_key(); // a second init() is a new run, as in the reference
    if (_dev) {
        bbgpu_sde_destroy(_dev);
        _dev = nullptr;
    }
    gpu_check(bbgpu_sde_create(bbopt::device_context(), _n, _np, _repaircr ? 1 : 0, _stream_key, _lower.data(), _upper.data(), &_dev));
    const bool have = (int) _initial.size() == _np * _n;
    gpu_check(bbgpu_sde_init(_dev, _obj.get(), have ? _initial.data() : nullptr));
    _fev = _np;
    _radius_fresh = false;
}

// :94-212
void SaNSDESearch::iterate() {
    const int refresh = _it % _ncrref == 0 ? 1 : 0;
    sansde_generation_draws d;
    if (_replay && _replay(_it, d)) {
        if ((int) d.F.size() != _np || (int) d.u.size() != _np * _n) throw bbopt::GpuError("SaNSDE replay: wrong draw shapes");
        bbgpu_sde_draws raw = bbgpu_sde_draws();
        raw.zcr = d.zcr.empty() ? nullptr : d.zcr.data();
        raw.ifstrat = d.ifstrat.data(); raw.F = d.F.data(); raw.r1 = d.r1.data(); raw.r2 = d.r2.data(); raw.r3 = d.r3.data();
        raw.istrat = d.istrat.data(); raw.jrand = d.jrand.data(); raw.u = d.u.data();
        gpu_check(bbgpu_sde_generation(_dev, _obj.get(), refresh, _crm, _fp, _p, &raw, &_acc, &_rounds));
    } else {
        gpu_check(bbgpu_sde_generation(_dev, _obj.get(), refresh, _crm, _fp, _p, nullptr, &_acc, &_rounds));
    }
    _fev += _np;
    _it++;
    _radius_fresh = true;
    if (_it > 0 && _it % _npup == 0) { // :180-187
        const int pnsf0 = _acc.pns[0] + _acc.pnf[0];
        const int pnsf1 = _acc.pns[1] + _acc.pnf[1];
        _p = (1. * _acc.pns[0] * pnsf1) / (_acc.pns[1] * pnsf0 + _acc.pns[0] * pnsf1);
        _acc.pns[0] = _acc.pns[1] = _acc.pnf[0] = _acc.pnf[1] = 0;
    }
    if (_it > 0 && _it % _ncrup == 0) { // :190-195
        if (_acc.crdeltaf > 0) _crm = _acc.crrec / _acc.crdeltaf;
        _acc.crrec = _acc.crdeltaf = 0.;
    }
    if (_it > 0 && _it % _ncrup == 0) { // :198-205
        const double fpnsf0 = _acc.fpns[0] + _acc.fpnf[0];
        const double fpnsf1 = _acc.fpns[1] + _acc.fpnf[1];
        _fp = (1. * _acc.fpns[0] * fpnsf1) / (_acc.fpns[1] * fpnsf0 + _acc.fpns[0] * fpnsf1);
        _acc.fpns[0] = _acc.fpns[1] = _acc.fpnf[0] = _acc.fpnf[1] = 0.;
    }
}

// :235-253: standard deviation of the members' radii
bool SaNSDESearch::converged() {
    if (!_radius_fresh) { // (after a generation the statistic came back with the sums)
        gpu_check(bbgpu_sde_radius(_dev, &_acc.rmean, &_acc.rm2));
        _radius_fresh = true;
    }
    return _acc.rm2 <= (_np - 1) * _tol * _tol;
}

multivariate_solution SaNSDESearch::solution() { // :214-217: sorts, then the best member
    gpu_check(bbgpu_sde_sort(_dev));
    std::vector<double> x = swarm();
    x.resize(_n);
    return {x, _fev, converged()};
}

multivariate_solution SaNSDESearch::optimize(const multivariate_problem &f, const double *guess) { // :219-233
    init(f, guess);
    bool converge = false;
    while (_fev < _mfev) {
        iterate();
        if (converged()) {
            converge = true;
            break;
        }
    }
    gpu_check(bbgpu_sde_sort(_dev));
    std::vector<double> x = swarm();
    x.resize(_n);
    return {x, _fev, converge};
}

std::vector<double> SaNSDESearch::swarm() {
    std::vector<double> x((size_t) _np * _n);
    gpu_check(bbgpu_sde_get_swarm(_dev, x.data(), nullptr, nullptr));
    return x;
}

std::vector<double> SaNSDESearch::fitness() {
    std::vector<double> f(_np);
    gpu_check(bbgpu_sde_get_swarm(_dev, nullptr, f.data(), nullptr));
    return f;
}

std::vector<double> SaNSDESearch::crossoverRates() {
    std::vector<double> c(_np);
    gpu_check(bbgpu_sde_get_swarm(_dev, nullptr, nullptr, c.data()));
    return c;
}
