// Host logic of SSDE (see ssde.h); citations: /root/reference/src/multivariate/de/ssde.cpp.
#include "ssde.h"

#include <algorithm>
#include <cmath>
#include <iostream>

using bbopt::gpu_check;

SSDESearch::SSDESearch(int mfev, int npinit, double tol, int patience, int npmin, double ptop, int h, bool usede, bool repaircr) { // :46-57
    _mfev = mfev;
    _npinit = npinit;
    _tol = tol;
    _patience = patience;
    _ptop = ptop;
    _npmin = npmin;
    _H = h;
    _usede = usede;
    _repaircr = repaircr;
    _n = _np = _fev = _convcount = _it = 0;
    _k = _kCR = 1;
    _bestf = 0.;
    _stream_key = bbopt::Random::next_stream_key();
}

SSDESearch::~SSDESearch() {
    if (_dev) bbgpu_ssde_destroy(_dev);
}

// :59-132: npinit uniform points (with usede also their opposites; the best npinit of the 2 npinit stay), sorted; the
// success memories start at 0.5.  The guess is not used.
void SSDESearch::init(const multivariate_problem &f, const double *guess) {
    (void) guess;
    if (f._hasc || f._hasbbc) {
        std::cerr << "Warning [SSDE]: problem constraints will be ignored." << std::endl;
    }
    _f = f;
    _n = f._n;
    _lower = std::vector<double>(f._lower, f._lower + _n);
    _upper = std::vector<double>(f._upper, f._upper + _n);
    _fev = 0;
    _it = 0;
    _convcount = 0;
    _k = _kCR = 1;
    _np = _npinit;
    _iwork.resize(_n);
    for (int i = 0; i < _n; i++) _iwork[i] = i; // shuffled in place by every generation (randA, :413)
    _obj.bind(_f);
    if (!_stream_key_fixed) _stream_key = bbopt::Random::next_stream_key();
    if (_dev) {
        bbgpu_ssde_destroy(_dev);
        _dev = nullptr;
    }
    gpu_check(bbgpu_ssde_create(bbopt::device_context(), _n, _npinit, _H, _usede ? 1 : 0, _repaircr ? 1 : 0, _stream_key, _lower.data(),
            _upper.data(), &_dev));
    const bool have = (int) _initial.size() == _npinit * _n;
    gpu_check(bbgpu_ssde_init(_dev, _obj.get(), have ? _initial.data() : nullptr, &_fev));
    std::vector<double> fit = fitness();
    _bestf = fit[0];
    _radius_fresh = false;
}

// :134-361
void SSDESearch::iterate() {
    const double oldbestf = _bestf;
    _np_trace = _np;
    bbgpu_ssde_state st = bbgpu_ssde_state();
    ssde_generation_draws d;
    if (_replay && _replay(_it, d)) {
        if ((int) d.iwork.size() != _n || (int) d.iL.size() != _np || (int) d.ub.size() != _np * _n)
            throw bbopt::GpuError("SSDE replay: wrong draw shapes");
        if (_usede && ((int) d.ucr.size() != _np * _n || (int) d.redraw.size() != _np * _n || (int) d.zcr.size() != _np
                || (int) d.j0.size() != _np)) throw bbopt::GpuError("SSDE replay: wrong second-stage draw shapes");
        _iwork = d.iwork;
        bbgpu_ssde_draws raw = bbgpu_ssde_draws();
        raw.iL = d.iL.data(); raw.zrank = d.zrank.data(); raw.ub = d.ub.data(); raw.p = d.p.data(); raw.q = d.q.data(); raw.r = d.r.data();
        raw.pbest = d.pbest.data(); raw.cauchy = d.cauchy.data();
        if (_usede) { raw.zcr = d.zcr.data(); raw.j0 = d.j0.data(); raw.ucr = d.ucr.data(); raw.redraw = d.redraw.data(); }
        gpu_check(bbgpu_ssde_generation(_dev, _obj.get(), _iwork.data(), _fev, _mfev, _ptop, _npmin, &raw, &st, &_rounds));
    } else {
        std::shuffle(_iwork.begin(), _iwork.end(), bbopt::Random::engine()); // :413
        gpu_check(bbgpu_ssde_generation(_dev, _obj.get(), _iwork.data(), _fev, _mfev, _ptop, _npmin, nullptr, &st, &_rounds));
    }
    _fev += st.evals;
    _k = st.k;
    _kCR = st.kcr;
    _nsucc = st.nsucc;
    _it++;
    _bestf = st.fbest; // the swarm is sorted
    if (_bestf < oldbestf) _convcount = 0; // :347-351
    else ++_convcount;
    _np = st.np_next; // :354-360, applied on the device with this very formula on _fev
    _rmean = st.rmean;
    _rm2 = st.rm2;
    _radius_fresh = true;
}

// :383-408: standard deviation of the members' radii, or no improvement for `patience` generations
bool SSDESearch::converged() {
    if (!_radius_fresh) { // (after a generation the statistic came back with its state)
        gpu_check(bbgpu_ssde_radius(_dev, &_rmean, &_rm2));
        _radius_fresh = true;
    }
    if (_rm2 <= (_np - 1) * _tol * _tol) return true;
    return _convcount > _patience;
}

multivariate_solution SSDESearch::solution() { // :363-366 (the swarm is kept sorted)
    std::vector<double> x = swarm();
    x.resize(_n);
    return {x, _fev, converged()};
}

multivariate_solution SSDESearch::optimize(const multivariate_problem &f, const double *guess) { // :368-381
    init(f, guess);
    bool converge = false;
    while (_fev < _mfev) {
        iterate();
        if (converged()) {
            converge = true;
            break;
        }
    }
    std::vector<double> x = swarm();
    x.resize(_n);
    return {x, _fev, converge};
}

std::vector<double> SSDESearch::swarm() {
    std::vector<double> x((size_t) _np * _n);
    gpu_check(bbgpu_ssde_get_swarm(_dev, x.data(), nullptr));
    return x;
}

std::vector<double> SSDESearch::fitness() {
    std::vector<double> f(_np);
    gpu_check(bbgpu_ssde_get_swarm(_dev, nullptr, f.data()));
    return f;
}

std::vector<double> SSDESearch::memory(int which) {
    std::vector<double> m(_H);
    gpu_check(bbgpu_ssde_get_memory(_dev, which == 0 ? m.data() : nullptr, which == 1 ? m.data() : nullptr, which == 2 ? m.data() : nullptr));
    return m;
}

void SSDESearch::trace(std::vector<int32_t> &succ, std::vector<double> &ftrial) {
    succ.assign(_np_trace, 0);
    ftrial.assign(_np_trace, 0.);
    if (_np_trace > 0) gpu_check(bbgpu_ssde_get_trace(_dev, succ.data(), ftrial.data()));
}
