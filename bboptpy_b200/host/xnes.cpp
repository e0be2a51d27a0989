// Host logic of xNES (see xnes.h); citations: /root/reference/src/multivariate/nes/xnes.cpp.
#include "xnes.h"

#include <algorithm>
#include <cmath>
#include <iostream>

using bbopt::gpu_check;

xNES::xNES(int mfev, double tol, double a0, double etamu) { // :42-48
    _tol = tol;
    _mfev = mfev;
    _a0 = a0;
    _etamu = etamu;
    _n = _np = _fev = 0;
    _etasigma = _etab = _sigma = _Gsigma = 0.;
    _stream_key = bbopt::Random::next_stream_key();
    _rep = bbgpu_cma_report();
}

xNES::~xNES() {
    if (_dev) bbgpu_cma_destroy(_dev);
}

// :50-108.  The reference starts the mean at the ORIGIN (:82) -- the guess is not used -- and B = (a0 / sigma) I = I.
void xNES::init(const multivariate_problem &f, const double *guess) {
    (void) guess;
    if (f._hasc || f._hasbbc) {
        std::cerr << "Warning [xNES]: problem constraints will be ignored." << std::endl;
    }
    _f = f;
    _n = f._n;
    _lower = std::vector<double>(f._lower, f._lower + _n);
    _upper = std::vector<double>(f._upper, f._upper + _n);
    _np = 4 + static_cast<int>(3. * std::log(_n));
    _etasigma = 3. * (3. + std::log(_n)) / (5. * _n * std::sqrt(_n));
    _etab = _etasigma;
    _u = std::vector<double>(_np);
    double sumu = 0.;
    for (int i = 1; i <= _np; i++) {
        const double ui = std::max(0., std::log(1 + 0.5 * _np) - std::log(i));
        _u[i - 1] = ui;
        sumu += ui;
    }
    for (int i = 0; i < _np; i++) _u[i] = _u[i] / sumu - (1. / _np);
    _sigma = _a0;
    _fev = 0;
    _obj.bind(_f);
    if (_dev) {
        bbgpu_cma_destroy(_dev);
        _dev = nullptr;
    }
    bbgpu_cma_config cfg = bbgpu_cma_config();
    cfg.kind = BBGPU_CMA_XNES;
    cfg.n = _n;
    cfg.lambda = _np;
    cfg.mu = _np;
    cfg.bound = 0;
    cfg.shard_count = 1;
    cfg.seed = _stream_key;
    gpu_check(bbgpu_cma_create(bbopt::device_context(), &cfg, _u.data(), nullptr, nullptr, &_dev));
    std::vector<double> origin(_n, 0.);
    gpu_check(bbgpu_cma_reset(_dev, origin.data(), 0));
}

// :110-191
void xNES::iterate() {
    if (_replay.enabled) { // np * n normals, point-major, like the reference's loop (:113-117)
        _zbuf.resize((size_t) _np * _n);
        _replay.fill_normals(_zbuf);
        gpu_check(bbgpu_cma_sample(_dev, _sigma, _gen, _zbuf.data()));
    } else {
        gpu_check(bbgpu_cma_sample(_dev, _sigma, _gen, nullptr));
    }
    _gen++;
    gpu_check(bbgpu_cma_evaluate(_dev, _obj.get()));
    _fev += _np;
    gpu_check(bbgpu_cma_rank(_dev, 0));
    gpu_check(bbgpu_cma_xnes_update(_dev, _sigma, _etamu, _etab, &_Gsigma));
    gpu_check(bbgpu_cma_fetch(_dev, &_rep));
    _sigma *= std::exp(0.5 * _etasigma * _Gsigma);
}

// :211-219: best against worst of the last generation
bool xNES::converged() {
    return std::abs(_rep.ybw[0] - _rep.ybw[3]) < _tol;
}

multivariate_solution xNES::solution() { // :193-195
    std::vector<double> x(_n, 0.);
    if (_fev > 0) gpu_check(bbgpu_cma_get_candidate(_dev, _rep.ibw[0], x.data()));
    return {x, _fev, converged()};
}

multivariate_solution xNES::optimize(const multivariate_problem &f, const double *guess) { // :197-209
    init(f, guess);
    bool converge = false;
    while (_fev < _mfev) {
        iterate();
        if (converged()) {
            converge = true;
            break;
        }
    }
    std::vector<double> x(_n, 0.);
    if (_fev > 0) gpu_check(bbgpu_cma_get_candidate(_dev, _rep.ibw[0], x.data()));
    return {x, _fev, converge};
}

std::vector<double> xNES::state(const std::string &name) {
    struct Entry { const char *key; int field; int64_t len; };
    const int64_t n = _n, np = _np;
    const Entry table[] = { { "mu", BBGPU_CMA_XMEAN, n }, { "B", BBGPU_CMA_B, n * n }, { "x", BBGPU_CMA_ARX, np * n },
            { "z", BBGPU_CMA_Z, np * n }, { "fitness", BBGPU_CMA_FITNESS, np }, { "f", BBGPU_CMA_RANKVAL, np } };
    if (name == "rank") {
        std::vector<int32_t> idx(_np);
        gpu_check(bbgpu_cma_get_rank(_dev, idx.data(), _np));
        return std::vector<double>(idx.begin(), idx.end());
    }
    if (name == "u") return _u;
    for (const Entry &e : table) {
        if (name == e.key) {
            std::vector<double> out((size_t) e.len);
            gpu_check(bbgpu_cma_get(_dev, e.field, out.data(), e.len));
            return out;
        }
    }
    throw bbopt::GpuError("unknown state field: " + name);
}
