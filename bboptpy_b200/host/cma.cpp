// Host logic of the CMA-ES family (see cma.h).  Strategy constants and control flow follow the
// reference line by line (citations inline); every O(lambda n^2) / O(n^3) step is a C-ABI call.
#include "cma.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <iostream>
#include <limits>

using bbopt::gpu_check;

void cmaes_history::add(double value) {
    _buffer = (_buffer + 1) % _cap;
    _hist[_buffer] = value;
    if (_len < _cap) _len++;
}

double cmaes_history::get(int i) {
    return _hist[(_cap + _buffer - i) % _cap];
}

BaseCmaes::BaseCmaes(int mfev, double tol, int np, double sigma0, bool bound) {
    _tol = tol;
    _lambda = np;
    _sigma0 = sigma0;
    _mfev = mfev;
    _adaptpop = _adaptit = false;
    _bound = bound;
    _n = _mit = _mu = _it = _hlen = _ik = _fev = 0;
    _mueff = _cc = _cs = _c1 = _cmu = _damps = _chi = _sigma = _fbest = _fworst = 0.;
    _stream_key = bbopt::Random::next_stream_key();
    _rep = bbgpu_cma_report();
}

BaseCmaes::~BaseCmaes() {
    if (_dev) bbgpu_cma_destroy(_dev);
}

// host half of BaseCmaes::init (cma/base_cmaes.cpp:54-134)
void BaseCmaes::init(const multivariate_problem &f, const double *guess) {
    if (f._hasc || f._hasbbc) {
        std::cerr << "Warning [CMA]: problem constraints will be ignored." << std::endl;
    }
    _f = f;
    _n = f._n;
    _lower = std::vector<double>(f._lower, f._lower + _n);
    _upper = std::vector<double>(f._upper, f._upper + _n);
    if (_adaptpop) _lambda = 4 + static_cast<int>(3. * std::log(_n));
    _mu = _lambda / 2;
    if (_adaptit) {
        _mit = static_cast<int>(100 + 50 * (_n + 3) * (_n + 3) / std::sqrt(1. * _lambda));
        _mfev = _mit * _lambda;
    } else {
        _mit = _mfev / _lambda;
    }
    _ibw.assign(4, 0);
    _ybw.assign(4, 0.);
    // log-linear recombination weights, normalised (:94-100)
    _weights.assign(_mu, 0.);
    double sum = 0.;
    for (int i = 0; i < _mu; i++) {
        _weights[i] = std::log(0.5 * (_lambda + 1.)) - std::log(i + 1.);
        sum += _weights[i];
    }
    const double inv = 1. / sum;
    for (int i = 0; i < _mu; i++) _weights[i] *= inv;
    double lenw = 0.;
    for (int i = 0; i < _mu; i++) lenw = lenw + _weights[i] * _weights[i];
    _mueff = 1. / lenw;
    // strategy parameters (:107-117)
    _chi = std::sqrt(_n) * (1. - 1. / (4. * _n) + 1. / (21. * _n * _n));
    _sigma = _sigma0;
    _cc = (4. + _mueff / _n) / (_n + 4. + 2. * _mueff / _n);
    _cs = (_mueff + 2.) / (5. + _n + _mueff);
    _c1 = 2. / ((1.3 + _n) * (1.3 + _n) + _mueff);
    _cmu = std::min(1. - _c1, 2. * (_mueff - 2. + 1. / _mueff) / ((2. + _n) * (2. + _n) + _mueff));
    _damps = 1. + _cs + 2. * std::max(0., std::sqrt((_mueff - 1.) / (_n + 1.)) - 1.);
    _it = _fev = 0;
    _hlen = 10 + static_cast<int>(std::ceil((30. * _n) / _lambda));
    _ik = static_cast<int>(std::ceil(0.1 + _lambda / 4.));
    _best = cmaes_history { _hlen, -1, 0, std::vector<double>(_hlen, 0.) };
    _kth = cmaes_history { _hlen, -1, 0, std::vector<double>(_hlen, 0.) };
    _fbest = -std::numeric_limits<double>::infinity();
    _fworst = std::numeric_limits<double>::infinity();
    _fkth = 0.;
    _rep = bbgpu_cma_report();
    _rep_fresh = false;
    _zbuf.clear();
    // the guess is uploaded by ensureDevice(), called at the end of the most-derived init()
    _ybw[0] = 0.;
    _xguess.assign(guess, guess + _n);
    _obj.bind(_f);
}

void BaseCmaes::fillConfig(bbgpu_cma_config &cfg) const {
    cfg.kind = deviceKind();
    cfg.n = _n;
    cfg.lambda = _lambda;
    cfg.mu = _mu;
    cfg.bound = _bound ? 1 : 0;
    cfg.shard_rank = _comm ? bbgpu_comm_rank(_comm) : 0;
    cfg.shard_count = _comm ? bbgpu_comm_world(_comm) : 1;
    cfg.seed = _stream_key;
    cfg.cc = _cc; cfg.cs = _cs; cfg.c1 = _c1; cfg.cmu = _cmu; cfg.mueff = _mueff; cfg.chi = _chi;
    cfg.cneg = 0.; cfg.alphaold = 0.; cfg.cm = 1.; cfg.ccov = 0.;
}

// (re)create the device state for the current (kind, n, lambda, constants) and reset it to the
// reference's init state.  The reference re-inits the SAME object on every IPOP/BIPOP restart and
// its `_b` / `_c` keep their off-diagonal entries (cmaes.cpp:53-60 `resize` without clear; only
// the diagonals are set back to 1, and `_invsqrtc` is cleared): carried over here.
void BaseCmaes::ensureDevice() {
    bbgpu_ctx *ctx = bbopt::device_context();
    const bool full = deviceKind() == BBGPU_CMA_FULL || deviceKind() == BBGPU_CMA_ACTIVE;
    std::vector<double> oldC, oldB;
    const bool carry = full && _dev != nullptr && _dev_used && _dev_n == _n &&
            (_dev_kind == BBGPU_CMA_FULL || _dev_kind == BBGPU_CMA_ACTIVE);
    if (carry) {
        oldC.resize((size_t) _n * _n);
        oldB.resize((size_t) _n * _n);
        gpu_check(bbgpu_cma_get(_dev, BBGPU_CMA_C, oldC.data(), (int64_t) oldC.size()));
        gpu_check(bbgpu_cma_get(_dev, BBGPU_CMA_B, oldB.data(), (int64_t) oldB.size()));
    }
    if (_dev) {
        bbgpu_cma_destroy(_dev);
        _dev = nullptr;
    }
    bbgpu_cma_config cfg;
    fillConfig(cfg);
    gpu_check(bbgpu_cma_create(ctx, &cfg, _weights.data(), _lower.data(), _upper.data(), &_dev));
    _dev_kind = cfg.kind;
    _dev_n = _n;
    _dev_lambda = _lambda;
    if (carry) {
        gpu_check(bbgpu_cma_set(_dev, BBGPU_CMA_C, oldC.data(), (int64_t) oldC.size()));
        gpu_check(bbgpu_cma_set(_dev, BBGPU_CMA_B, oldB.data(), (int64_t) oldB.size()));
    }
    gpu_check(bbgpu_cma_reset(_dev, _xguess.data(), carry ? 1 : 0));
    if (_comm && bbgpu_comm_world(_comm) > 1) gpu_check(bbgpu_cma_attach_comm(_dev, _comm));
    _dev_used = false;
}

void BaseCmaes::setParams(int np, double sigma, int mfev) {
    _lambda = np;
    _sigma0 = sigma;
    _mfev = mfev;
    _adaptpop = _adaptit = false;
    if (_bound) {
        _bound = false;
        std::cerr << "Warning [CMA]: box bounding is no longer enabled." << std::endl;
    }
}

// cma/base_cmaes.cpp:150-156
void BaseCmaes::iterate() {
    samplePopulation();
    evaluateAndSortPopulation();
    updateDistribution();
    updateHistory();
    _it++;
}

multivariate_solution BaseCmaes::solution() {
    return {bestSolution(), _fev, converged()};
}

// cma/base_cmaes.cpp:162-174
multivariate_solution BaseCmaes::optimize(const multivariate_problem &f, const double *guess) {
    init(f, guess);
    bool converge = false;
    while (_fev < _mfev) {
        iterate();
        if (converged()) {
            converge = true;
            break;
        }
    }
    return {bestSolution(), _fev, converge};
}

// sample(): cmaes.cpp:65-80 / sep_cmaes.cpp:69-79 on the device.  Draws: the device Philox
// stream keyed by (instance key, generation counter), or -- in replay mode -- lambda*n host
// normals in the reference's candidate-major draw order.
void BaseCmaes::samplePopulation() {
    _dev_used = true;
    if (_replay.enabled) {
        _zbuf.resize((size_t) _lambda * _n);
        _replay.fill_normals(_zbuf);
        gpu_check(bbgpu_cma_sample(_dev, _sigma, _gen, _zbuf.data()));
    } else {
        gpu_check(bbgpu_cma_sample(_dev, _sigma, _gen, nullptr));
    }
    _gen++;
}

// cma/base_cmaes.cpp:211-230: evaluation and ranking stay on the device; ibw/ybw come back
// with the per-generation report
void BaseCmaes::evaluateAndSortPopulation() {
    gpu_check(bbgpu_cma_evaluate(_dev, _obj.get()));
    _fev += _lambda;
    gpu_check(bbgpu_cma_rank(_dev, _ik));
    _rep_fresh = false;
}

void BaseCmaes::fetchReport() {
    gpu_check(bbgpu_cma_fetch(_dev, &_rep));
    for (int q = 0; q < 4; q++) {
        _ibw[q] = _rep.ibw[q];
        _ybw[q] = _rep.ybw[q];
    }
    _fkth = _rep.f_kth;
    _rep_fresh = true;
}

// cma/base_cmaes.cpp:176-189 with ||p_sigma|| computed on the device
void BaseCmaes::updateSigma() {
    const double pslen = _rep.pslen;
    _sigma *= std::exp(std::min(1., (_cs / _damps) * (pslen / _chi - 1.)));
    if (_ybw[0] == _fkth) _sigma *= std::exp(0.2 + _cs / _damps);
    if (_it >= _hlen && _fworst - _fbest == 0.) _sigma *= std::exp(0.2 + _cs / _damps);
}

// cma/base_cmaes.cpp:191-209
void BaseCmaes::updateHistory() {
    if (_it >= _mit) return;
    if (!_rep_fresh) fetchReport();
    _best.add(_ybw[0]);
    _kth.add(_fkth);
    if (_best._len == _best._cap) {
        _fbest = std::numeric_limits<double>::infinity();
        _fworst = -std::numeric_limits<double>::infinity();
        for (double fx : _best._hist) {
            _fbest = std::min(fx, _fbest);
            _fworst = std::max(fx, _fworst);
        }
    }
}

// cma/base_cmaes.cpp:232-238
std::vector<double> BaseCmaes::bestSolution() {
    std::vector<double> x(_n);
    if (_it <= 0) {
        gpu_check(bbgpu_cma_get(_dev, BBGPU_CMA_XMEAN, x.data(), _n));
    } else {
        if (!_rep_fresh) fetchReport();
        gpu_check(bbgpu_cma_get_candidate(_dev, _ibw[0], x.data())); // collective on a sharded run
    }
    return x;
}

std::vector<double> BaseCmaes::state(const std::string &name) {
    struct Entry { const char *key; int field; int64_t len; };
    const bool full = deviceKind() != BBGPU_CMA_SEP;
    const int64_t n = _n, lam = _lambda;
    const Entry table[] = { { "xmean", BBGPU_CMA_XMEAN, n }, { "xold", BBGPU_CMA_XOLD, n }, { "pc", BBGPU_CMA_PC, n },
            { "ps", BBGPU_CMA_PS, n }, { "D", BBGPU_CMA_D, n }, { "C", BBGPU_CMA_C, full ? n * n : n },
            { "B", BBGPU_CMA_B, n * n }, { "invsqrtC", BBGPU_CMA_INVSQRTC, n * n }, { "arx", BBGPU_CMA_ARX, lam * n },
            { "z", BBGPU_CMA_Z, (deviceKind() == BBGPU_CMA_LM ? (lam + 1) / 2 : lam) * n }, { "fitness", BBGPU_CMA_FITNESS, lam }, { "rankval", BBGPU_CMA_RANKVAL, lam },
            { "ycoeff", BBGPU_CMA_YCOEFF, (int64_t) _mu }, { "weights", BBGPU_CMA_WEIGHTS, (int64_t) _mu },
            { "A", BBGPU_CMA_A, n * n }, { "pcmat", BBGPU_CMA_PCMAT, memoryRows() * n }, { "vmat", BBGPU_CMA_VMAT, memoryRows() * n },
            { "b", BBGPU_CMA_LMB, (int64_t) memoryRows() }, { "d", BBGPU_CMA_LMD, (int64_t) memoryRows() } };
    if (name == "rank") {
        std::vector<int32_t> idx(_lambda);
        gpu_check(bbgpu_cma_get_rank(_dev, idx.data(), _lambda));
        return std::vector<double>(idx.begin(), idx.end());
    }
    for (const Entry &e : table) {
        if (name == e.key) {
            std::vector<double> out((size_t) e.len);
            gpu_check(bbgpu_cma_get(_dev, e.field, out.data(), e.len));
            return out;
        }
    }
    throw bbopt::GpuError("unknown state field: " + name);
}

// ---------------------------------------------------------------------------------------
// checkpoint / resume
// ---------------------------------------------------------------------------------------
// One walk over the fields serves both directions: a StateIO either appends to a blob or reads from one.
struct BaseCmaes::StateIO {
    std::string *out = nullptr;      // writing
    const std::string *in = nullptr; // reading
    size_t pos = 0;
    bool writing() const { return out != nullptr; }
    void raw(void *p, size_t bytes) {
        if (out) out->append(static_cast<const char*>(p), bytes);
        else {
            if (pos + bytes > in->size()) throw bbopt::GpuError("checkpoint: truncated blob");
            std::memcpy(p, in->data() + pos, bytes);
            pos += bytes;
        }
    }
    template<typename T> void pod(T &v) { raw(&v, sizeof(T)); }
    void flag(bool &b) { uint8_t v = b ? 1 : 0; pod(v); b = v != 0; }
    template<typename T> void vec(std::vector<T> &v) {
        uint64_t len = v.size();
        pod(len);
        if (!out) {
            if (len > (in->size() - pos) / sizeof(T)) throw bbopt::GpuError("checkpoint: corrupt vector length");
            v.resize((size_t) len);
        }
        if (len) raw(v.data(), (size_t) len * sizeof(T));
    }
    void ring(cmaes_history &h) { pod(h._cap); pod(h._buffer); pod(h._len); vec(h._hist); }
};

void BaseCmaes::extraState(StateIO &) {}

namespace {
const char CK_MAGIC[8] = { 'B', 'B', 'O', 'P', 'T', 'C', 'K', '1' };
const uint32_t CK_VERSION = 1;
// device fields per kind, in blob order
std::vector<int> checkpoint_fields(int kind) {
    std::vector<int> f = { BBGPU_CMA_XMEAN, BBGPU_CMA_XOLD, BBGPU_CMA_PC, BBGPU_CMA_PS, BBGPU_CMA_ARX, BBGPU_CMA_FITNESS };
    if (kind == BBGPU_CMA_FULL || kind == BBGPU_CMA_ACTIVE) {
        // D before B: the engine rebuilds its scaled copies of the basis when either arrives
        for (int x : { BBGPU_CMA_C, BBGPU_CMA_D, BBGPU_CMA_B, BBGPU_CMA_INVSQRTC }) f.push_back(x);
    } else if (kind == BBGPU_CMA_SEP) {
        f.push_back(BBGPU_CMA_C);
        f.push_back(BBGPU_CMA_D);
    } else if (kind == BBGPU_CMA_LM) {
        for (int x : { BBGPU_CMA_PCMAT, BBGPU_CMA_VMAT, BBGPU_CMA_LMB, BBGPU_CMA_LMD }) f.push_back(x);
    } else {
        f.push_back(BBGPU_CMA_A);
    }
    return f;
}
int64_t field_length(int field, int kind, int64_t n, int64_t lam, int64_t mem = 0) {
    switch (field) {
    case BBGPU_CMA_PCMAT: case BBGPU_CMA_VMAT: return mem * n;
    case BBGPU_CMA_LMB: case BBGPU_CMA_LMD: return mem;
    case BBGPU_CMA_ARX: return lam * n;
    case BBGPU_CMA_FITNESS: return lam;
    case BBGPU_CMA_C: return kind == BBGPU_CMA_SEP ? n : n * n;
    case BBGPU_CMA_B: case BBGPU_CMA_INVSQRTC: case BBGPU_CMA_A: return n * n;
    default: return n;
    }
}
}

std::string BaseCmaes::checkpoint() {
    if (!_dev) throw bbopt::GpuError("checkpoint: the optimizer has not been initialised");
    if (_comm && bbgpu_comm_world(_comm) > 1) throw bbopt::GpuError("checkpoint: not available for a candidate-sharded run");
    std::string blob(CK_MAGIC, CK_MAGIC + 8);
    StateIO io;
    io.out = &blob;
    uint32_t version = CK_VERSION;
    int32_t kind = deviceKind(), n = _n, lambda = _lambda, mu = _mu;
    io.pod(version); io.pod(kind); io.pod(n); io.pod(lambda); io.pod(mu);
    restoreOrSaveScalars(io);
    for (int f : checkpoint_fields(kind)) {
        std::vector<double> v((size_t) field_length(f, kind, _n, _lambda, memoryRows()));
        gpu_check(bbgpu_cma_get(_dev, f, v.data(), (int64_t) v.size()));
        io.vec(v);
    }
    return blob;
}

void BaseCmaes::restore(const std::string &blob) {
    if (!_dev) throw bbopt::GpuError("restore: initialize(f, lower, upper, guess) the optimizer first");
    if (blob.size() < 8 || std::memcmp(blob.data(), CK_MAGIC, 8) != 0) throw bbopt::GpuError("restore: not a bboptpy_b200 checkpoint");
    StateIO io;
    io.in = &blob;
    io.pos = 8;
    uint32_t version = 0;
    int32_t kind = 0, n = 0, lambda = 0, mu = 0;
    io.pod(version); io.pod(kind); io.pod(n); io.pod(lambda); io.pod(mu);
    if (version != CK_VERSION) throw bbopt::GpuError("restore: unsupported checkpoint version " + std::to_string(version));
    if (kind != deviceKind()) throw bbopt::GpuError("restore: the checkpoint is of another optimizer class");
    if (n != _n) throw bbopt::GpuError("restore: the checkpoint has dimension " + std::to_string(n));
    restoreOrSaveScalars(io);
    if (_lambda != lambda || _mu != mu) throw bbopt::GpuError("restore: inconsistent checkpoint header");
    // the device handle is keyed by the Philox stream key and sized by lambda (which differ from this object's own, e.g. a
    // run checkpointed inside an IPOP restart): re-create it from the restored scalars
    _dev_used = false;
    ensureDevice();
    for (int f : checkpoint_fields(kind)) {
        std::vector<double> v;
        io.vec(v);
        if ((int64_t) v.size() != field_length(f, kind, _n, _lambda, memoryRows())) throw bbopt::GpuError("restore: field of the wrong length");
        gpu_check(bbgpu_cma_set(_dev, f, v.data(), (int64_t) v.size()));
    }
    if (io.pos != blob.size()) throw bbopt::GpuError("restore: trailing bytes in the checkpoint");
    _dev_used = true;
    _rep_fresh = true; // _ibw / _ybw / _fkth came with the blob
}

void BaseCmaes::restoreOrSaveScalars(StateIO &io) {
    io.flag(_adaptpop); io.flag(_adaptit); io.flag(_bound);
    io.pod(_mfev); io.pod(_mit); io.pod(_lambda); io.pod(_mu); io.pod(_it); io.pod(_hlen); io.pod(_ik); io.pod(_fev);
    io.pod(_tol); io.pod(_sigma0); io.pod(_mueff); io.pod(_cc); io.pod(_cs); io.pod(_c1); io.pod(_cmu); io.pod(_damps);
    io.pod(_chi); io.pod(_sigma); io.pod(_fbest); io.pod(_fworst); io.pod(_fkth);
    io.ring(_best); io.ring(_kth);
    io.vec(_ibw); io.vec(_ybw); io.vec(_weights);
    io.pod(_stream_key); io.pod(_gen);
    io.pod(_rep);
    extraState(io);
}

// ---------------------------------------------------------------------------------------
// Cmaes
// ---------------------------------------------------------------------------------------

// cma/cmaes.cpp:44-63
void Cmaes::init(const multivariate_problem &f, const double *guess) {
    BaseCmaes::init(f, guess);
    _eigenfreq = _eigenrate * _lambda / (_c1 + _cmu) / _n;
    _eigenlastev = 0;
    _flag = 0;
    if (deviceKind() == BBGPU_CMA_FULL) ensureDevice();
}

// cma/cmaes.cpp:82-149 (the same flow serves ActiveCmaes, active_cmaes.cpp:71-168: the
// variant-specific arithmetic is selected by the device handle's kind)
void Cmaes::updateDistribution() {
    gpu_check(bbgpu_cma_update(_dev, _sigma, _fev));
    fetchReport();
    updateSigma();
    eigenDecomposition();
}

// cma/cmaes.cpp:229-283: the gate stays on the host, the decomposition runs on the device
void Cmaes::eigenDecomposition() {
    if (_fev - _eigenlastev <= _eigenfreq) return;
    _eigenlastev = _fev;
    _eigcount++;
    if (_eigreplay) {
        std::vector<double> B, D;
        if (_eigreplay(_fev, _n, B, D)) {
            if ((int) B.size() != _n * _n || (int) D.size() != _n) throw bbopt::GpuError("eigen replay: wrong basis shape");
            gpu_check(bbgpu_cma_eig(_dev, B.data(), D.data()));
            return;
        }
    }
    gpu_check(bbgpu_cma_eig(_dev, nullptr, nullptr));
}

// cma/cmaes.cpp:151-227
bool Cmaes::converged() {
    if (_it >= _mit) { _flag = 1; return true; }                                  // MaxIter
    if (_it >= _hlen && _fworst - _fbest < _tol) { _flag = 2; return true; }      // TolHistFun
    if (_best._len >= _n && _kth._len >= _n) {                                    // EqualFunVals
        int countEq = 0;
        for (int i = 0; i < _n; i++) {
            if (_best.get(i) == _kth.get(i)) {
                countEq++;
                if (3 * countEq >= _n) { _flag = 3; return true; }
            }
        }
    }
    int iaxis = _n - 1 - ((_it - 1) % _n);
    iaxis = std::max(0, std::min(iaxis, _n - 1)); // _it == 0 indexes out of range in the reference
    gpu_check(bbgpu_cma_probe(_dev, _sigma, iaxis));
    gpu_check(bbgpu_cma_fetch(_dev, &_rep));
    // TolX: all_i max(pc_i, sqrt(C_ii)) sigma/sigma0 < tol  <=>  the max satisfies it
    if (!(_rep.tolx_max * _sigma / _sigma0 >= _tol)) { _flag = 4; return true; }
    if (_sigma / _sigma0 > 1.0e20 * _rep.d_max) { _flag = 5; return true; }       // TolUpSigma
    if (_rep.d_max > 1.0e7 * _rep.d_min) { _flag = 7; return true; }              // ConditionCov
    if (_rep.noeffect_axis) { _flag = 8; return true; }                           // NoEffectAxis
    if (_rep.noeffect_coor) { _flag = 9; return true; }                           // NoEffectCoor
    return false;
}

// ---------------------------------------------------------------------------------------
// ActiveCmaes -- cma/active_cmaes.cpp:42-69
// ---------------------------------------------------------------------------------------
void ActiveCmaes::init(const multivariate_problem &f, const double *guess) {
    Cmaes::init(f, guess);
    _cm = 1.;
    _alphaold = 0.5;
    _cc = (4. + 0. * _mueff / _n) / (_n + 4. + 0. * 2. * _mueff / _n);
    _cs = (_mueff + 2.) / (3. + _n + _mueff);
    _c1 = _alphacov * std::min(1., _lambda / 6.) / ((_n + 1.3) * (_n + 1.3) + _mueff);
    _cmu = 1. - _c1;
    _cmu = std::min(_cmu, _alphacov * (_mueff - 2. + 1. / _mueff) / ((2. + _n) * (2. + _n) + _alphacov * _mueff / 2.));
    _cneg = (1. - _cmu) * (_alphacov / 8.) * _mueff / (std::pow(_n + 2., 1.5) + 2. * _mueff);
    _damps = 1. + _cs + 2. * std::max(0., std::sqrt((_mueff - 1.) / (_n + 1.)) - 1.);
    _eigenfreq = _eigenrate * (1. / (_c1 + _cmu + _cneg)) / _n;
    _eigenlastev = 0;
    ensureDevice();
}

void ActiveCmaes::fillConfig(bbgpu_cma_config &cfg) const {
    BaseCmaes::fillConfig(cfg);
    cfg.cneg = _cneg;
    cfg.alphaold = _alphaold;
    cfg.cm = _cm;
}

// ---------------------------------------------------------------------------------------
// SepCmaes -- cma/sep_cmaes.cpp:40-67, 81-135, 137-207
// ---------------------------------------------------------------------------------------
void SepCmaes::init(const multivariate_problem &f, const double *guess) {
    BaseCmaes::init(f, guess);
    _cc = 4. / (_n + 4.);
    _cs = (_mueff + 2.) / (3. + _n + _mueff);
    _damps = 1. + _cs + 2. * std::max(0., std::sqrt((_mueff - 1.) / (_n + 1.)) - 1.);
    _ccov = 2. / ((_n + std::sqrt(2.)) * (_n + std::sqrt(2.)) * _mueff);
    _ccov += std::min(1., (2. * _mueff - 1.) / ((_n + 2.) * (_n + 2.) + _mueff)) * (1. - 1. / _mueff);
    if (_adjustlr) _ccov *= ((_n + 2.) / 3.);
    _flag = 0;
    ensureDevice();
}

void SepCmaes::fillConfig(bbgpu_cma_config &cfg) const {
    BaseCmaes::fillConfig(cfg);
    cfg.ccov = _ccov;
}

void SepCmaes::updateDistribution() {
    gpu_check(bbgpu_cma_update(_dev, _sigma, _fev));
    fetchReport();
    updateSigma();
}

bool SepCmaes::converged() {
    if (_it >= _mit) { _flag = 1; return true; }
    if (_it >= _hlen && _fworst - _fbest < _tol) { _flag = 2; return true; }
    if (_best._len >= _n && _kth._len >= _n) {
        int countEq = 0;
        for (int i = 0; i < _n; i++) {
            if (_best.get(i) == _kth.get(i)) {
                countEq++;
                if (3 * countEq >= _n) { _flag = 3; return true; }
            }
        }
    }
    int iaxis = _n - 1 - ((_it - 1) % _n);
    iaxis = std::max(0, std::min(iaxis, _n - 1));
    gpu_check(bbgpu_cma_probe(_dev, _sigma, iaxis));
    gpu_check(bbgpu_cma_fetch(_dev, &_rep));
    if (!(_rep.tolx_max * _sigma / _sigma0 >= _tol)) { _flag = 4; return true; }
    // the reference indexes the UNSORTED ends _diagd[n-1] / _diagd[0] here (sep_cmaes.cpp:181,187)
    if (_sigma / _sigma0 > 1.0e20 * _rep.d_max) { _flag = 5; return true; }
    if (_rep.d_max > 1.0e7 * _rep.d_min) { _flag = 7; return true; }
    if (_rep.noeffect_axis) { _flag = 8; return true; }
    if (_rep.noeffect_coor) { _flag = 9; return true; }
    return false;
}

// ---------------------------------------------------------------------------------------
// CholeskyCmaes -- cma/cholesky_cmaes.cpp:40-52, 54-118, 137-161
// ---------------------------------------------------------------------------------------
void CholeskyCmaes::init(const multivariate_problem &f, const double *guess) {
    BaseCmaes::init(f, guess);
    ensureDevice(); // A = I: the reference clears `_a` on every init (:45-51), nothing is carried over a restart
}

void CholeskyCmaes::updateDistribution() {
    gpu_check(bbgpu_cma_update(_dev, _sigma, _fev));
    fetchReport();
    updateSigma();
}

bool CholeskyCmaes::converged() {
    if (!_rep_fresh) fetchReport();
    if (std::abs(_ybw[0] - _ybw[3]) > _tol) return false;
    // standard deviation of the candidates' radii (:146-160): M2 <= (lambda - 1) stol^2
    double mean = 0., m2 = 0.;
    gpu_check(bbgpu_cma_swarm_radius(_dev, &mean, &m2));
    return m2 <= (_lambda - 1) * _stol * _stol;
}

// ---- subclass scalars of a checkpoint
void Cmaes::extraState(StateIO &io) {
    io.pod(_flag); io.pod(_eigenlastev); io.pod(_eigenfreq); io.pod(_eigenrate); io.pod(_eigcount);
}

void ActiveCmaes::extraState(StateIO &io) {
    Cmaes::extraState(io);
    io.pod(_alphacov); io.pod(_cm); io.pod(_cneg); io.pod(_alphaold);
}

void SepCmaes::extraState(StateIO &io) {
    io.flag(_adjustlr); io.pod(_flag); io.pod(_ccov);
}

void CholeskyCmaes::extraState(StateIO &io) {
    io.pod(_stol);
}

// ---------------------------------------------------------------------------------------
// LmCmaes -- cma/lm_cmaes.cpp
// ---------------------------------------------------------------------------------------
void LmCmaes::init(const multivariate_problem &f, const double *guess) { // :42-89
    BaseCmaes::init(f, guess);
    if (_adaptmemory) _memsize = large_size(_n);
    _memlen = 0;
    if (_new) {
        _nsteps = _n;
        _t = std::max(1, static_cast<int>(std::log(1. * _n)));
        _cc = 0.5 / std::sqrt(1. * _n);
    } else {
        _nsteps = _memsize;
        _t = 1;
        _cc = 1. / _memsize;
    }
    _cs = 0.3;
    _c1 = 0.1 / std::log(_n + 1.);
    _cmu = NAN;
    _ccc = std::sqrt(_cc * (2. - _cc) * _mueff);
    _damps = 1.;
    _sqrt1mc1 = std::sqrt(1. - _c1);
    _zstar = 0.25;
    _s = 0.;
    _jarr.assign(_memsize, 0);
    _larr.assign(_memsize, 0);
    _fp.assign(_lambda, 0.);
    _fcur.assign(_lambda, 0.);
    ensureDevice();
}

void LmCmaes::fillConfig(bbgpu_cma_config &cfg) const {
    BaseCmaes::fillConfig(cfg);
    cfg.cmu = 0.; // NAN in the reference (:65), never used
    cfg.lm_memory = _memsize;
    cfg.lm_rademacher = _samplemode;
    cfg.lm_usenew = _new ? 1 : 0;
}

// :91-127.  Replay: per pair of candidates n base draws -- N(0,1) from _Z or Rademacher signs from uniform_real(0,1) < 0.5 --
// then, in the 2015 variant with more than one live vector, the one normal of selectSubset (:308-322); all on the same
// engine, in this order.
void LmCmaes::samplePopulation() {
    _dev_used = true;
    if (_replay.enabled) {
        const int pairs = (_lambda + 1) / 2;
        _zbuf.resize((size_t) pairs * _n);
        std::vector<int32_t> i0(pairs, 0);
        std::uniform_real_distribution<double> u01(0., 1.);
        for (int p = 0; p < pairs; p++) {
            double *z = _zbuf.data() + (size_t) p * _n;
            if (_samplemode == 0) {
                for (int i = 0; i < _n; i++) z[i] = _replay.normal(_replay.eng());
            } else {
                for (int i = 0; i < _n; i++) z[i] = u01(_replay.eng()) < 0.5 ? 1. : -1.;
            }
            if (_new && _memlen > 1) {
                const int msigma = p == 0 ? 40 : 4;
                int mstar = static_cast<int>(msigma * std::abs(_replay.normal(_replay.eng())));
                mstar = std::min(mstar, _memlen);
                i0[p] = _memlen - mstar;
            }
        }
        gpu_check(bbgpu_cma_lm_sample(_dev, _sigma, _gen, _memlen, _jarr.data(), _zbuf.data(), i0.data()));
    } else {
        gpu_check(bbgpu_cma_lm_sample(_dev, _sigma, _gen, _memlen, _jarr.data(), nullptr, nullptr));
    }
    _gen++;
}

// :228-238: the previous generation's sorted fitness is kept for the pooled ranking
void LmCmaes::evaluateAndSortPopulation() {
    if (_it > 0) _fp = _fcur;
    BaseCmaes::evaluateAndSortPopulation();
    // (the sorted fitness is read in updateDistribution, together with the report: one host round trip per generation)
}

// :274-306 (algorithm 5)
int LmCmaes::updateSet() {
    const int it = _it / _t;
    int imin = 1;
    if (it < _memsize) {
        _jarr[it] = it;
    } else if (_memsize > 1) {
        int iminval = _larr[_jarr[1]] - _larr[_jarr[0]];
        for (int i = 1; i < _memsize - 1; i++) {
            const int val = _larr[_jarr[i + 1]] - _larr[_jarr[i]];
            if (val < iminval) {
                iminval = val;
                imin = i + 1;
            }
        }
        if (iminval >= _nsteps) imin = 0;
        const int jtmp = _jarr[imin];
        for (int i = imin; i < _memsize - 1; i++) _jarr[i] = _jarr[i + 1];
        _jarr[_memsize - 1] = jtmp;
    }
    const int jcur = _jarr[std::min(it, _memsize - 1)];
    _larr[jcur] = _it;
    return imin == 1 ? 0 : imin;
}

// :129-181
void LmCmaes::updateDistribution() {
    const bool due = _it % _t == 0;
    int imin = 0;
    if (due) {
        imin = updateSet();
        if (_memlen < _memsize) _memlen++;
    }
    gpu_check(bbgpu_cma_lm_update(_dev, _sigma, due ? 1 : 0, imin, _memlen, _jarr.data()));
    gpu_check(bbgpu_cma_get(_dev, BBGPU_CMA_RANKVAL, _fcur.data(), _lambda)); // this generation's sorted fitness (:228-238)
    fetchReport();
    updateSigma();
}

// :240-258: rank-based success rule on the pooled previous + current fitness
void LmCmaes::updateSigma() {
    if (_it == 0) return;
    struct pooled { int _index; double _value; }; // cmaes_index, base_cmaes.h:30-38
    std::vector<pooled> mixed(2 * (size_t) _lambda);
    for (int n = 0; n < _lambda; n++) {
        mixed[n] = pooled { n, _fp[n] };
        mixed[n + _lambda] = pooled { n + _lambda, _fcur[n] };
    }
    std::sort(mixed.begin(), mixed.end(), [](const pooled &x, const pooled &y) { return x._value < y._value; });
    double zpsr = 0.;
    for (int n = 0; n < (_lambda << 1); n++) {
        const double f = (1. * n) / _lambda;
        if (mixed[n]._index < _lambda) zpsr += f;
        else zpsr -= f;
    }
    zpsr /= (1. * _lambda);
    zpsr -= _zstar;
    _s = (1. - _cs) * _s + _cs * zpsr;
    _sigma *= std::exp(_s / _damps);
}

// :183-226
bool LmCmaes::converged() {
    if (_it >= _mit) return true;
    if (_sigma < _stolmin) return true;
    if (_it >= _hlen && _fworst - _fbest < _tol) return true;
    if (_best._len >= _n && _kth._len >= _n) {
        int countEq = 0;
        for (int i = 0; i < _n; i++) {
            if (_best.get(i) == _kth.get(i)) {
                countEq++;
                if (3 * countEq >= _n) return true;
            }
        }
    }
    return false;
}

void LmCmaes::extraState(StateIO &io) {
    io.flag(_new); io.flag(_adaptmemory);
    io.pod(_samplemode); io.pod(_memlen); io.pod(_memsize); io.pod(_nsteps); io.pod(_t);
    io.pod(_stolmin); io.pod(_sqrt1mc1); io.pod(_zstar); io.pod(_s); io.pod(_ccc);
    io.vec(_jarr); io.vec(_larr); io.vec(_fp); io.vec(_fcur);
}
