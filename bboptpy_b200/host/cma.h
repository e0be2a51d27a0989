// Host classes of the CMA-ES family over the C ABI (include/bbgpu.h).
//
// Same class names, constructor signatures, public virtuals and semantics as the reference
// (/root/reference/src/multivariate/cma/base_cmaes.h:50-92, cmaes.h:40-70,
// active_cmaes.h:40-56, sep_cmaes.h:37-61, cholesky_cmaes.h:32-58).  The host keeps what the reference keeps in scalars
// (sigma, counters, history rings, stop flags); the vectors and matrices live on the GPU.
#ifndef BBOPT_HOST_CMA_H_
#define BBOPT_HOST_CMA_H_

#include <cmath>
#include <functional>
#include <string>
#include <vector>

#include "device.h"
#include "multivariate.h"
#include "random.hpp"

struct cmaes_history {
    int _cap, _buffer, _len;
    std::vector<double> _hist;
    void add(double value);
    double get(int i);
};

// eigenbasis replay hook: called at every decomposition event with (fev, n); returns true and
// fills B (n*n, B[i*n+j] = component i of eigenvector j) and D (n) to impose a basis
typedef std::function<bool(int, int, std::vector<double>&, std::vector<double>&)> eigen_replay_fn;

class BaseCmaes: public MultivariateOptimizer {
protected:
    bool _adaptpop, _adaptit, _bound;
    int _n, _mfev, _mit, _lambda, _mu, _it, _hlen, _ik, _fev;
    double _tol, _sigma0, _mueff, _cc, _cs, _c1, _cmu, _damps, _chi, _sigma, _fbest, _fworst;
    multivariate_problem _f;
    cmaes_history _best, _kth;
    std::vector<int> _ibw;
    std::vector<double> _lower, _upper, _ybw, _weights;
    // device side
    bbgpu_cma *_dev = nullptr;
    int _dev_kind = -1, _dev_n = 0, _dev_lambda = 0;
    bool _dev_used = false; // a previous run left B/C in the handle (re-init quirk, cmaes.cpp:53-60)
    bbopt::ObjectiveHandle _obj;
    bbgpu_cma_report _rep;
    bool _rep_fresh = false;
    uint64_t _stream_key = 0;
    double _fkth = 0.;
    bbopt::ReplayStream _replay;
    std::vector<double> _zbuf, _xguess;
    uint64_t _gen = 0; // generation counter of the Philox stream; never reset, so restarts draw fresh values
    bbgpu_comm *_comm = nullptr; // candidate sharding over the ranks of this communicator (not owned)

    virtual int deviceKind() const = 0;
    virtual void fillConfig(bbgpu_cma_config &cfg) const;
    // subclass scalars of a checkpoint, in a fixed order (same walk for writing and reading)
    struct StateIO;
    virtual void extraState(StateIO &io);
    void restoreOrSaveScalars(StateIO &io);
    virtual int memoryRows() const { return 0; } // LM kind: rows of the direction-vector matrices
    void ensureDevice();
    void fetchReport();

public:
    BaseCmaes(int mfev, double tol, int np, double sigma0 = 2., bool bound = false);
    virtual ~BaseCmaes();

    void init(const multivariate_problem &f, const double *guess);
    void iterate();
    multivariate_solution solution();
    multivariate_solution optimize(const multivariate_problem &f, const double *guess);
    void setParams(int np, double sigma, int mfev);

    virtual void samplePopulation();
    virtual void updateDistribution() = 0;
    virtual bool converged() = 0;
    virtual void updateSigma();
    virtual void updateHistory();
    virtual void evaluateAndSortPopulation();
    std::vector<double> bestSolution();

    // ---- additions (not in the reference): replay hooks and state access for parity checks
    void setDrawReplay(uint32_t seed) { _replay.enable(seed); }
    void setDrawReplayGlobal() { _replay.enable_global(); }
    // advance the replay engine past `count` uniform_real draws the caller made on the reference side
    void burnUniforms(int count) {
        std::uniform_real_distribution<double> u(0., 1.);
        for (int i = 0; i < count; i++) u(_replay.eng());
    }
    std::vector<double> state(const std::string &name);
    // ---- checkpoint / resume (the reference has no on-disk format: random.hpp:281-300 can serialise
    // the engine but nobody calls it).  A versioned little-endian blob: header ("BBOPTCK1", version, kind, n, lambda, mu),
    // every host scalar of the run (sigma, counters, strategy constants, history rings, stop flags, the Philox key and
    // generation counter of the free-running stream) and the device state vectors / matrices of the kind.  Contract:
    // construct an optimizer of the same class, initialize(f, lower, upper, guess) it, then restore(blob); iterate() then
    // continues exactly where the checkpointed run stood (bit for bit with free-running draws: the Philox stream is a
    // function of (key, generation, candidate), not of history).  Layout: INTEGRATION.md.
    std::string checkpoint();
    void restore(const std::string &blob);
    double sigma() const { return _sigma; }
    int evaluations() const { return _fev; }
    int iterations() const { return _it; }
    int lambda() const { return _lambda; }
    bbopt::ObjectiveHandle &objective() { return _obj; }
    // One large-lambda run sharded by candidates over the GPUs of `comm` (one process per GPU): every rank
    // constructs the same optimizer, calls this, then optimize()/iterate() with the same arguments; state is
    // replicated, each rank samples and evaluates lambda/world candidates.  Needs setStreamKey() with one
    // common key so that all ranks draw from the same Philox stream.  Call before init().
    void setCommunicator(bbgpu_comm *comm) { _comm = comm; }
    void setStreamKey(uint64_t key) { _stream_key = key; }
};

class Cmaes: public BaseCmaes {
protected:
    int _flag, _eigenlastev;
    double _eigenfreq, _eigenrate;
    eigen_replay_fn _eigreplay;
    int _eigcount = 0;
    int deviceKind() const override { return BBGPU_CMA_FULL; }
    void pathsAndCovariance();
    void extraState(StateIO &io) override;

public:
    Cmaes(int mfev, double tol, int np, double sigma0 = 2., bool bound = false, double eigenrate = 0.25) :
            BaseCmaes(mfev, tol, np, sigma0, bound) {
        _eigenrate = eigenrate;
    }
    void init(const multivariate_problem &f, const double *guess);
    void updateDistribution();
    bool converged();
    void eigenDecomposition();
    void setEigenReplay(eigen_replay_fn fn) { _eigreplay = fn; }
    int flag() const { return _flag; }
    int decompositions() const { return _eigcount; }
};

class ActiveCmaes: public Cmaes {
protected:
    double _alphacov, _cm, _cneg, _alphaold;
    int deviceKind() const override { return BBGPU_CMA_ACTIVE; }
    void fillConfig(bbgpu_cma_config &cfg) const override;
    void extraState(StateIO &io) override;

public:
    ActiveCmaes(int mfev, double tol, int np, double sigma0 = 2., bool bound = false, double alphacov = 2.,
            double eigenrate = 0.25) :
            Cmaes(mfev, tol, np, sigma0, bound, eigenrate) {
        _alphacov = alphacov;
    }
    void init(const multivariate_problem &f, const double *guess);
};

class SepCmaes: public BaseCmaes {
protected:
    bool _adjustlr;
    int _flag;
    double _ccov;
    int deviceKind() const override { return BBGPU_CMA_SEP; }
    void fillConfig(bbgpu_cma_config &cfg) const override;
    void extraState(StateIO &io) override;

public:
    SepCmaes(int mfev, double tol, int np, double sigma0 = 2., bool bound = false, bool adjustlr = false) :
            BaseCmaes(mfev, tol, np, sigma0, bound) {
        _adjustlr = adjustlr;
    }
    void init(const multivariate_problem &f, const double *guess);
    void updateDistribution();
    bool converged();
    int flag() const { return _flag; }
};

// cma/cholesky_cmaes.h:32-58.  The factor A lives on the device; the reference's mu + 1 sequential rank-one updates are
// one SYRK + one Cholesky factorisation there (csrc/chol.cuh), so a generation needs no eigendecomposition.
class CholeskyCmaes: public BaseCmaes {
protected:
    double _stol;
    int deviceKind() const override { return BBGPU_CMA_CHOLESKY; }
    void extraState(StateIO &io) override;

public:
    CholeskyCmaes(int mfev, double tol, double stol, int np, double sigma0 = 2., bool bound = false) :
            BaseCmaes(mfev, tol, np, sigma0, bound) {
        _stol = stol;
    }
    void init(const multivariate_problem &f, const double *guess);
    void updateDistribution();
    bool converged();
};

// cma/lm_cmaes.h:36-92 (Loshchilov 2014 / 2015).  Host: the memory bookkeeping (jarr / larr, algorithm 5), the pooled-ranking
// step size (:240-258) and the stop tests; device: mirrored sampling through the direction vectors and their rebuild.
class LmCmaes: public BaseCmaes {
protected:
    bool _new, _adaptmemory;
    int _samplemode, _memlen, _memsize, _nsteps, _t;
    double _stolmin, _sqrt1mc1, _zstar, _s, _ccc;
    std::vector<int> _jarr, _larr;
    std::vector<double> _fp, _fcur;
    int deviceKind() const override { return BBGPU_CMA_LM; }
    int memoryRows() const override { return _memsize; }
    void fillConfig(bbgpu_cma_config &cfg) const override;
    void extraState(StateIO &io) override;
    int updateSet();

public:
    LmCmaes(int mfev, double tol, int np, int memory = 0, double sigma0 = 2., bool bound = false, bool rademacher = true,
            bool usenew = true) :
            BaseCmaes(mfev, tol, np, sigma0, bound) {
        _samplemode = rademacher ? 1 : 0;
        _new = usenew;
        _stolmin = 1e-16;
        _memsize = memory;
        _adaptmemory = memory <= 0;
        _memlen = _nsteps = 0;
        _t = 1;
        _sqrt1mc1 = _zstar = _s = _ccc = 0.;
    }
    static int large_size(const int x) { return static_cast<int>(2. * std::sqrt(1. * x)); }
    void init(const multivariate_problem &f, const double *guess);
    void samplePopulation();
    void updateDistribution();
    bool converged();
    void evaluateAndSortPopulation();
    void updateSigma();
    int memoryLength() const { return _memlen; }
};

#endif
