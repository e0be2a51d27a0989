// Host class of SSDE over the C ABI (include/bbgpu.h, bbgpu_ssde_*).
// Same constructor signature and public interface as the reference (/root/reference/src/multivariate/de/ssde.h:42-83): the
// host keeps the evaluation counter, the patience counter, the population-reduction schedule and the shuffled index list
// behind the random rotation; the GPU keeps the swarm and the success memories and resolves the in-place generation.
#ifndef BBOPT_HOST_SSDE_H_
#define BBOPT_HOST_SSDE_H_

#include <functional>
#include <vector>

#include "device.h"
#include "multivariate.h"
#include "random.hpp"

// tagged draws of one generation, as the oracle's recorder produces them (draw replay)
struct ssde_generation_draws {
    std::vector<int32_t> iwork, iL, p, q, r, pbest, j0;
    std::vector<double> zrank, ub, cauchy, zcr, ucr, redraw;
};
typedef std::function<bool(int /*generation*/, ssde_generation_draws&)> ssde_replay_fn;

class SSDESearch: public MultivariateOptimizer {
protected:
    bool _usede, _repaircr;
    int _n, _np, _npmin, _npinit, _mfev, _fev, _patience, _convcount, _H, _k, _kCR, _it;
    double _tol, _ptop, _bestf, _rmean = 0., _rm2 = 0.;
    bool _radius_fresh = false;
    multivariate_problem _f;
    std::vector<int32_t> _iwork;
    std::vector<double> _lower, _upper;
    bbgpu_ssde *_dev = nullptr;
    bbopt::ObjectiveHandle _obj;
    uint64_t _stream_key = 0;
    bool _stream_key_fixed = false;
    ssde_replay_fn _replay;
    std::vector<double> _initial;
    int _rounds = 0, _nsucc = 0, _np_trace = 0;
    bool converged();

public:
    SSDESearch(int mfev, int npinit, double tol, int patience = 1000, int npmin = 4, double ptop = 0.11, int h = 100,
            bool usede = true, bool repaircr = true);
    virtual ~SSDESearch();
    void init(const multivariate_problem &f, const double *guess);
    void iterate();
    multivariate_solution solution();
    multivariate_solution optimize(const multivariate_problem &f, const double *guess);
    // additions: replay hooks and state access
    void setDrawReplay(ssde_replay_fn fn) { _replay = fn; }
    void setInitialPopulation(const std::vector<double> &x) { _initial = x; }
    void setStreamKey(uint64_t key) { _stream_key = key; _stream_key_fixed = true; }
    std::vector<double> swarm();
    std::vector<double> fitness();
    std::vector<double> memory(int which); // 0: L1 (mask probability), 1: L2 (step), 2: LCR
    // outcome of the last generation by candidate, before the sort: 0 rejected / 1 first stage / 2 second stage, last trial fitness
    void trace(std::vector<int32_t> &succ, std::vector<double> &ftrial);
    int evaluations() const { return _fev; }
    int population() const { return _np; }
    int rounds() const { return _rounds; }
    int accepted() const { return _nsucc; }
    int convcount() const { return _convcount; }
    int k() const { return _k; }
    int kCR() const { return _kCR; }
    bool isConverged() { return converged(); }
};

#endif
