// Host class of SaNSDE over the C ABI (include/bbgpu.h, bbgpu_sde_*).
// Same constructor signature and public interface as the reference (/root/reference/src/multivariate/de/sansde.h:44-90): the
// host keeps the adaptation state (p, fp, crm and the sums behind them), the GPU keeps the swarm and resolves the in-place
// generation.
#ifndef BBOPT_HOST_SANSDE_H_
#define BBOPT_HOST_SANSDE_H_

#include <functional>
#include <vector>

#include "device.h"
#include "multivariate.h"
#include "random.hpp"

// tagged draws of one generation, as the oracle's recorder produces them (draw replay)
struct sansde_generation_draws {
    std::vector<double> zcr, F, u;
    std::vector<int32_t> ifstrat, r1, r2, r3, istrat, jrand;
};
typedef std::function<bool(int /*generation*/, sansde_generation_draws&)> sansde_replay_fn;

class SaNSDESearch: public MultivariateOptimizer {
protected:
    bool _repaircr;
    int _n, _np, _it, _fev, _mfev, _ncrref, _npup, _ncrup;
    double _tol, _p, _fp, _crm;
    bbgpu_sde_accum _acc;
    multivariate_problem _f;
    std::vector<double> _lower, _upper;
    bbgpu_sde *_dev = nullptr;
    bbopt::ObjectiveHandle _obj;
    uint64_t _stream_key = 0;
    bool _stream_key_fixed = false;
    sansde_replay_fn _replay;
    std::vector<double> _initial;
    int _rounds = 0;
    bool _radius_fresh = false;
    bool converged();

public:
    SaNSDESearch(int mfev, int np, double tol, bool repaircr = true, int crref = 5, int pupdate = 50, int crupdate = 25);
    virtual ~SaNSDESearch();
    void init(const multivariate_problem &f, const double *guess);
    void iterate();
    multivariate_solution solution();
    multivariate_solution optimize(const multivariate_problem &f, const double *guess);
    // additions: replay hooks and state access
    void setDrawReplay(sansde_replay_fn fn) { _replay = fn; }
    void setInitialPopulation(const std::vector<double> &x) { _initial = x; }
    void setStreamKey(uint64_t key) { _stream_key = key; _stream_key_fixed = true; }
    std::vector<double> swarm();
    std::vector<double> fitness();
    std::vector<double> crossoverRates();
    int evaluations() const { return _fev; }
    int population() const { return _np; }
    int rounds() const { return _rounds; }
    double p() const { return _p; }
    double fp() const { return _fp; }
    double crm() const { return _crm; }
    const bbgpu_sde_accum &sums() const { return _acc; }
    bool isConverged() { return converged(); }
};

#endif
