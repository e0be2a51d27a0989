// Host classes of the DE family over the C ABI: JADE and L-SHADE.
// Same constructor signatures and public interface as the reference
// (/root/reference/src/multivariate/de/jade.h:49-87, shade.h:41-75).
#ifndef BBOPT_HOST_DE_H_
#define BBOPT_HOST_DE_H_

#include <functional>
#include <vector>

#include "device.h"
#include "multivariate.h"
#include "random.hpp"

// tagged draws of one generation, as the oracle recorder produces them (draw replay)
struct de_generation_draws {
    std::vector<double> zcr, cauchy, u, pi;
    std::vector<int32_t> ibest, r1, r2, jrand, slot, ri, trim;
};
typedef std::function<bool(int /*generation*/, de_generation_draws&)> de_replay_fn;

class DeSearchBase: public MultivariateOptimizer {
protected:
    bool _archive, _repaircr;
    int _n, _np, _npinit, _fev, _mfev, _h, _npmin, _gen;
    double _pelite, _c, _tol, _sigma;
    multivariate_problem _f;
    std::vector<double> _lower, _upper;
    bbgpu_de *_dev = nullptr;
    bbopt::ObjectiveHandle _obj;
    bbgpu_de_report _rep;
    uint64_t _stream_key;
    bool _stream_key_fixed = false;
    de_replay_fn _replay;
    std::vector<double> _initial;
    virtual int kind() const = 0;
    virtual const char *label() const = 0;
    bool converged();
public:
    DeSearchBase();
    virtual ~DeSearchBase();
    void init(const multivariate_problem &f, const double *guess);
    void iterate();
    multivariate_solution solution();
    multivariate_solution optimize(const multivariate_problem &f, const double *guess);
    // additions: replay hooks and state access
    void setDrawReplay(de_replay_fn fn) { _replay = fn; }
    void setInitialPopulation(const std::vector<double> &x) { _initial = x; }
    void setStreamKey(uint64_t key) { _stream_key = key; _stream_key_fixed = true; }
    std::vector<double> swarm();
    std::vector<double> fitness();
    const bbgpu_de_report &report() const { return _rep; }
    int evaluations() const { return _fev; }
    int population() const { return _np; }
};

class JadeSearch: public DeSearchBase {
    int kind() const override { return BBGPU_DE_JADE; }
    const char *label() const override { return "JADE"; }
public:
    JadeSearch(int mfev, int np, double tol, bool archive = true, bool repaircr = true, double pelite = 0.05,
            double cdamp = 0.1, double sigma = 0.07);
};

class ShadeSearch: public DeSearchBase {
    int kind() const override { return BBGPU_DE_SHADE; }
    const char *label() const override { return "SHADE"; }
public:
    ShadeSearch(int mfev, int npinit, double tol, bool archive = true, bool repaircr = true, int h = 100, int npmin = 4);
};

#endif
