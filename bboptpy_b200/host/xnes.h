// Host class of xNES over the C ABI (include/bbgpu.h, kind BBGPU_CMA_XNES).
// Same constructor signature and public interface as the reference (/root/reference/src/multivariate/nes/xnes.h:46-80);
// the host keeps sigma, the learning rates and the rank utilities, the GPU keeps mu, B and the points.
#ifndef BBOPT_HOST_XNES_H_
#define BBOPT_HOST_XNES_H_

#include <string>
#include <vector>

#include "device.h"
#include "multivariate.h"
#include "random.hpp"

class xNES: public MultivariateOptimizer {
protected:
    int _n, _mfev, _np, _fev;
    double _tol, _a0, _etamu, _etasigma, _etab, _sigma, _Gsigma;
    multivariate_problem _f;
    std::vector<double> _lower, _upper, _u;
    bbgpu_cma *_dev = nullptr;
    bbopt::ObjectiveHandle _obj;
    bbgpu_cma_report _rep;
    uint64_t _stream_key = 0, _gen = 0;
    bbopt::ReplayStream _replay;
    std::vector<double> _zbuf;
    bool converged();

public:
    xNES(int mfev, double tol, double a0 = 1., double etamu = 1.);
    virtual ~xNES();
    void init(const multivariate_problem &f, const double *guess);
    void iterate();
    multivariate_solution solution();
    multivariate_solution optimize(const multivariate_problem &f, const double *guess);
    // additions: replay hook and state access for parity checks
    void setDrawReplay(uint32_t seed) { _replay.enable(seed); }
    std::vector<double> state(const std::string &name);
    double sigma() const { return _sigma; }
    double gsigma() const { return _Gsigma; }
    int evaluations() const { return _fev; }
    int population() const { return _np; }
    bool isConverged() { return converged(); }
};

#endif
