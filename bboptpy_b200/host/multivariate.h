// Optimizer interface of the accelerated path.
//
// Mirrors the reference's drop-in boundary (/root/reference/src/multivariate/multivariate.h:
// `multivariate` :32, `multivariate_problem` :38-79, `multivariate_solution` :81-115,
// `MultivariateOptimizer` :132-146): same type names, same member names, same virtuals, so
// the pybind layer and user C++ code read the same.  The one addition is `_dev`: a device
// objective descriptor.  When it is set the population is evaluated on the GPU (registry
// objective or user CUDA function); when only the CPU function `_f` is given it is driven
// through the batched host-callback bridge (one device->host copy of the whole population per
// generation instead of the reference's per-candidate numpy copy, py/multivariate_py.cpp:385-388).
#ifndef BBOPT_HOST_MULTIVARIATE_H_
#define BBOPT_HOST_MULTIVARIATE_H_

#include <functional>
#include <memory>
#include <string>
#include <vector>

#include "bbgpu.h"

typedef std::function<double(const double*)> multivariate;
typedef std::function<void(const double*, double*)> constraints;
typedef std::function<bool(const double*)> blackbox_constraints;

// where the objective runs
struct device_objective {
    int _builtin = -1;                               // bbgpu_builtin_objective, or -1
    bbgpu_device_objective_fn _fn = nullptr;         // user CUDA objective (device pointers)
    void *_user = nullptr;
    bool onDevice() const { return _builtin >= 0 || _fn != nullptr; }
};

struct multivariate_problem {
    multivariate _f;
    int _n;
    double *_lower;
    double *_upper;
    bool _hasc;
    constraints _c;
    int _neq, _nineq;
    bool _hasbbc;
    blackbox_constraints _bbc;
    device_objective _dev;

    multivariate_problem() : _n(0), _lower(nullptr), _upper(nullptr), _hasc(false), _neq(0), _nineq(0), _hasbbc(false) {}

    multivariate_problem(const multivariate f, const int n, double *lower, double *upper) :
            _f(f), _n(n), _lower(lower), _upper(upper), _hasc(false), _neq(0), _nineq(0), _hasbbc(false) {}

    multivariate_problem(const device_objective dev, const int n, double *lower, double *upper) :
            _n(n), _lower(lower), _upper(upper), _hasc(false), _neq(0), _nineq(0), _hasbbc(false), _dev(dev) {}
};

struct multivariate_solution {
    const std::vector<double> _sol;
    const int _fev, _cev, _bbev;
    const bool _converged;

    multivariate_solution(const std::vector<double> &sol, const int fev, const bool converged) :
            _sol(sol), _fev(fev), _cev(0), _bbev(0), _converged(converged) {}

    // same text layout as the reference's toString() (multivariate.h:98-114)
    std::string toString() {
        std::string s = "x*: ";
        for (double v : _sol) s += std::to_string(v) + " ";
        s += "\nobjective calls: " + std::to_string(_fev);
        s += "\nconstraint calls: " + std::to_string(_cev);
        s += "\nB/B constraint calls: " + std::to_string(_bbev);
        s += "\nconverged: ";
        s += _converged ? "yes" : "no/unknown";
        return s;
    }
};

class MultivariateOptimizer {
public:
    virtual ~MultivariateOptimizer() {}
    virtual void init(const multivariate_problem &f, const double *guess) = 0;
    virtual void iterate() = 0;
    virtual multivariate_solution solution() = 0;
    virtual multivariate_solution optimize(const multivariate_problem &f, const double *guess) = 0;
};

#endif
