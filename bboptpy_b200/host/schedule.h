// Restart schedules of IPOP / NIPOP and BIPOP / NBIPOP as state machines, detached from HOW a base run executes.
//
// The reference interleaves the schedule with the base optimizer call (ipop_cmaes.cpp:65-162, bipop_cmaes.cpp:61-267:
// setParams -> optimize -> f(solution) -> bookkeeping).  Here the bookkeeping is one object per restart sequence:
//     request r = s.first(n);            // init(): the default-lambda run from the user's guess
//     ... run a base CMA-ES with (r.lambda, r.sigma, r.mfev), get (fev, fx = f(best)) ...
//     s.finish(fev, fx);                 // budgets, best-of tracking, +1 evaluation
//     r = s.next(lower, upper, x0, u);   // iterate(): x0 ~ U(lower, upper), regime choice, lambda / sigma / budget
// so the same code drives one run behind the reference API (IPopCmaes / BiPopCmaes, restart.cpp: base run on one device
// handle) and thousands of runs of a batch (batch.cpp: base runs as CTAs of the batched engine).  Draw order inside
// next() is the reference's: n uniforms for x0 first, then (BIPOP, small regime) one for lambda, one for sigma.
#ifndef BBOPT_HOST_SCHEDULE_H_
#define BBOPT_HOST_SCHEDULE_H_

#include <algorithm>
#include <cmath>
#include <functional>
#include <vector>

struct base_run_request {
    int lambda;
    double sigma;
    int mfev;
    int regime; // 0 = the initial run, 1 = large-population regime (IPOP: every restart), 2 = small-population regime
};

typedef std::function<double(double, double)> uniform_fn; // uniform_real(a, b) draw

// ipop_cmaes.cpp
class IPopSchedule {
public:
    int _n = 0, _mfev = 0, _it = 0, _fev = 0, _lambdadef = 0, _lambdamax = 0, _lambda = 0;
    double _sigmadef = 2., _sigma = 0., _ksigmadec = 1.6, _fbest = 0., _fx = 0.;
    bool _nipop = true, _boundlambda = true;

    void configure(int mfev, double sigma0, bool nipop, double ksigmadec, bool boundlambda) {
        _mfev = mfev;
        _sigmadef = sigma0;
        _nipop = nipop;
        _ksigmadec = ksigmadec;
        _boundlambda = boundlambda;
    }
    // ipop_cmaes.cpp:74-92
    base_run_request first(int n) {
        _n = n;
        _fev = 0;
        _it = -1;
        _lambda = _lambdadef = 4 + static_cast<int>(3. * std::log(1. * _n));
        _lambdamax = 10 * _n * _n;
        _sigma = _sigmadef;
        return {_lambdadef, _sigmadef, maxEvaluations(_lambdadef), 0};
    }
    // ipop_cmaes.cpp:112-145
    base_run_request next(const double *lower, const double *upper, std::vector<double> &x0, const uniform_fn &u) {
        x0.resize(_n);
        for (int i = 0; i < _n; i++) x0[i] = u(lower[i], upper[i]);
        _lambda <<= 1;
        if (_boundlambda && _lambda > _lambdamax) {
            if (_lambda - _lambdamax < _lambdamax - (_lambda >> 1)) _lambda = _lambdamax;
            else _lambda = _lambdadef;
        }
        if (_nipop) {
            _sigma /= _ksigmadec;
            _sigma = std::max(_sigma, 0.01 * _sigmadef);
        }
        return {_lambda, _sigma, maxEvaluations(_lambda), 1};
    }
    // ipop_cmaes.cpp:93-99,146-158; true when this run is the best so far (the caller keeps its x)
    bool finish(int fev, double fx) {
        _fx = fx;
        _fev += fev + 1;
        const bool initial = _it < 0;
        _it++;
        if (initial || fx < _fbest) {
            _fbest = fx;
            return true;
        }
        return false;
    }
    bool done() const { return _fev >= _mfev; } // ipop_cmaes.cpp:171
    // ipop_cmaes.cpp:178-189
    int maxEvaluations(int lambda) const {
        const int maxit = static_cast<int>(100. + 50. * (_n + 3) * (_n + 3) / std::sqrt(1. * lambda));
        return std::min(maxit * lambda, _mfev - _fev);
    }
};

// bipop_cmaes.cpp
class BiPopSchedule {
public:
    int _n = 0, _mfev = 0, _fev = 0, _maxlargeruns = 9, _lambdadef = 0, _largelambda = 0, _smalllambda = 0,
            _largebudget = 0, _smallbudget = 0, _largerestarts = 0, _smallrestarts = 0, _it = 0, _bestregime = 1,
            _regime = 0;
    double _fx = 0., _sigmadef = 2., _largesigma = 0., _smallsigma = 0., _fxbest = 0., _ksigmadec = 1.6, _kbudget = 2.;
    bool _nbipop = true;

    void configure(int mfev, double sigma0, int maxlargeruns, bool nbipop, double ksigmadec, double kbudget) {
        _mfev = mfev;
        _sigmadef = sigma0;
        _maxlargeruns = maxlargeruns;
        _nbipop = nbipop;
        _ksigmadec = ksigmadec;
        _kbudget = kbudget;
    }
    // bipop_cmaes.cpp:70-92
    base_run_request first(int n) {
        _n = n;
        _fev = 0;
        _lambdadef = 4 + static_cast<int>(3. * std::log(1. * _n));
        _largebudget = _smallbudget = 0;
        _largerestarts = _smallrestarts = 0;
        _largelambda = _smalllambda = 0;
        _bestregime = 1;
        _it = -1;
        _regime = 0;
        return {_lambdadef, _sigmadef, maxEvaluations(_lambdadef), 0};
    }
    // bipop_cmaes.cpp:109-131, 204-216, 238-251
    base_run_request next(const double *lower, const double *upper, std::vector<double> &x0, const uniform_fn &u) {
        x0.resize(_n);
        for (int i = 0; i < _n; i++) x0[i] = u(lower[i], upper[i]);
        if (_nbipop) {
            if (_bestregime == 1) _regime = (_largebudget <= _smallbudget * _kbudget) ? 1 : 2;
            else _regime = (_smallbudget <= _kbudget * _largebudget) ? 2 : 1;
        } else {
            _regime = (_largebudget <= _smallbudget) ? 1 : 2;
        }
        if (_regime == 1) {
            _largelambda = static_cast<int>(_lambdadef * std::pow(2, _largerestarts + 1));
            if (_nbipop) {
                _largesigma = _sigmadef * std::pow(1. / _ksigmadec, _largerestarts + 1);
                _largesigma = std::max(_largesigma, 0.01 * _sigmadef);
            } else {
                _largesigma = _sigmadef;
            }
            return {_largelambda, _largesigma, maxEvaluations(_largelambda), 1};
        }
        const double v = u(0., 1.);
        _smalllambda = static_cast<int>(_lambdadef * std::pow((0.5 * _largelambda) / _lambdadef, v * v));
        _smallsigma = _sigmadef * std::pow(10., -2. * u(0., 1.));
        int maxfev = maxEvaluations(_smalllambda);
        maxfev = std::min(maxfev, _largebudget >> 1);
        return {_smalllambda, _smallsigma, maxfev, 2};
    }
    // bipop_cmaes.cpp:84-92, 217-235, 252-266
    bool finish(int fev, double fx) {
        _fx = fx;
        _fev += fev + 1;
        _it++;
        if (_regime == 0) {
            _fxbest = fx;
            return true;
        }
        bool best = false;
        if (fx < _fxbest) {
            _fxbest = fx;
            _bestregime = _regime;
            best = true;
        }
        if (_regime == 1) {
            _largebudget += fev;
            _largerestarts++;
        } else {
            _smallbudget += fev;
            _smallrestarts++;
        }
        return best;
    }
    // the exits of optimize() (bipop_cmaes.cpp:170-189), tested after every iterate()
    bool hitLargeRunLimit() const { return _largerestarts >= _maxlargeruns; }
    bool done() const { return hitLargeRunLimit() || _fev >= _mfev; }
    int currentLambda() const { return _regime == 0 ? _lambdadef : _regime == 1 ? _largelambda : _smalllambda; }
    double currentSigma() const { return _regime == 0 ? _sigmadef : _regime == 1 ? _largesigma : _smallsigma; }
    // bipop_cmaes.cpp:191-202
    int maxEvaluations(int lambda) const {
        const int maxit = static_cast<int>(100. + 50. * (_n + 3) * (_n + 3) / std::sqrt(1. * lambda));
        return std::min(maxit * lambda, _mfev - _fev);
    }
};

#endif
