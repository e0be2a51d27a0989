#include "restart.h"

#include <cmath>
#include <iomanip>
#include <iostream>
#include <limits>
#include <sstream>

using bbopt::Random;

void RowPrinter::printRow(const std::vector<std::string> &cells) {
    int total = 0;
    for (size_t i = 0; i < _w.size(); i++) {
        std::cout << " | " << std::setw(_w[i]) << (i < cells.size() ? cells[i] : "");
        total += _w[i];
    }
    std::cout << " | " << std::endl;
    if (_rows == 0) std::cout << " |" << std::string(total + 3 * ((int) _w.size() - 1) + 2, '=') << "| " << std::endl;
    _rows++;
}

std::string RowPrinter::num(double v) {
    std::stringstream ss;
    ss << std::setprecision(std::numeric_limits<double>::max_digits10) << v;
    return ss.str();
}

// ---------------------------------------------------------------------------------------
// IPOP / NIPOP -- cma/ipop_cmaes.cpp
// ---------------------------------------------------------------------------------------
IPopCmaes::IPopCmaes(BaseCmaes *base, int mfev, bool print, double sigma0, bool nipop, double ksigmadec,
        bool boundlambda) {
    _base = base;
    _mfev = mfev;
    _print = print;
    _sigmadef = sigma0;
    _nipop = nipop;
    _ksigmadec = ksigmadec;
    _boundlambda = boundlambda;
    _n = _it = _fev = _lambdadef = _lambdamax = _lambda = 0;
    _sigma = _fbest = _fx = 0.;
}

// ipop_cmaes.cpp:65-110
void IPopCmaes::init(const multivariate_problem &f, const double *guess) {
    if (f._hasc || f._hasbbc) {
        std::cerr << "Warning [IPOP-CMAES]: problem constraints will be ignored." << std::endl;
    }
    _f = f;
    _n = f._n;
    _lower = std::vector<double>(f._lower, f._lower + _n);
    _upper = std::vector<double>(f._upper, f._upper + _n);
    _guess = std::vector<double>(guess, guess + _n);
    _fev = 0;
    _lambda = _lambdadef = 4 + static_cast<int>(3. * std::log(1. * _n));
    _lambdamax = 10 * _n * _n;
    _sigma = _sigmadef;
    const int maxfev = computeMaxEvaluations(_lambdadef);
    _base->setParams(_lambdadef, _sigmadef, maxfev);
    const multivariate_solution sol = _base->optimize(_f, &_guess[0]);
    _fx = _base->objective().evalOne(&sol._sol[0]); // the reference's extra f(x) per run (:94)
    _fev += sol._fev + 1;
    _it = 0;
    _fbest = _fx;
    _xbest = sol._sol;
    _x0 = std::vector<double>(_n);
    if (_print) {
        _table.setWidth( { 5, 10, 5, 25, 25, 25 });
        _table.printRow( { "run", "budget", "pop", "sigma", "f*", "best f*" });
        _table.printRow( { RowPrinter::num(_it), RowPrinter::num(_fev), RowPrinter::num(_lambdadef),
                RowPrinter::num(_sigmadef), RowPrinter::num(_fx), RowPrinter::num(_fbest) });
    }
}

// ipop_cmaes.cpp:112-162
void IPopCmaes::iterate() {
    for (int i = 0; i < _n; i++) _x0[i] = Random::get(_lower[i], _upper[i]);
    if (_boundlambda) {
        _lambda <<= 1;
        if (_lambda > _lambdamax) {
            if (_lambda - _lambdamax < _lambdamax - (_lambda >> 1)) _lambda = _lambdamax;
            else _lambda = _lambdadef;
        }
    } else {
        _lambda <<= 1;
    }
    if (_nipop) {
        _sigma /= _ksigmadec;
        _sigma = std::max(_sigma, 0.01 * _sigmadef);
    }
    const int maxfev = computeMaxEvaluations(_lambda);
    _base->setParams(_lambda, _sigma, maxfev);
    const multivariate_solution sol = _base->optimize(_f, &_x0[0]);
    _fx = _base->objective().evalOne(&sol._sol[0]);
    _fev += sol._fev + 1;
    if (_fx < _fbest) {
        _fbest = _fx;
        _xbest = sol._sol;
    }
    _it++;
    if (_print) {
        _table.printRow( { RowPrinter::num(_it), RowPrinter::num(_fev), RowPrinter::num(_lambda), RowPrinter::num(_sigma),
                RowPrinter::num(_fx), RowPrinter::num(_fbest) });
    }
}

multivariate_solution IPopCmaes::solution() {
    return {_xbest, _fev, false};
}

multivariate_solution IPopCmaes::optimize(const multivariate_problem &f, const double *guess) {
    init(f, guess);
    while (_fev < _mfev) iterate();
    return {_xbest, _fev, false};
}

// ipop_cmaes.cpp:178-189
int IPopCmaes::computeMaxEvaluations(const int lambda) {
    const int maxit = static_cast<int>(100. + 50. * (_n + 3) * (_n + 3) / std::sqrt(1. * lambda));
    return std::min(maxit * lambda, _mfev - _fev);
}

// ---------------------------------------------------------------------------------------
// BIPOP / NBIPOP -- cma/bipop_cmaes.cpp
// ---------------------------------------------------------------------------------------
BiPopCmaes::BiPopCmaes(BaseCmaes *base, int mfev, bool print, double sigma0, int maxlargeruns, bool nbipop,
        double ksigmadec, double kbudget) {
    _sigmadef = sigma0;
    _mfev = mfev;
    _print = print;
    _base = base;
    _maxlargeruns = maxlargeruns;
    _nbipop = nbipop;
    _ksigmadec = ksigmadec;
    _kbudget = kbudget;
    _n = _fev = _lambdadef = _largelambda = _smalllambda = _largebudget = _smallbudget = 0;
    _largerestarts = _smallrestarts = _it = 0;
    _bestregime = 1;
    _fx = _largesigma = _smallsigma = _fxbest = 0.;
}

// bipop_cmaes.cpp:61-107
void BiPopCmaes::init(const multivariate_problem &f, const double *guess) {
    if (f._hasc || f._hasbbc) {
        std::cerr << "Warning [BIPOP-CMAES]: problem constraints will be ignored." << std::endl;
    }
    _f = f;
    _n = f._n;
    _lower = std::vector<double>(f._lower, f._lower + _n);
    _upper = std::vector<double>(f._upper, f._upper + _n);
    _guess = std::vector<double>(guess, guess + _n);
    _fev = 0;
    _lambdadef = 4 + static_cast<int>(3. * std::log(1. * _n));
    const int maxfev = computeMaxEvaluations(_lambdadef);
    _base->setParams(_lambdadef, _sigmadef, maxfev);
    const multivariate_solution sol = _base->optimize(_f, &_guess[0]);
    _fx = _base->objective().evalOne(&sol._sol[0]);
    _fev += sol._fev + 1;
    _largebudget = _smallbudget = 0;
    _largerestarts = _smallrestarts = 0;
    _bestregime = 1;
    _it = 0;
    _fxbest = _fx;
    _xbest = sol._sol;
    _x0 = std::vector<double>(_n);
    if (_print) {
        _table.setWidth( { 5, 5, 5, 5, 10, 10, 10, 5, 25, 25, 25 });
        _table.printRow( { "run", "regime", "run1", "run2", "budget1", "budget2", "fev", "pop", "sigma", "f*", "best f*" });
        _table.printRow( { RowPrinter::num(_it), "0", RowPrinter::num(_largerestarts), RowPrinter::num(_smallrestarts),
                RowPrinter::num(_largebudget), RowPrinter::num(_smallbudget), RowPrinter::num(_fev),
                RowPrinter::num(_lambdadef), RowPrinter::num(_sigmadef), RowPrinter::num(_fx), RowPrinter::num(_fxbest) });
    }
}

// bipop_cmaes.cpp:109-164
void BiPopCmaes::iterate() {
    for (int i = 0; i < _n; i++) _x0[i] = Random::get(_lower[i], _upper[i]);
    int regime;
    if (_nbipop) {
        if (_bestregime == 1) regime = (_largebudget <= _smallbudget * _kbudget) ? 1 : 2;
        else regime = (_smallbudget <= _kbudget * _largebudget) ? 2 : 1;
    } else {
        regime = (_largebudget <= _smallbudget) ? 1 : 2;
    }
    if (regime == 1) runFirstRegime();
    else runSecondRegime();
    _it++;
    if (_print) {
        const int lam = regime == 1 ? _largelambda : _smalllambda;
        const double sg = regime == 1 ? _largesigma : _smallsigma;
        _table.printRow( { RowPrinter::num(_it), RowPrinter::num(regime), RowPrinter::num(_largerestarts),
                RowPrinter::num(_smallrestarts), RowPrinter::num(_largebudget), RowPrinter::num(_smallbudget),
                RowPrinter::num(_fev), RowPrinter::num(lam), RowPrinter::num(sg), RowPrinter::num(_fx),
                RowPrinter::num(_fxbest) });
    }
}

multivariate_solution BiPopCmaes::solution() {
    return {_xbest, _fev, false};
}

// bipop_cmaes.cpp:170-189
multivariate_solution BiPopCmaes::optimize(const multivariate_problem &f, const double *guess) {
    init(f, guess);
    while (true) {
        iterate();
        if (_largerestarts >= _maxlargeruns) {
            if (_print) {
                std::cerr << "Warning [BIPOP-CMAES]: reached the maximum number of large population restarts." << std::endl;
            }
            break;
        }
        if (_fev >= _mfev) break;
    }
    return {_xbest, _fev, false};
}

int BiPopCmaes::computeMaxEvaluations(const int lambda) {
    const int maxit = static_cast<int>(100. + 50. * (_n + 3) * (_n + 3) / std::sqrt(1. * lambda));
    return std::min(maxit * lambda, _mfev - _fev);
}

// bipop_cmaes.cpp:204-236
void BiPopCmaes::runFirstRegime() {
    _largelambda = static_cast<int>(_lambdadef * std::pow(2, _largerestarts + 1));
    if (_nbipop) {
        _largesigma = _sigmadef * std::pow(1. / _ksigmadec, _largerestarts + 1);
        _largesigma = std::max(_largesigma, 0.01 * _sigmadef);
    } else {
        _largesigma = _sigmadef;
    }
    const int maxfev = computeMaxEvaluations(_largelambda);
    _base->setParams(_largelambda, _largesigma, maxfev);
    const multivariate_solution sol = _base->optimize(_f, &_x0[0]);
    _fx = _base->objective().evalOne(&sol._sol[0]);
    _fev += sol._fev + 1;
    if (_fx < _fxbest) {
        _fxbest = _fx;
        _xbest = sol._sol;
        _bestregime = 1;
    }
    _largebudget += sol._fev;
    _largerestarts++;
}

// bipop_cmaes.cpp:238-267
void BiPopCmaes::runSecondRegime() {
    const double u = Random::get(0., 1.);
    _smalllambda = static_cast<int>(_lambdadef * std::pow((0.5 * _largelambda) / _lambdadef, u * u));
    _smallsigma = _sigmadef * std::pow(10., -2. * Random::get(0., 1.));
    int maxfev = computeMaxEvaluations(_smalllambda);
    maxfev = std::min(maxfev, _largebudget >> 1);
    _base->setParams(_smalllambda, _smallsigma, maxfev);
    const multivariate_solution sol = _base->optimize(_f, &_x0[0]);
    _fx = _base->objective().evalOne(&sol._sol[0]);
    _fev += sol._fev + 1;
    if (_fx < _fxbest) {
        _fxbest = _fx;
        _xbest = sol._sol;
        _bestregime = 2;
    }
    _smallbudget += sol._fev;
    _smallrestarts++;
}
