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
// IPOP / NIPOP -- cma/ipop_cmaes.cpp.  The bookkeeping lives in IPopSchedule (schedule.h); this class runs each
// requested base run on the one-handle device engine, as the reference runs it on its BaseCmaes.
// ---------------------------------------------------------------------------------------
static const uniform_fn global_uniform = [](double a, double b) { return Random::get(a, b); };

IPopCmaes::IPopCmaes(BaseCmaes *base, int mfev, bool print, double sigma0, bool nipop, double ksigmadec,
        bool boundlambda) {
    _base = base;
    _print = print;
    _s.configure(mfev, sigma0, nipop, ksigmadec, boundlambda);
}

// setParams -> optimize -> the extra f(x) (ipop_cmaes.cpp:90-99, 143-158)
void IPopCmaes::runBase(const base_run_request &req, const double *x0) {
    _base->setParams(req.lambda, req.sigma, req.mfev);
    const multivariate_solution sol = _base->optimize(_f, x0);
    const double fx = _base->objective().evalOne(&sol._sol[0]);
    if (_s.finish(sol._fev, fx)) _xbest = sol._sol;
    if (_print) {
        if (req.regime == 0) {
            _table.setWidth( { 5, 10, 5, 25, 25, 25 });
            _table.printRow( { "run", "budget", "pop", "sigma", "f*", "best f*" });
        }
        _table.printRow( { RowPrinter::num(_s._it), RowPrinter::num(_s._fev), RowPrinter::num(req.lambda),
                RowPrinter::num(req.sigma), RowPrinter::num(_s._fx), RowPrinter::num(_s._fbest) });
    }
}

// ipop_cmaes.cpp:65-110
void IPopCmaes::init(const multivariate_problem &f, const double *guess) {
    if (f._hasc || f._hasbbc) {
        std::cerr << "Warning [IPOP-CMAES]: problem constraints will be ignored." << std::endl;
    }
    _f = f;
    const int n = f._n;
    _lower = std::vector<double>(f._lower, f._lower + n);
    _upper = std::vector<double>(f._upper, f._upper + n);
    _guess = std::vector<double>(guess, guess + n);
    _x0 = std::vector<double>(n);
    runBase(_s.first(n), &_guess[0]);
}

// ipop_cmaes.cpp:112-162
void IPopCmaes::iterate() {
    const base_run_request req = _s.next(&_lower[0], &_upper[0], _x0, global_uniform);
    runBase(req, &_x0[0]);
}

multivariate_solution IPopCmaes::solution() {
    return {_xbest, _s._fev, false};
}

multivariate_solution IPopCmaes::optimize(const multivariate_problem &f, const double *guess) {
    init(f, guess);
    while (!_s.done()) iterate();
    return {_xbest, _s._fev, false};
}

// ---------------------------------------------------------------------------------------
// BIPOP / NBIPOP -- cma/bipop_cmaes.cpp over BiPopSchedule
// ---------------------------------------------------------------------------------------
BiPopCmaes::BiPopCmaes(BaseCmaes *base, int mfev, bool print, double sigma0, int maxlargeruns, bool nbipop,
        double ksigmadec, double kbudget) {
    _print = print;
    _base = base;
    _s.configure(mfev, sigma0, maxlargeruns, nbipop, ksigmadec, kbudget);
}

void BiPopCmaes::printRow() {
    _table.printRow( { RowPrinter::num(_s._it), RowPrinter::num(_s._regime), RowPrinter::num(_s._largerestarts),
            RowPrinter::num(_s._smallrestarts), RowPrinter::num(_s._largebudget), RowPrinter::num(_s._smallbudget),
            RowPrinter::num(_s._fev), RowPrinter::num(_s.currentLambda()), RowPrinter::num(_s.currentSigma()),
            RowPrinter::num(_s._fx), RowPrinter::num(_s._fxbest) });
}

// setParams -> optimize -> the extra f(x) (bipop_cmaes.cpp:80-92, 217-235, 252-266)
void BiPopCmaes::runBase(const base_run_request &req, const double *x0) {
    _base->setParams(req.lambda, req.sigma, req.mfev);
    const multivariate_solution sol = _base->optimize(_f, x0);
    const double fx = _base->objective().evalOne(&sol._sol[0]);
    if (_s.finish(sol._fev, fx)) _xbest = sol._sol;
    if (_print) {
        if (req.regime == 0) {
            _table.setWidth( { 5, 5, 5, 5, 10, 10, 10, 5, 25, 25, 25 });
            _table.printRow( { "run", "regime", "run1", "run2", "budget1", "budget2", "fev", "pop", "sigma", "f*", "best f*" });
        }
        printRow();
    }
}

// bipop_cmaes.cpp:61-107
void BiPopCmaes::init(const multivariate_problem &f, const double *guess) {
    if (f._hasc || f._hasbbc) {
        std::cerr << "Warning [BIPOP-CMAES]: problem constraints will be ignored." << std::endl;
    }
    _f = f;
    const int n = f._n;
    _lower = std::vector<double>(f._lower, f._lower + n);
    _upper = std::vector<double>(f._upper, f._upper + n);
    _guess = std::vector<double>(guess, guess + n);
    _x0 = std::vector<double>(n);
    runBase(_s.first(n), &_guess[0]);
}

// bipop_cmaes.cpp:109-164
void BiPopCmaes::iterate() {
    const base_run_request req = _s.next(&_lower[0], &_upper[0], _x0, global_uniform);
    runBase(req, &_x0[0]);
}

multivariate_solution BiPopCmaes::solution() {
    return {_xbest, _s._fev, false};
}

// bipop_cmaes.cpp:170-189
multivariate_solution BiPopCmaes::optimize(const multivariate_problem &f, const double *guess) {
    init(f, guess);
    while (true) {
        iterate();
        if (_s.hitLargeRunLimit()) {
            if (_print) {
                std::cerr << "Warning [BIPOP-CMAES]: reached the maximum number of large population restarts." << std::endl;
            }
            break;
        }
        if (_s._fev >= _s._mfev) break;
    }
    return {_xbest, _s._fev, false};
}
