// Restart drivers over the device CMA-ES classes: IPOP / NIPOP and BIPOP / NBIPOP.
// Pure host control logic, same constructor signatures and behaviour as the reference
// (/root/reference/src/multivariate/cma/ipop_cmaes.h:52-84, bipop_cmaes.h:46-85).
#ifndef BBOPT_HOST_RESTART_H_
#define BBOPT_HOST_RESTART_H_

#include <string>
#include <vector>

#include "cma.h"

// console table of the `print=True` mode (same column layout as the reference's Tabular)
class RowPrinter {
    std::vector<int> _w;
    int _rows = 0;
public:
    void setWidth(const std::vector<int> &w) { _w = w; _rows = 0; }
    void printRow(const std::vector<std::string> &cells);
    static std::string num(double v);
    static std::string num(int v) { return std::to_string(v); }
};

class IPopCmaes: public MultivariateOptimizer {
protected:
    bool _print, _nipop, _boundlambda;
    int _n, _mfev, _it, _fev, _lambdadef, _lambdamax, _lambda;
    double _sigmadef, _sigma, _ksigmadec, _fbest, _fx;
    multivariate_problem _f;
    BaseCmaes *_base;
    RowPrinter _table;
    std::vector<double> _lower, _upper, _guess, _xbest, _x0;
public:
    IPopCmaes(BaseCmaes *base, int mfev, bool print = false, double sigma0 = 2., bool nipop = true,
            double ksigmadec = 1.6, bool boundlambda = true);
    void init(const multivariate_problem &f, const double *guess);
    void iterate();
    multivariate_solution solution();
    multivariate_solution optimize(const multivariate_problem &f, const double *guess);
    int lambda() const { return _lambda; }
    double sigma() const { return _sigma; }
    int evaluations() const { return _fev; }
private:
    int computeMaxEvaluations(const int lambda);
};

class BiPopCmaes: public MultivariateOptimizer {
protected:
    bool _print, _nbipop;
    int _n, _mfev, _fev, _maxlargeruns, _lambdadef, _largelambda, _smalllambda, _largebudget, _smallbudget,
            _largerestarts, _smallrestarts, _it, _bestregime;
    double _fx, _sigmadef, _largesigma, _smallsigma, _fxbest, _ksigmadec, _kbudget;
    multivariate_problem _f;
    BaseCmaes *_base;
    RowPrinter _table;
    std::vector<double> _lower, _upper, _guess, _x0, _xbest;
public:
    BiPopCmaes(BaseCmaes *base, int mfev, bool print = false, double sigma0 = 2., int maxlargeruns = 9,
            bool nbipop = true, double ksigmadec = 1.6, double kbudget = 2.);
    void init(const multivariate_problem &f, const double *guess);
    void iterate();
    multivariate_solution solution();
    multivariate_solution optimize(const multivariate_problem &f, const double *guess);
    int evaluations() const { return _fev; }
    int largeRestarts() const { return _largerestarts; }
    int smallRestarts() const { return _smallrestarts; }
private:
    int computeMaxEvaluations(const int lambda);
    void runFirstRegime();
    void runSecondRegime();
};

#endif
