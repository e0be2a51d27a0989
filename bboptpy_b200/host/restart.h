// Restart drivers over the device CMA-ES classes: IPOP / NIPOP and BIPOP / NBIPOP.
// Pure host control logic, same constructor signatures and behaviour as the reference
// (/root/reference/src/multivariate/cma/ipop_cmaes.h:52-84, bipop_cmaes.h:46-85).
#ifndef BBOPT_HOST_RESTART_H_
#define BBOPT_HOST_RESTART_H_

#include <string>
#include <vector>

#include "cma.h"
#include "schedule.h"

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
    bool _print;
    IPopSchedule _s; // the restart bookkeeping (schedule.h), shared with the batched driver
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
    int lambda() const { return _s._lambda; }
    double sigma() const { return _s._sigma; }
    int evaluations() const { return _s._fev; }
private:
    void runBase(const base_run_request &req, const double *x0);
};

class BiPopCmaes: public MultivariateOptimizer {
protected:
    bool _print;
    BiPopSchedule _s;
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
    int evaluations() const { return _s._fev; }
    int largeRestarts() const { return _s._largerestarts; }
    int smallRestarts() const { return _s._smallrestarts; }
private:
    void runBase(const base_run_request &req, const double *x0);
    void printRow();
};

#endif
