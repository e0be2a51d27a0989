// Batches of independent IPOP / BIPOP runs (BASELINE config 5: 4096 BIPOP-CMA-ES runs at n = 40).
//
// Every run of the batch is what the reference builds as `BiPopCmaes(new Cmaes(...), mfev, ...)` (or IPopCmaes) and
// runs with optimize(): the restart bookkeeping is one schedule object per run (schedule.h, the same code that drives
// the one-run classes of restart.h), the base CMA-ES runs are CTAs of the batched device engine (bbgpu_cmab_*,
// csrc/cmab.cu).  The reference has no batch API -- its unit of work is one optimizer object -- so this class is an
// addition, not a mirror; per run it keeps the reference's semantics (budgets, +1 evaluation per base run, regime
// choice, lambda / sigma sequences, the B / C re-init quirk of restarting the same base object).
//
// Multi-GPU: runs are independent, so rank g of a communicator owns the runs r with r % world == g and nothing is
// exchanged until bestOf() (one NCCL best-of reduce, bbgpu_comm_best_of).
#ifndef BBOPT_HOST_BATCH_H_
#define BBOPT_HOST_BATCH_H_

#include <cstdint>
#include <random>
#include <vector>

#include "device.h"
#include "multivariate.h"
#include "schedule.h"

struct restart_batch_stats {
    long long base_runs = 0;        // base CMA-ES runs started (initial runs + restarts)
    long long generations = 0;      // run-generations executed on the device
    long long evaluations = 0;      // objective evaluations, the reference's count (includes the +1 per base run)
    long long decompositions = 0;   // in-CTA eigendecompositions
    long long launches = 0;         // step launches
    double device_ms = 0.;          // summed device time of the step launches over all streams (they overlap), if timing was on
};

// Limits (bbgpu_cmab_limits): n <= 64; lambda of a base run <= 16384 (8192 for the largest n): the ranking lives in the
// CTA's shared memory.  optimize() rejects a larger n up front; a sequence whose next restart would exceed the lambda cap
// (IPOP's lambda can reach 10 n^2) stops there with its best so far, with a warning.
class RestartBatch {
public:
    enum Kind { IPOP = 0, BIPOP = 1 };

    // base Cmaes parameters: tol, eigenrate (np, sigma0, mfev of a base run come from the schedule via setParams)
    RestartBatch(Kind kind, int runs, int mfev, double tol, double sigma0, double eigenrate);
    ~RestartBatch();
    void configureIPop(bool nipop, double ksigmadec, bool boundlambda);
    void configureBiPop(int maxlargeruns, bool nbipop, double ksigmadec, double kbudget);
    // shard the runs over the ranks of `comm` (not owned): this process runs r % world == rank
    void setCommunicator(bbgpu_comm *comm) { _comm = comm; }
    void setQuantum(int evals) { _quantum = evals; }
    // number of concurrent streams the local runs are dealt to (0 = automatic: 2 from 1024 runs on)
    void setStreams(int streams) { _streams = streams; }
    void setTiming(bool on) { _timing = on; }

    // guesses: runs * n (ALL runs of the batch, also on a sharded batch).  Returns the solutions of THIS rank's runs,
    // in increasing global run id (see runIds()).
    std::vector<multivariate_solution> optimize(const multivariate_problem &f, const double *guesses);
    int totalRuns() const { return _runs; }
    const std::vector<int> &runIds() const { return _ids; }
    const std::vector<double> &bestValues() const { return _fbest; }
    const restart_batch_stats &stats() const { return _stats; }
    // best run of the whole batch (all ranks): fills x (n), returns f; owner = global run id
    double bestOf(std::vector<double> &x, int *owner);

private:
    int engineMaxLambda(int n) const; // validates n, returns the engine's lambda cap
    Kind _kind;
    int _runs, _mfev, _n = 0, _quantum = 3000, _streams = 0;
    struct Group;
    double _tol, _sigma0, _eigenrate;
    bool _nipop = true, _boundlambda = true, _nbipop = true, _timing = false;
    int _maxlargeruns = 9;
    double _ksigmadec = 1.6, _kbudget = 2.;
    bbgpu_comm *_comm = nullptr;
    std::vector<int> _ids;            // global ids of the local runs
    std::vector<double> _fbest, _xbest; // per local run
    std::vector<int> _fev;
    restart_batch_stats _stats;
};

#endif
