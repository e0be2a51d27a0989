// TEST INFRASTRUCTURE: drives the host classes (bboptpy_b200/host) against the mock C ABI.
#include <cstdio>
#include <cstring>
#include <iostream>

#include "cma.h"
#include "de.h"
#include "restart.h"
#include "batch.h"
#include "sansde.h"
#include "ssde.h"

int main(int argc, char **argv) {
    const int n = 6;
    std::vector<double> lo(n, -5.12), hi(n, 5.12), guess(n, 1.5);
    device_objective dev;
    dev._builtin = BBGPU_OBJ_RASTRIGIN;
    multivariate_problem prob(dev, n, lo.data(), hi.data());
    const bool print = argc > 1 && !std::strcmp(argv[1], "print");
    int total = 0;
    for (int variant = 0; variant < 3; variant++) {
        BaseCmaes *base;
        if (variant == 0) base = new Cmaes(1000, 1e-8, 10);
        else if (variant == 1) base = new ActiveCmaes(1000, 1e-8, 10);
        else base = new SepCmaes(1000, 1e-8, 10);
        IPopCmaes ipop(base, 20000, print);
        multivariate_solution s1 = ipop.optimize(prob, guess.data());
        BiPopCmaes bipop(base, 20000, print);
        multivariate_solution s2 = bipop.optimize(prob, guess.data());
        std::printf("variant %d ipop fev %d bipop fev %d (large %d small %d)\n", variant, s1._fev, s2._fev, bipop.largeRestarts(), bipop.smallRestarts());
        total += s1._fev + s2._fev;
        delete base;
    }
    multivariate f = [n](const double *x) { double s = 0; for (int i = 0; i < n; i++) s += x[i] * x[i]; return s; };
    multivariate_problem hostprob(f, n, lo.data(), hi.data());
    ActiveCmaes a(2000, 1e-4, 20);
    multivariate_solution s3 = a.optimize(hostprob, guess.data());
    std::printf("host-callback run: fev %d converged %d\n%s\n", s3._fev, (int) s3._converged, s3.toString().c_str());
    JadeSearch jade(5000, 20, 1e-3);
    multivariate_solution s4 = jade.optimize(prob, guess.data());
    ShadeSearch shade(5000, 20, 1e-3);
    multivariate_solution s5 = shade.optimize(prob, guess.data());
    std::printf("jade fev %d conv %d, shade fev %d conv %d\n", s4._fev, (int) s4._converged, s5._fev, (int) s5._converged);
    {   // SaNSDE / SSDE: budgets and the population-reduction schedule (de/ssde.cpp:354-360) with the stand-in's np evaluations per generation
        SaNSDESearch sde(3000, 20, 0.0);
        multivariate_solution s6 = sde.optimize(prob, guess.data());
        const int npinit = 40, npmin = 4, mfev = 2000;
        SSDESearch ssde(mfev, npinit, 0.0, 1000000, npmin, 0.11, 10, false, true);
        multivariate_solution s7 = ssde.optimize(prob, guess.data());
        int np = npinit, fev = npinit, gens = 0;
        while (fev < mfev) {
            fev += np;
            gens++;
            const int npnew = static_cast<int>(npinit + (npmin - npinit) * (1. * fev) / mfev);
            if (np > npnew) np = npnew;
        }
        std::printf("sansde fev %d, ssde fev %d (expected %d) final population %d (expected %d) after %d generations\n", s6._fev, s7._fev, fev,
                ssde.population(), np, gens);
        if (s7._fev != fev || ssde.population() != np || s6._fev < 3000) return 3;
    }
    for (int kind = 0; kind < 2; kind++) {
        const int runs = 37;
        RestartBatch batch(kind == 0 ? RestartBatch::IPOP : RestartBatch::BIPOP, runs, 20000, 1e-8, 2., 0.25);
        std::vector<double> guesses((size_t) runs * n, 1.0);
        std::vector<multivariate_solution> sols = batch.optimize(prob, guesses.data());
        long long fev = 0;
        for (const multivariate_solution &s : sols) fev += s._fev;
        std::printf("batch kind %d: %zu runs, %lld base runs, %lld evaluations (sum %lld), %lld launches\n", kind, sols.size(),
                batch.stats().base_runs, batch.stats().evaluations, fev, batch.stats().launches);
        if (fev != batch.stats().evaluations || (int) sols.size() != runs) return 2;
    }
    return total > 0 ? 0 : 1;
}
