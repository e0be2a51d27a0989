"""ncu launch-list target: a few generations of every hot path at its BASELINE configuration.
  A  Cmaes n=1000 lambda=4096 Rosenbrock      (3 generations + 1 eigendecomposition)
  A' ActiveCmaes, same shape                  (2 generations)
  S  SepCmaes n=10^4 lambda=4096 Ellipsoid    (3 generations)
  D  JADE / L-SHADE n=1000 np=1000 Rastrigin  (3 generations each, single run; 2 generations of a 16-run batch)
Usage: python tools/profile_all.py [A|S|D|all]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
from bboptpy_b200 import _cabi  # noqa: E402
from bboptpy_b200.strategy import cma_constants, update_sigma  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "all"
ctx = _cabi.Context(0)


def cma(kind, n, lam, objective, gens, eig):
    w, c = cma_constants(kind, n, lam)
    g = _cabi.Cma(ctx, kind, n, lam, w, c, seed=3)
    g.reset(np.random.default_rng(1).uniform(-10, 10, n))
    obj = _cabi.Objective(ctx, n, builtin=objective)
    sigma, fev = 2.0, 0
    for gen in range(gens):
        fev += lam
        g.generation(obj, sigma, gen, c["ik"], fev)
        rep = g.fetch()
        sigma = update_sigma(sigma, rep.pslen, c, rep.ybw[0], rep.f_kth)
        if eig and gen == 0:
            g.eig()
    return rep.ybw[0]


if what in ("A", "all"):
    print("A  cmaes", cma(_cabi.CMA_FULL, 1000, 4096, _cabi.OBJ_ROSENBROCK, 3, True))
    print("A' active", cma(_cabi.CMA_ACTIVE, 1000, 4096, _cabi.OBJ_ROSENBROCK, 2, True))
if what in ("S", "all"):
    print("S  sep", cma(_cabi.CMA_SEP, 10000, 4096, _cabi.OBJ_ELLIPSOID, 3, False))
if what in ("D", "all"):
    n, npop = 1000, 1000
    lo, hi = -5.12 * np.ones(n), 5.12 * np.ones(n)
    obj = _cabi.Objective(ctx, n, builtin=_cabi.OBJ_RASTRIGIN)
    for kind in (0, 1):
        g = _cabi.De(ctx, kind, n, npop, lo, hi, mfev=10 ** 7, seed=7)
        g.init(obj)
        for _ in range(3):
            rep = g.generation(obj)
        print("D  de kind", kind, rep.fbest, "rounds", rep.rounds)
    runs = [_cabi.De(ctx, 0, n, npop, lo, hi, mfev=10 ** 7, seed=7 + r) for r in range(16)]
    for g in runs:
        g.init(obj)
    for _ in range(2):
        reps = _cabi.De.generation_batch(runs, obj)
    print("D  de batch16 rounds", reps[0].rounds)
