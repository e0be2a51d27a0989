"""A few generations of CholeskyCmaes at the headline shape (n = 1000, lambda = 4096) through the C ABI: the short command
ncu wraps."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from bboptpy_b200 import _cabi
from bboptpy_b200.strategy import cma_constants, update_sigma
gens = int(sys.argv[1]) if len(sys.argv) > 1 else 4
n, lam = 1000, 4096
ctx = _cabi.Context(0)
w, c = cma_constants(3, n, lam)
g = _cabi.Cma(ctx, _cabi.CMA_CHOLESKY, n, lam, w, c, seed=1)
g.reset(np.random.default_rng(0).uniform(-10, 10, n))
obj = _cabi.Objective(ctx, n, builtin=_cabi.OBJ_ROSENBROCK)
sigma, fev = 2.0, 0
for gen in range(gens):
    fev += lam
    g.generation(obj, sigma, gen, c["ik"], fev)
    rep = g.fetch()
    sigma = update_sigma(sigma, rep.pslen, c, rep.ybw[0], rep.f_kth)
print("ok", rep.ybw[0], sigma, ctx.launches)
