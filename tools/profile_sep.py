"""A few generations of BASELINE config 3 (SepCmaes n = 10^4, lambda = 4096, Ellipsoid) through the C ABI: the short command
ncu wraps.  Prints the CUDA-event time per generation when run without a profiler."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from bboptpy_b200 import _cabi
from bboptpy_b200.strategy import cma_constants, update_sigma
gens = int(sys.argv[1]) if len(sys.argv) > 1 else 4
n, lam = 10000, 4096
ctx = _cabi.Context(0)
w, c = cma_constants(2, n, lam)
g = _cabi.Cma(ctx, _cabi.CMA_SEP, n, lam, w, c, seed=1)
g.reset(np.random.default_rng(0).uniform(-10, 10, n))
obj = _cabi.Objective(ctx, n, builtin=_cabi.OBJ_ELLIPSOID)
sigma, fev = 2.0, 0
for rep_ in range(2):
    ctx.sync()
    t0 = time.perf_counter()
    for gen in range(gens):
        fev += lam
        g.generation(obj, sigma, gen, c["ik"], fev)
    rep = g.fetch()
    dt = time.perf_counter() - t0
print("ok", rep.ybw[0], "wall per generation %.1f us" % (dt / gens * 1e6), ctx.launches)
