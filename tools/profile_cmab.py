"""One step launch of the batched small-n engine for ncu (and plain timing):
    python tools/profile_cmab.py [runs] [n] [lambda] [gens]
Starts `runs` runs at (n, lambda) on Rastrigin, one warm-up launch of `gens` generations, then timed launches."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bboptpy_b200 import _cabi  # noqa: E402

runs = int(sys.argv[1]) if len(sys.argv) > 1 else 1184
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
lam = int(sys.argv[3]) if len(sys.argv) > 3 else 15
gens = int(sys.argv[4]) if len(sys.argv) > 4 else 40
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 3
ctx = _cabi.Context(0)
obj = _cabi.Objective(ctx, n, builtin=_cabi.OBJ_RASTRIGIN)
batch = _cabi.CmaBatch(ctx, n, runs)
params = [_cabi.CmaBatch.run_params(n, lam, 10 ** 9, 2.0, 1e-12, seed=r) for r in range(runs)]
batch.start(range(runs), params, np.random.default_rng(0).uniform(-5, 5, (runs, n)))
ids = list(range(runs))
batch.step(obj, ids, quantum=gens * lam, max_gens=gens)
ctx.sync()
for rep in range(reps):
    ctx.timer_start()
    batch.step(obj, ids, quantum=gens * lam, max_gens=gens)
    ms = ctx.timer_stop()
    st = batch.fetch()
    g = sum(s.generations_this_step for s in st)
    e = sum(s.eig_count for s in st)
    print(f"runs {runs} n {n} lambda {lam}: {g} run-generations in {ms:.3f} ms = {g / ms / 1e3:.3f} M run-gens/s, "
          f"{ms * 1e3 / gens:.1f} us per generation per CTA wave, eig so far {e}, QL iterations/eig {sum(s.ql_iters for s in st) / max(e, 1):.1f}", flush=True)
