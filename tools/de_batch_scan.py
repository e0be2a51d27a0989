"""Batched-runs DE throughput scan: R independent runs per launch (tools; not part of the bench contract)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from bboptpy_b200 import _cabi
ctx = _cabi.Context(0)
n, npop = 1000, 1000
lo, hi = -5.12 * np.ones(n), 5.12 * np.ones(n)
obj = _cabi.Objective(ctx, n, builtin=_cabi.OBJ_RASTRIGIN)
Rs = [int(a) for a in sys.argv[1:]] or [1, 4, 16, 64]
for kind in (0, 1):
    for R in Rs:
        runs = [_cabi.De(ctx, kind, n, npop, lo, hi, mfev=10**7, seed=7 + r) for r in range(R)]
        for g in runs:
            g.init(obj)
        for _ in range(5):
            reps = _cabi.De.generation_batch(runs, obj)
        gens = 20
        ctx.sync()
        t0 = time.perf_counter()
        for _ in range(gens):
            _cabi.De.generation_batch(runs, obj, want_report=False)
        ctx.sync()
        dt = time.perf_counter() - t0
        reps = _cabi.De.generation_batch(runs, obj)
        ev = R * npop * gens / dt
        print("kind %d R %3d: %.1f us/gen  %.2f M evals/s  %.1f GB/s algorithmic (%.3f of 6460)  rounds %d" %
              (kind, R, dt / gens * 1e6, ev / 1e6, ev * 40000 / 1e9, ev * 40000 / 1e9 / 6460, reps[0].rounds), flush=True)
        del runs
