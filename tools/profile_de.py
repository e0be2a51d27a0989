"""ncu target: JADE n=1000 np=1000 Rastrigin, R runs per launch (argv[1], default 1), 8 generations."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from bboptpy_b200 import _cabi
R = int(sys.argv[1]) if len(sys.argv) > 1 else 1
ctx = _cabi.Context(0)
n, npop = 1000, 1000
lo, hi = -5.12 * np.ones(n), 5.12 * np.ones(n)
obj = _cabi.Objective(ctx, n, builtin=_cabi.OBJ_RASTRIGIN)
runs = [_cabi.De(ctx, 0, n, npop, lo, hi, mfev=10**7, seed=7 + r) for r in range(R)]
for g in runs:
    g.init(obj)
for _ in range(8):
    reps = _cabi.De.generation_batch(runs, obj)
print("rounds", reps[0].rounds, "fbest", reps[0].fbest)
