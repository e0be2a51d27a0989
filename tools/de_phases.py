import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from bboptpy_b200 import _cabi
ctx = _cabi.Context(0)
n, npop = 1000, 1000
lo, hi = -5.12 * np.ones(n), 5.12 * np.ones(n)
for kind in (0, 1):
    g = _cabi.De(ctx, kind, n, npop, lo, hi, mfev=10**7, seed=7)
    obj = _cabi.Objective(ctx, n, builtin=_cabi.OBJ_RASTRIGIN)
    g.init(obj)
    for gen in range(6):
        rep = g.generation(obj)
        t = g.phase_times().astype(np.int64)
        t0 = t[0]
        rounds = rep.rounds
        marks = [("sort+params", 1)] + [("r%d" % r, 2 + r) for r in range(min(rounds, 12))] + [("finalize", 16), ("sort", 17), ("end", 18)]
        prev = t0
        parts = []
        for name, idx in marks:
            parts.append("%s=%.1f" % (name, (t[idx] - prev) / 1e3))
            prev = t[idx]
        print("kind", kind, "gen", gen, "rounds", rounds, "total_us %.1f |" % ((t[18] - t0) / 1e3), " ".join(parts))
