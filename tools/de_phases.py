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
        if gen == 5:
            for r in range(min(rounds, 8)):
                base = t[1] if r == 0 else t[32 + 4 * (r - 1) + 3]
                q = t[32 + 4 * r: 36 + 4 * r]
                print("   round %d: scan %.1f  signatures %.1f  builds %.1f  grid.sync %.1f" %
                      (r, (q[0] - base) / 1e3, (q[1] - q[0]) / 1e3, (q[2] - q[1]) / 1e3, (q[3] - q[2]) / 1e3))
