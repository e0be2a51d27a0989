"""Walk one SSDE replay configuration until the device and the oracle disagree; print the per-candidate outcomes."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
from oracle import oracle as O
import bboptpy_b200 as pkg

INT_TAGS = ("iwork", "iL", "p", "q", "r", "pbest", "j0")
DBL_TAGS = ("zrank", "ub", "cauchy", "zcr", "ucr", "redraw")
P = O.port()
obj, n, npinit, gens = O.RASTRIGIN, 12, 36, 150
kw = dict(usede=True, repaircr=False, h=7, npmin=6)
mfev = npinit * (gens + 4) * 2
lo, hi = -5.12 * np.ones(n), 0.8 * 5.12 * np.ones(n)
P.seed(7)
h = P.ssde(mfev, npinit, 1e-12, **kw)
h.init(obj, lo, hi)
alg = pkg.SSDE(mfev=mfev, npinit=npinit, tol=1e-12, **kw)
alg.set_initial_population(h.getv("t_x0").reshape(npinit, n))
alg.set_draw_replay(lambda gen: {k: h.getv("t_" + k) for k in INT_TAGS + DBL_TAGS})
alg.initialize(pkg.Objective.builtin(obj), lo, hi, lo)
for g in range(1, gens + 1):
    f_before = alg.fitness()
    lcr = h.getv("LCR")
    h.iterate()
    alg.iterate()
    succ, ft = alg.trace()
    ts, tf = h.getv("t_succ"), h.getv("t_f")
    if not np.array_equal(succ, ts.astype(np.int32)) or not np.array_equal(ft, tf):
        print("gen", g, "rounds", alg.rounds, "np", len(succ))
        print("LCR before", lcr)
        print("iL", h.getv("t_iL").astype(int))
        for i in range(len(succ)):
            print(i, "dev", succ[i], ft[i], "| ref", int(ts[i]), tf[i], "| f", f_before[i], "zcr", h.getv("t_zcr")[i], "j0", h.getv("t_j0")[i])
        break
else:
    print("no mismatch")
