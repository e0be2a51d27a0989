"""Restart-driver replay side by side: the oracle's per-restart trace against the engine's printed table.
    python tools/debug_restart.py [sep|chol] [budget]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bboptpy_b200 as pkg
from oracle import oracle as O
P = O.port()
which = sys.argv[1] if len(sys.argv) > 1 else "sep"
budget = int(sys.argv[2]) if len(sys.argv) > 2 else 30000
n = 6
lo, hi = -5.12 * np.ones(n), 5.12 * np.ones(n)
for rk, drv in ((0, "IPopCMAES"), (1, "BiPopCMAES")):
    P.seed(5)
    guess = P.uniform(-5, 5, n)
    base_o = P.cma(O.SEP, 1000, 1e-8, 10, adjustlr=False) if which == "sep" else P.cma(O.CHOLESKY, 1000, 1e-8, 10, stol=1e-8)
    x_o, fev_o, trace = P.restart_optimize(rk, base_o, budget, O.RASTRIGIN, lo, hi, guess)
    print("oracle", drv, fev_o)
    for r in trace: print("   lam %d sigma %.17g fx %.17g fev %d regime %d" % (r[0], r[1], r[2], r[3], r[4]))
    pkg.seed(5)
    g2 = np.array([pkg.uniform(-5, 5) for _ in range(n)])
    base = pkg.SepCMAES(mfev=1000, tol=1e-8, np=10, adjustlr=False) if which == "sep" else pkg.CholeskyCMAES(mfev=1000, tol=1e-8, stol=1e-8, np=10)
    base.set_draw_replay_global()
    alg = getattr(pkg, drv)(base, mfev=budget, print=True)
    sol = alg.optimize(pkg.Rastrigin, lo, hi, guess)
    print("gpu", sol.n_evals)
