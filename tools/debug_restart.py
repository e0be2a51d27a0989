import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bboptpy_b200 as pkg
from oracle import oracle as O
P = O.port()
n = 6
lo, hi = -5.12 * np.ones(n), 5.12 * np.ones(n)
for rk, drv in ((0, "IPopCMAES"), (1, "BiPopCMAES")):
    P.seed(5)
    guess = P.uniform(-5, 5, n)
    base_o = P.cma(O.SEP, 1000, 1e-8, 10, adjustlr=False)
    x_o, fev_o, trace = P.restart_optimize(rk, base_o, 30000, O.RASTRIGIN, lo, hi, guess)
    print("oracle", drv, fev_o)
    for r in trace: print("   lam %d sigma %.17g fx %.17g fev %d regime %d" % (r[0], r[1], r[2], r[3], r[4]))
    pkg.seed(5)
    g2 = np.array([pkg.uniform(-5, 5) for _ in range(n)])
    base = pkg.SepCMAES(mfev=1000, tol=1e-8, np=10, adjustlr=False)
    base.set_draw_replay_global()
    alg = getattr(pkg, drv)(base, mfev=30000, print=True)
    sol = alg.optimize(pkg.Rastrigin, lo, hi, guess)
    print("gpu", sol.n_evals)
