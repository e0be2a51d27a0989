"""Timing of the tridiagonal divide & conquer alone (CUDA events around one stedc_device call, after a warm-up call).
    python tools/stedc_time.py [n ...]"""
import os, sys
import numpy as np
import scipy.linalg
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bboptpy_b200 import _cabi
ctx = _cabi.Context(0)
for n in [int(a) for a in sys.argv[1:]] or [1000]:
    rng = np.random.default_rng(n)
    M = rng.standard_normal((n, n))
    A = np.eye(n) + 1e-2 * (M + M.T) + np.diag(rng.uniform(0, 1, n))   # the spectrum of a CMA covariance early in a run
    T = scipy.linalg.hessenberg(A)
    d, e = np.diag(T).copy(), np.diag(T, -1).copy()
    best = 1e9
    for _ in range(5):
        w, Z, ms = ctx.stedc(d, e)
        best = min(best, ms)
    Tm = np.diag(d) + np.diag(e, 1) + np.diag(e, -1)
    res = min(np.abs(Z @ Tm @ Z.T - np.diag(w)).max(), np.abs(Z.T @ Tm @ Z - np.diag(w)).max())
    orth = np.abs(Z @ Z.T - np.eye(n)).max()
    print(f"n = {n}: stedc {best:.3f} ms (best of 5), residual {res:.2e}, orthogonality {orth:.2e}", flush=True)
