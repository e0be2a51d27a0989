"""Timing of the tridiagonalisation at a few sizes: BBGPU_SYTRD=smem selects the shared-memory kernel.
    python tools/sytrd_time.py [n ...]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bboptpy_b200 import _cabi
ctx = _cabi.Context(0)
for n in [int(a) for a in sys.argv[1:]] or [1000]:
    rng = np.random.default_rng(n)
    M = rng.standard_normal((n, n))
    A = np.eye(n) + 1e-2 * (M + M.T) + np.diag(rng.uniform(0, 1, n))
    d, e, tau, V, ms = ctx.sytrd(A, reps=5)
    T = np.diag(d) + np.diag(e, 1) + np.diag(e, -1)
    err = np.abs(np.linalg.eigvalsh(T) - np.linalg.eigvalsh(A)).max()
    print(f"n = {n}: sytrd {ms:.3f} ms (best of 5), eigenvalue error {err:.2e}, kernel = {os.environ.get('BBGPU_SYTRD', 'auto')}", flush=True)
