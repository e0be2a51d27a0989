"""Timing of the cooperative tridiagonalisation kernel (csrc/sytrd.cu) against cuSOLVER's Dsytrd (tools/eig_parts.cu):
    python tools/sytrd_probe.py [n ...]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bboptpy_b200 import _cabi  # noqa: E402

ctx = _cabi.Context(0)
for n in [int(a) for a in sys.argv[1:]] or [500, 1000, 1500]:
    rng = np.random.default_rng(n)
    M = rng.standard_normal((n, n))
    A = np.eye(n) + 1e-2 * (M + M.T)
    d, e, tau, V, ms = ctx.sytrd(A, reps=5)
    w = np.linalg.eigvalsh(np.diag(d) + np.diag(e, 1) + np.diag(e, -1))
    err = np.abs(w - np.linalg.eigvalsh(A)).max()
    print(f"n = {n}: sytrd_coop_kernel {ms:.3f} ms ({ms * 1e3 / n:.2f} us per column), eigenvalues of T vs A: max diff {err:.2e}", flush=True)
    w2, Z, ms2 = ctx.stedc(d, e)
    T = np.diag(d) + np.diag(e, 1) + np.diag(e, -1)
    print(f"n = {n}: tridiagonal divide & conquer {ms2:.3f} ms, orthogonality {np.abs(Z.T @ Z - np.eye(n)).max():.2e}, "
          f"residual {np.abs(T @ Z - Z * w2).max():.2e}", flush=True)
