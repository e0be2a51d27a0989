"""The eigendecomposition of one CMA handle (bbgpu_cma_eig, our own solver at n >= 256) on a covariance matrix with a graded
spectrum: timing with CUDA events through the profile counters, or a target for ncu:
    python tools/eig_probe_handle.py [n] [reps]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bboptpy_b200 import _cabi  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
lam = 16
mu = lam // 2
w = np.log(mu + 0.5) - np.log(np.arange(1, mu + 1))
w /= w.sum()
mueff = 1.0 / (w ** 2).sum()
consts = dict(cc=4.0 / (n + 4), cs=(mueff + 2) / (n + mueff + 3), c1=2.0 / ((n + 1.3) ** 2 + mueff), cmu=1e-3, mueff=mueff,
              chi=np.sqrt(n) * (1 - 1 / (4 * n) + 1 / (21 * n * n)))
ctx = _cabi.Context(0)
g = _cabi.Cma(ctx, 0, n, lam, w, consts, seed=1)
g.reset(np.ones(n))
rng = np.random.default_rng(n)
Q, _ = np.linalg.qr(rng.standard_normal((n, n)))
spectrum = np.sort(np.exp(rng.uniform(0.0, np.log(1e3), n)))
C = (Q * spectrum) @ Q.T
C = 0.5 * (C + C.T)
for r in range(reps):
    g.set(_cabi.F_C, C)
    t0 = time.perf_counter()
    g.eig()
    rep = g.fetch()
    dt = time.perf_counter() - t0
    assert rep.eig_info == 0
    print(f"n = {n}: eig + fetch {dt * 1e3:.3f} ms (wall)", flush=True)
D = g.get(_cabi.F_D)
print("eigenvalues vs numpy: max rel diff", float(np.abs(D ** 2 / np.linalg.eigvalsh(C) - 1).max()))
