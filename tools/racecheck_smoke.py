"""Small instances of every pipeline with shared-memory hand-offs (TMA + mbarrier GEMMs, the DE round kernel, the divide &
conquer setup scan, the Cholesky panels, the LM / Sep sampling kernels), sized for compute-sanitizer:
    compute-sanitizer --tool racecheck python tools/racecheck_smoke.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from bboptpy_b200 import _cabi
from bboptpy_b200.strategy import cma_constants

ctx = _cabi.Context(0)
which = sys.argv[1:] or ["cma", "eig", "chol", "sep", "lm", "de", "rank"]
if "cma" in which:
    n, lam = 130, 256
    for kind in (0, 1):
        w, c = cma_constants(kind, n, lam)
        g = _cabi.Cma(ctx, kind, n, lam, w, c, seed=1)
        g.reset(np.linspace(-3, 3, n))
        obj = _cabi.Objective(ctx, n, builtin=_cabi.OBJ_ROSENBROCK)
        g.generation(obj, 1.5, 0, c["ik"], lam)
        g.fetch()
    print("cma ok", flush=True)
if "eig" in which:
    n = 300
    w, c = cma_constants(0, n, 16)
    g = _cabi.Cma(ctx, 0, n, 16, w, c, seed=1)
    g.reset(np.ones(n))
    M = np.random.default_rng(0).standard_normal((n, n))
    g.set(_cabi.F_C, np.eye(n) + 1e-2 * (M + M.T))
    g.eig()
    assert g.fetch().eig_info == 0
    print("eig ok", flush=True)
if "chol" in which:
    n, lam = 130, 64
    w, c = cma_constants(3, n, lam)
    g = _cabi.Cma(ctx, _cabi.CMA_CHOLESKY, n, lam, w, c, seed=1)
    g.reset(np.linspace(-3, 3, n))
    obj = _cabi.Objective(ctx, n, builtin=_cabi.OBJ_ROSENBROCK)
    for gen in range(2):
        g.generation(obj, 1.5, gen, c["ik"], lam * (gen + 1))
    assert g.fetch().eig_info == 0
    print("chol ok", flush=True)
if "sep" in which:
    n, lam = 3000, 64
    w, c = cma_constants(2, n, lam)
    g = _cabi.Cma(ctx, _cabi.CMA_SEP, n, lam, w, c, seed=1)
    g.reset(np.linspace(-3, 3, n))
    obj = _cabi.Objective(ctx, n, builtin=_cabi.OBJ_ELLIPSOID)
    g.generation(obj, 1.5, 0, c["ik"], lam)
    g.fetch()
    print("sep ok", flush=True)
if "lm" in which:
    import bboptpy_b200 as pkg
    alg = pkg.LmCMAES(mfev=10 ** 6, tol=1e-12, np=16, memory=5)
    alg.initialize(pkg.Rastrigin, -5 * np.ones(64), 5 * np.ones(64), np.ones(64))
    for _ in range(8):
        alg.iterate()
    print("lm ok", flush=True)
if "de" in which:
    import bboptpy_b200 as pkg
    for cls in (pkg.JADE, pkg.SHADE):
        alg = cls(10 ** 6, 64, 1e-12)
        alg.initialize(pkg.Rastrigin, -5 * np.ones(50), 5 * np.ones(50), np.ones(50))
        for _ in range(2):
            alg.iterate()
    print("de ok", flush=True)
if "rank" in which:
    f = np.random.default_rng(1).standard_normal(9000)
    assert np.array_equal(ctx.rank(f), np.argsort(f, kind="stable"))
    print("rank ok", flush=True)
