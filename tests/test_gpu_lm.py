"""LM-CMA-ES (cma/lm_cmaes.cpp) on the device against the oracle.

T1, C ABI, teacher-forced per generation: oracle state in (mean, p_c, the live direction vectors and their scalars) ->
replayed base draws and subset starts -> X (1e-10); on the oracle's X: fitness bits and ranking bit-exact; then mean, p_c and
the rebuilt rows of V / b / d at 1e-10 (the memory bookkeeping -- algorithm 5 -- is integer logic: equal).
T2, Python API: whole trajectories under draw replay (Rademacher / Gaussian draws, 2014 / 2015 variant, a small memory that
keeps replacing vectors), the pooled-ranking step size, converged(), optimize(), IPOP with an LM base, checkpoint / resume."""
import numpy as np
import pytest

from bboptpy_b200 import _cabi
from oracle import oracle as O
from util import assert_close

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    return _cabi.Context(0)


@pytest.fixture(scope="module")
def pkg():
    import bboptpy_b200
    return bboptpy_b200


def update_set(it, t, memsize, nsteps, jarr, larr):
    """lm_cmaes.cpp:274-306 (algorithm 5) in Python: returns imin as updateDistribution uses it; mutates jarr / larr."""
    itt = it // t
    imin = 1
    if itt < memsize:
        jarr[itt] = itt
    elif memsize > 1:
        iminval = larr[jarr[1]] - larr[jarr[0]]
        for i in range(1, memsize - 1):
            val = larr[jarr[i + 1]] - larr[jarr[i]]
            if val < iminval:
                iminval, imin = val, i + 1
        if iminval >= nsteps:
            imin = 0
        jtmp = jarr[imin]
        for i in range(imin, memsize - 1):
            jarr[i] = jarr[i + 1]
        jarr[memsize - 1] = jtmp
    jcur = jarr[min(itt, memsize - 1)]
    larr[jcur] = it
    return 0 if imin == 1 else imin


CASES = [
    # n, lambda, gens, bound, kwargs
    (10, 12, 60, False, dict()),
    (10, 12, 60, False, dict(rademacher=False)),
    (30, 16, 80, True, dict(memory=5)),
    (30, 17, 80, False, dict(memory=6, usenew=False, rademacher=False)),
    (1000, 64, 12, False, dict()),
    (4096, 16, 14, False, dict(memory=6, rademacher=False)),   # sampling split over 8 column slabs, rebuild wavefront
    (12500, 8, 12, True, dict(memory=5)),                      # a row beyond the registers of a CTA: rebuild scratch in global memory
]


@pytest.mark.parametrize("n,lam,gens,bound,kw", CASES)
def test_teacher_forced_generations(ctx, port_oracle, n, lam, gens, bound, kw):
    P = port_oracle
    P.seed(42)
    lo, hi = -5.12 * np.ones(n), 5.12 * np.ones(n)
    guess = P.uniform(-5.0, 5.0, n)
    h = P.cma(O.LM, 10 ** 9, 1e-12, lam, bound=bound, **kw)
    h.init(O.RASTRIGIN, lo, hi, guess)
    memsize, nsteps, t = h.geti("memsize"), h.geti("nsteps"), h.geti("t")
    consts = {k: h.getd(k) for k in ("cc", "cs", "c1", "mueff", "chi")}
    consts.update(lm_memory=memsize, lm_rademacher=int(kw.get("rademacher", True)), lm_usenew=int(kw.get("usenew", True)))
    g = _cabi.Cma(ctx, _cabi.CMA_LM, n, lam, h.getv("weights"), consts, bound=bound, lower=lo, upper=hi)
    g.reset(guess)
    objective = _cabi.Objective(ctx, n, builtin=O.RASTRIGIN)
    ik = h.geti("ik")
    for gen in range(gens):
        s0 = h.snapshot()
        memlen, jarr, larr = s0["memlen"], s0["jarr"].astype(np.int32), s0["larr"].astype(np.int32)
        g.set(_cabi.F_XMEAN, s0["xmean"])
        g.set(_cabi.F_PC, s0["pc"])
        g.set(_cabi.F_PCMAT, s0["pcmat"])
        g.set(_cabi.F_VMAT, s0["vmat"])
        g.set(_cabi.F_LMB, s0["b"])
        g.set(_cabi.F_LMD, s0["d"])
        sigma = s0["sigma"]
        h.sample()
        z, i0 = h.getv("z"), h.getv("i0").astype(np.int32)
        g.lm_sample(sigma, gen, jarr, memlen, z=z, i0=i0)
        X = h.mat("arx")
        assert_close(g.get(_cabi.F_ARX), X, what=f"gen {gen} arx")
        g.set(_cabi.F_ARX, X)
        h.evalsort()
        g.evaluate(objective)
        g.rank(ik)
        fit_sorted, fit_idx = h.getv("fit_val"), h.getv("fit_idx").astype(np.int32)
        if len(set(fit_sorted.tolist())) == lam:  # mirrored Rademacher pairs can tie on symmetric objectives
            assert np.array_equal(g.rank_idx(), fit_idx), f"gen {gen} ranking"
        assert np.array_equal(np.sort(g.get(_cabi.F_FITNESS)), fit_sorted), f"gen {gen} fitness bits"
        if not np.array_equal(g.rank_idx(), fit_idx):
            pytest.skip("tied fitness values: std::sort's tie order is implementation defined")
        it = h.geti("it")
        due = it % t == 0
        imin = 0
        if due:
            imin = update_set(it, t, memsize, nsteps, jarr, larr)
            memlen = min(memlen + 1, memsize)
        h.update()
        assert h.geti("memlen") == memlen and np.array_equal(h.getv("jarr").astype(np.int32), jarr)
        g.lm_update(sigma, due, imin, memlen, jarr)
        g.fetch()
        assert_close(g.get(_cabi.F_XMEAN), h.getv("xmean"), what=f"gen {gen} xmean")
        assert_close(g.get(_cabi.F_PC), h.getv("pc"), what=f"gen {gen} pc")
        rows = jarr[:memlen]
        assert_close(g.get(_cabi.F_PCMAT)[rows], h.mat("pcmat")[rows], what=f"gen {gen} pcmat")
        assert_close(g.get(_cabi.F_VMAT)[rows], h.mat("vmat")[rows], what=f"gen {gen} vmat")
        assert_close(g.get(_cabi.F_LMB)[rows], h.getv("b")[rows], what=f"gen {gen} b")
        assert_close(g.get(_cabi.F_LMD)[rows], h.getv("d")[rows], what=f"gen {gen} d")
        h.history()


@pytest.mark.parametrize("n,lam,gens,kw", [(10, 12, 150, dict(rademacher=False)), (30, 16, 120, dict(memory=5, rademacher=False)),
                                           (30, 17, 120, dict(memory=6, usenew=False, rademacher=False)),
                                           (200, 24, 40, dict(rademacher=False))])
def test_trajectory_replay(pkg, port_oracle, n, lam, gens, kw):
    """T2: Gaussian base draws (continuous: tie-free), the reference's own stop, sigma from the pooled ranking."""
    P = port_oracle
    lo, hi = -10 * np.ones(n), 10 * np.ones(n)
    guess = np.random.default_rng(17).uniform(-10, 10, n)
    P.seed(42)
    h = P.cma(O.LM, 10 ** 9, 1e-12, lam, **kw)
    h.init(O.ROSENBROCK, lo, hi, guess)
    alg = pkg.LmCMAES(mfev=10 ** 9, tol=1e-12, np=lam, **kw)
    alg.set_draw_replay(42)
    alg.initialize(pkg.Rosenbrock, lo, hi, guess)
    for g in range(gens):
        h.iterate()
        alg.iterate()
        assert np.array_equal(alg.state("rank"), h.getv("fit_idx")), f"gen {g} ranking"
        assert_close(alg.state("arx"), h.getv("arx"), what=f"gen {g} arx")
        assert_close(alg.state("xmean"), h.getv("xmean"), what=f"gen {g} xmean")
        assert alg.sigma == h.getd("sigma"), f"gen {g} sigma (rank-based: exact)"
        assert_close(alg.state("pc"), h.getv("pc"), what=f"gen {g} pc")
        ml = h.geti("memlen")
        assert alg.memory_length == ml
        rows = h.getv("jarr")[:ml].astype(int)
        assert_close(alg.state("vmat").reshape(-1, n)[rows], h.mat("vmat")[rows], what=f"gen {g} vmat")
        assert alg.converged() == h.converged()
    assert alg.n_evals == h.geti("fev")


def test_rademacher_replay_and_full_optimize(pkg, port_oracle):
    """Default constructor arguments (Rademacher draws, 2015 variant) on a tie-free objective, to the reference's own stop."""
    P = port_oracle
    n = 12
    lo, hi = -10 * np.ones(n), 10 * np.ones(n)
    P.seed(42)
    guess = P.uniform(-10.0, 10.0, n)
    x_o, fev_o, conv_o = P.cma(O.LM, 30000, 1e-8, 16).optimize(O.ROSENBROCK, lo, hi, guess)
    alg = pkg.LmCMAES(mfev=30000, tol=1e-8, np=16)
    alg.set_draw_replay(42)
    alg.burn_uniforms(n)
    sol = alg.optimize(pkg.Rosenbrock, lo, hi, guess)
    assert sol.n_evals == fev_o and sol.converged == conv_o
    assert_close(sol.x, x_o, rel=1e-8, what="best x")


def test_ipop_with_lm_base(pkg, port_oracle):
    P = port_oracle
    n = 6
    lo, hi = -5.12 * np.ones(n), 5.12 * np.ones(n)
    P.seed(5)
    guess = P.uniform(-5, 5, n)
    base_o = P.cma(O.LM, 1000, 1e-8, 10, rademacher=False)
    x_o, fev_o, _ = P.restart_optimize(0, base_o, 20000, O.RASTRIGIN, lo, hi, guess)
    pkg.seed(5)
    from test_gpu_api import pkg_uniform
    assert np.array_equal(pkg_uniform(pkg, -5, 5, n), guess)
    base = pkg.LmCMAES(mfev=1000, tol=1e-8, np=10, rademacher=False)
    base.set_draw_replay_global()
    sol = pkg.IPopCMAES(base, mfev=20000).optimize(pkg.Rastrigin, lo, hi, guess)
    assert sol.n_evals == fev_o
    assert_close(sol.x, x_o, rel=1e-8, what="restart best x")


def test_free_running_solves_and_resumes(pkg):
    n = 40
    lo, hi = -10 * np.ones(n), 10 * np.ones(n)
    solved = 0
    for seed in range(1, 5):
        pkg.seed(seed)
        alg = pkg.LmCMAES(mfev=200000, tol=1e-12, np=16)
        sol = alg.optimize(pkg.Sphere, lo, hi, 3 * np.ones(n))
        solved += float(np.sum(np.asarray(sol.x) ** 2)) < 1e-8
    assert solved >= 3
    a = pkg.LmCMAES(mfev=10 ** 6, tol=1e-12, np=16, memory=7)
    a.initialize(pkg.Rastrigin, lo, hi, np.linspace(-3, 4, n))
    for _ in range(25):
        a.iterate()
    blob = a.checkpoint()
    for _ in range(15):
        a.iterate()
    b = pkg.LmCMAES(mfev=10 ** 6, tol=1e-12, np=16, memory=7)
    b.initialize(pkg.Rastrigin, lo, hi, np.zeros(n))
    b.restore(blob)
    for _ in range(15):
        b.iterate()
    assert a.sigma == b.sigma and a.n_evals == b.n_evals and a.memory_length == b.memory_length
    for f in ("xmean", "pc", "arx", "vmat", "pcmat", "b", "d"):
        assert np.array_equal(a.state(f), b.state(f)), f
