"""xNES (nes/xnes.cpp) on the device against the oracle.

T1, C ABI, teacher-forced per generation: oracle state in (mu, B, sigma) -> replayed draws -> points (1e-10); on the oracle's
points: fitness bits and ranking bit-exact; then G_sigma, mu and B = B expm(eta_B G_B / 2) at 1e-10 -- the matrix exponential
through the engine's own eigensolver (n >= 256) or cuSOLVER's (smaller n): basis independent, so no replay is needed.
T2, Python API (py/multivariate_py.cpp:153-157): trajectories under draw replay to the reference's own stop, optimize()."""
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


@pytest.mark.parametrize("obj,n,gens,a0,etamu", [(O.ROSENBROCK, 10, 40, 1.0, 1.0), (O.RASTRIGIN, 40, 25, 2.0, 0.5),
                                                  (O.ELLIPSOID, 130, 6, 1.0, 1.0), (O.ROSENBROCK, 300, 3, 1.0, 1.0),
                                                  (O.ROSENBROCK, 1000, 1, 1.0, 1.0)])
def test_teacher_forced_generations(ctx, port_oracle, obj, n, gens, a0, etamu):
    P = port_oracle
    P.seed(42)
    lo, hi = -5 * np.ones(n), 5 * np.ones(n)
    h = P.xnes(10 ** 9, 1e-12, a0, etamu)
    h.init(obj, lo, hi, np.zeros(n))
    npts = h.geti("np")
    g = _cabi.Cma(ctx, _cabi.CMA_XNES, n, npts, h.getv("u"), {}, seed=3)
    g.reset(np.zeros(n))
    assert np.array_equal(g.get(_cabi.F_B), np.eye(n))
    objective = _cabi.Objective(ctx, n, builtin=obj)
    etab = h.getd("etab")
    for gen in range(gens):
        mu0, B0, sigma = h.getv("mu"), h.mat("B"), h.getd("sigma")
        g.set(_cabi.F_XMEAN, mu0)
        g.set(_cabi.F_B, B0)
        h.iterate()
        # the oracle's points come back SORTED; the draws z are sorted with them, so replay them in that order: point k of
        # the engine is the oracle's k-th best, and the ranking must come out as the identity
        Z, X, f = h.mat("z"), h.mat("x"), h.getv("f")
        g.sample(sigma, gen, z=Z)
        assert_close(g.get(_cabi.F_ARX), X, what=f"gen {gen} points")
        g.set(_cabi.F_ARX, X)
        g.evaluate(objective)
        g.rank(0)
        assert len(set(f.tolist())) == npts, "parity cases must be tie-free"
        assert np.array_equal(g.get(_cabi.F_FITNESS), f), f"gen {gen} fitness bits"
        assert np.array_equal(g.rank_idx(), np.arange(npts, dtype=np.int32)), f"gen {gen} ranking"
        gs = g.xnes_update(sigma, etamu, etab)
        rep = g.fetch()
        assert rep.eig_info == 0
        assert_close(gs, h.getd("Gsigma"), rel=1e-9, what=f"gen {gen} Gsigma")  # a trace of n cancelling terms
        assert_close(g.get(_cabi.F_XMEAN), h.getv("mu"), what=f"gen {gen} mu")
        assert_close(g.get(_cabi.F_B), h.mat("B"), what=f"gen {gen} B")


@pytest.mark.parametrize("obj,n,gens", [(O.ROSENBROCK, 10, 120), (O.RASTRIGIN, 40, 60), (O.ELLIPSOID, 100, 15)])
def test_trajectory_replay(pkg, port_oracle, obj, n, gens):
    P = port_oracle
    lo, hi = -5 * np.ones(n), 5 * np.ones(n)
    P.seed(42)
    h = P.xnes(10 ** 9, 1e-12)
    h.init(obj, lo, hi, np.ones(n))
    alg = pkg.xNES(mfev=10 ** 9, tol=1e-12)
    alg.set_draw_replay(42)
    alg.initialize(pkg.Objective.builtin(obj), lo, hi, np.ones(n))
    assert alg.np == h.geti("np")
    for g in range(gens):
        h.iterate()
        alg.iterate()
        order = alg.state("rank").astype(int)
        assert_close(alg.state("x").reshape(-1, n)[order], h.mat("x"), what=f"gen {g} points (sorted)")
        assert_close(alg.state("f"), h.getv("f"), what=f"gen {g} fitness (sorted)")
        assert_close(alg.state("mu"), h.getv("mu"), what=f"gen {g} mu")
        assert_close(alg.sigma, h.getd("sigma"), what=f"gen {g} sigma")
        assert_close(alg.state("B"), h.getv("B"), what=f"gen {g} B")
        assert alg.converged() == h.converged()
    assert alg.n_evals == h.geti("fev")


def test_full_optimize_replay_and_free_run(pkg, port_oracle):
    P = port_oracle
    n = 8
    lo, hi = -5 * np.ones(n), 5 * np.ones(n)
    P.seed(7)
    x_o, fev_o, conv_o = P.xnes(20000, 1e-9).optimize(O.SPHERE, lo, hi, np.ones(n))
    alg = pkg.xNES(mfev=20000, tol=1e-9)
    alg.set_draw_replay(7)
    sol = alg.optimize(pkg.Sphere, lo, hi, np.ones(n))
    assert sol.n_evals == fev_o and sol.converged == conv_o
    assert_close(sol.x, x_o, rel=1e-8, what="best x")
    solved = 0
    for seed in range(1, 5):
        pkg.seed(seed)
        s = pkg.xNES(mfev=40000, tol=1e-14).optimize(pkg.Sphere, lo, hi, np.ones(n))
        solved += float(np.sum(np.asarray(s.x) ** 2)) < 1e-8
    assert solved >= 3
