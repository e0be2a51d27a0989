"""Tagged-draw replay parity of the JADE / L-SHADE device generation against the oracle
(tier T3), through the C ABI.

The oracle's recorder (oracle/port, pinned bit-for-bit to the compiled reference) emits the
draws the reference consumed; the engine replays them and must reproduce, per generation:
success mask / indices bit-exact, swarm order bit-exact, positions and fitness within 1e-10
(the first generation bit-exact), archive, and the adaptive means within 1e-10."""
import numpy as np
import pytest

from bboptpy_b200 import _cabi
from oracle import oracle as O
from util import assert_close

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    return _cabi.Context(0)


CASES = [
    # kind, obj, n, np, gens, box, mfev, kwargs
    (O.JADE, O.RASTRIGIN, 30, 50, 60, 5.12, 10 ** 6, {}),
    (O.SHADE, O.RASTRIGIN, 30, 50, 60, 5.12, 20000, {"h": 10}),
    (O.JADE, O.ROSENBROCK, 10, 20, 80, 10.0, 10 ** 6, {"archive": False, "repaircr": False}),
    (O.SHADE, O.ACKLEY, 8, 30, 60, 32.768, 4000, {"h": 5}),
    (O.JADE, O.SPHERE, 64, 200, 25, 5.0, 10 ** 6, {"pelite": 0.2}),
    (O.JADE, O.RASTRIGIN, 1000, 1000, 4, 5.12, 10 ** 7, {}),
    (O.SHADE, O.RASTRIGIN, 1000, 1000, 4, 5.12, 10 ** 7, {}),
]


@pytest.mark.parametrize("kind,obj,n,np_,gens,box,mfev,kw", CASES)
def test_tagged_draw_replay(ctx, port_oracle, kind, obj, n, np_, gens, box, mfev, kw):
    P = port_oracle
    P.seed(7)
    lo, hi = -box * np.ones(n), box * np.ones(n)
    h = P.de(kind, mfev, np_, 1e-12, **kw)
    h.init(obj, lo, hi)
    x0, f0 = h.swarm()
    g = _cabi.De(ctx, kind, n, np_, lo, hi, mfev=mfev, archive=kw.get("archive", True), repaircr=kw.get("repaircr", True),
                 pelite=kw.get("pelite", 0.05), h=kw.get("h", 100))
    objective = _cabi.Objective(ctx, n, builtin=obj)
    g.init(objective, x0)
    xg, fg = g.swarm()
    assert np.array_equal(xg, x0) and np.array_equal(fg, f0)
    max_rounds = 0
    for gen in range(gens):
        np_gen = h.geti("np")
        h.iterate()
        t = h.trace()
        rep = g.generation(objective, t)
        max_rounds = max(max_rounds, rep.rounds)
        succ, ft, F, CR = g.trace(np_gen)
        assert np.array_equal(succ > 0, t["succ"] > 0), f"gen {gen} success mask"
        if kind == O.SHADE:
            assert np.array_equal(succ == 2, t["succ"] == 2), f"gen {gen} strict-improvement mask"
        assert_close(F, t["F"], what=f"gen {gen} F")
        assert_close(CR, t["CR"], what=f"gen {gen} CR")
        assert_close(ft, t["ftrial"], what=f"gen {gen} trial fitness")
        if gen == 0:
            assert np.array_equal(ft, t["ftrial"]), "first generation is bit-exact"
        xo, fo = h.swarm()
        xg, fg = g.swarm()
        assert rep.np == h.geti("np") and rep.fev == h.geti("fev") and rep.larch == h.geti("larch")
        assert rep.nsucc == int((t["succ"] > 0).sum())
        assert_close(fg, fo, what=f"gen {gen} swarm fitness")
        assert_close(xg, xo, what=f"gen {gen} swarm positions")
        ao, ag = h.archive(), g.archive()
        assert ao.shape == ag.shape
        if ao.size:
            assert_close(ag, ao, what=f"gen {gen} archive")
        if kind == O.JADE:
            assert_close(rep.mucr, h.getd("mucr"), what="mucr")
            assert_close(rep.muf, h.getd("muf"), what="muf")
        else:
            mcr, mf = g.memory()
            hh = kw.get("h", 100)
            assert_close(mcr[:hh], h.getv("MCR"), what="MCR")
            assert_close(mf[:hh], h.getv("MF"), what="MF")
            assert rep.kslot == h.geti("k")
        # Welford M2 of the radii (de/jade.cpp:228-245) from the device norms
        r = np.linalg.norm(xo, axis=1)
        assert_close(rep.radius_m2, np.sum((r - r.mean()) ** 2), rel=1e-8, what="radius M2")
        assert rep.fbest == fg.min()
    assert max_rounds >= 2  # the in-place dependencies really were exercised


@pytest.mark.parametrize("kind", [O.JADE, O.SHADE])
def test_free_running_philox_de_converges(ctx, kind):
    n, np_ = 20, 64
    lo, hi = -5.0 * np.ones(n), 5.0 * np.ones(n)
    g = _cabi.De(ctx, kind, n, np_, lo, hi, mfev=40000, seed=5)
    obj = _cabi.Objective(ctx, n, builtin=O.SPHERE)
    g.init(obj)
    x, f = g.swarm()
    assert (x >= -5).all() and (x <= 5).all()
    f0 = f.min()
    rep = None
    for _ in range(150):
        rep = g.generation(obj)
    assert rep.fbest < 1e-3 * f0
    x, f = g.swarm()
    assert (x >= -5).all() and (x <= 5).all()
    assert np.allclose(f, (x * x).sum(1), rtol=1e-12)


@pytest.mark.parametrize("kind,obj,R,n,np_,gens", [(O.JADE, O.RASTRIGIN, 5, 50, 64, 12), (O.SHADE, O.ROSENBROCK, 5, 50, 64, 12),
                                                   (O.JADE, O.ACKLEY, 5, 50, 64, 12),
                                                   # >= one run per SM: every CTA owns whole runs, no grid-wide barrier
                                                   (O.SHADE, O.RASTRIGIN, 160, 24, 20, 8), (O.JADE, O.ROSENBROCK, 300, 600, 12, 5)])
def test_batched_runs_equal_single_runs(ctx, kind, obj, R, n, np_, gens):
    """bbgpu_de_generation_batch: R independent runs in one launch are bit-identical to the same runs
    stepped one handle at a time (same seeds)."""
    lo, hi = -5.0 * np.ones(n), 5.0 * np.ones(n)
    objective = _cabi.Objective(ctx, n, builtin=obj)
    single = [_cabi.De(ctx, kind, n, np_, lo, hi, mfev=40 * np_, seed=100 + r) for r in range(R)]
    batch = [_cabi.De(ctx, kind, n, np_, lo, hi, mfev=40 * np_, seed=100 + r) for r in range(R)]
    for g in single + batch:
        g.init(objective)
    for gen in range(gens):
        reps1 = [g.generation(objective) for g in single]
        reps2 = _cabi.De.generation_batch(batch, objective)
        for r in range(R):
            assert (reps1[r].np, reps1[r].larch, reps1[r].nsucc, reps1[r].fev) == \
                   (reps2[r].np, reps2[r].larch, reps2[r].nsucc, reps2[r].fev)
            assert reps1[r].fbest == reps2[r].fbest and reps1[r].radius_m2 == reps2[r].radius_m2
            x1, f1 = single[r].swarm()
            x2, f2 = batch[r].swarm()
            assert np.array_equal(x1, x2) and np.array_equal(f1, f2), f"gen {gen} run {r}"
            assert np.array_equal(single[r].archive(), batch[r].archive())
    # different seeds really are different runs
    assert not np.array_equal(batch[0].swarm()[1], batch[1].swarm()[1])


def test_de_large_n_tiles(ctx, port_oracle):
    """n > 1024 exercises the tiled term staging (Rosenbrock's neighbour coordinate across the tile edge)."""
    P = port_oracle
    P.seed(3)
    n, np_ = 2500, 24
    lo, hi = -10.0 * np.ones(n), 10.0 * np.ones(n)
    for obj in (O.ROSENBROCK, O.ACKLEY):
        h = P.de(O.JADE, 10 ** 6, np_, 1e-12)
        h.init(obj, lo, hi)
        x0, f0 = h.swarm()
        g = _cabi.De(ctx, O.JADE, n, np_, lo, hi, mfev=10 ** 6)
        objective = _cabi.Objective(ctx, n, builtin=obj)
        g.init(objective, x0)
        for gen in range(3):
            h.iterate()
            t = h.trace()
            g.generation(objective, t)
            succ, ft, F, CR = g.trace(np_)
            assert np.array_equal(succ > 0, t["succ"] > 0)
            assert_close(ft, t["ftrial"], what="trial fitness")
            xo, fo = h.swarm()
            xg, fg = g.swarm()
            assert_close(xg, xo, what="swarm")


@pytest.mark.parametrize("kind,obj", [(O.JADE, O.RASTRIGIN), (O.SHADE, O.ROSENBROCK)])
def test_external_objective_path_equals_fused(ctx, kind, obj):
    """A user device objective (here: a callback that forwards to the registry kernel, so the fitness bits are the
    same) drives the split generation -- build / gather / evaluate / select per round -- and must reproduce the fused
    kernel exactly: same draws, same fixed point."""
    n, np_ = 40, 48
    lo, hi = -5.0 * np.ones(n), 5.0 * np.ones(n)
    builtin = _cabi.Objective(ctx, n, builtin=obj)
    calls = []

    def dev_fn(d_X, ld, n_, count, d_f, stream, user):
        calls.append(count)
        return ctx.L.bbgpu_eval(ctx.h, builtin.h, d_X, ld, count, d_f)

    external = _cabi.Objective(ctx, n, device_fn=dev_fn)
    a = _cabi.De(ctx, kind, n, np_, lo, hi, mfev=4000, seed=21)
    b = _cabi.De(ctx, kind, n, np_, lo, hi, mfev=4000, seed=21)
    a.init(builtin)
    b.init(external)
    for gen in range(15):
        ra = a.generation(builtin)
        rb = b.generation(external)
        assert (ra.np, ra.larch, ra.nsucc, ra.fev, ra.rounds) == (rb.np, rb.larch, rb.nsucc, rb.fev, rb.rounds), f"gen {gen}"
        assert ra.fbest == rb.fbest and ra.radius_m2 == rb.radius_m2
        xa, fa = a.swarm()
        xb, fb = b.swarm()
        assert np.array_equal(xa, xb) and np.array_equal(fa, fb), f"gen {gen}"
        assert np.array_equal(a.archive(), b.archive())
    assert len(calls) > 15 and max(calls) <= np_


def test_de_python_callable_objective():
    """The reference's everyday entry: JADE / SHADE with an arbitrary Python callable (py/multivariate_py.cpp:385-388)."""
    import bboptpy_b200 as pkg
    n = 6
    lo, hi = -5.0 * np.ones(n), 5.0 * np.ones(n)
    evals = [0]

    def sphere(x):
        evals[0] += 1
        return float(np.sum((x - 1.0) ** 2))

    for alg in (pkg.JADE(mfev=6000, np=30, tol=1e-8), pkg.SHADE(mfev=6000, npinit=30, tol=1e-8)):
        evals[0] = 0
        sol = alg.optimize(sphere, lo, hi, lo)
        assert sphere(sol.x) < 1e-6
        # n_evals is the reference's count (one per candidate and generation); the dependency rounds may evaluate a
        # candidate's trial again after one of its parents changed, so the callable sees somewhat more calls
        assert sol.n_evals <= evals[0] - 1 <= 2 * sol.n_evals
