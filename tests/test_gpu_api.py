"""End-to-end parity through the reference-facing Python API (the pybind module over the host
C++ classes over the C ABI): trajectory replay (SURVEY.md 8(c) tier T2), the README example, the
restart drivers and the DE classes."""
import numpy as np
import pytest

from oracle import oracle as O
from util import assert_close

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pkg():
    import bboptpy_b200
    return bboptpy_b200


def fx(x):  # the README objective, examples/multivariate.py:6-10
    total = 0.0
    for x1, x2 in zip(x[:-1], x[1:]):
        total += 100 * (x2 - x1 ** 2) ** 2 + (1 - x1) ** 2
    return total


def test_readme_example_with_the_python_callable(pkg):
    n = 10
    solved = 0
    for seed in range(1, 6):  # free-running draws: a run may end in the 10-D Rosenbrock's local minimum (f = 3.99), as in the reference
        pkg.seed(seed)
        alg = pkg.ActiveCMAES(mfev=10000, tol=1e-4, np=20)
        sol = alg.optimize(fx, lower=-10 * np.ones(n), upper=10 * np.ones(n), guess=np.random.default_rng(0).uniform(-10, 10, n))
        assert sol.converged and sol.n_evals <= 10000
        assert "objective calls: %d" % sol.n_evals in str(sol)
        solved += fx(sol.x) < 1e-3
    assert solved >= 3


def test_iterative_api(pkg):
    n = 10
    alg = pkg.ActiveCMAES(mfev=10000, tol=1e-4, np=20)
    alg.initialize(f=pkg.Rosenbrock, lower=-10 * np.ones(n), upper=10 * np.ones(n), guess=3 * np.ones(n))
    s0 = alg.solution()
    assert s0.n_evals == 0 and np.array_equal(s0.x, 3 * np.ones(n))
    for _ in range(400):
        alg.iterate()
    s = alg.solution()
    assert s.n_evals == 400 * 20 and fx(s.x) < 1e-6


@pytest.mark.parametrize("cls,kind,obj,n,lam,gens", [
    ("CMAES", O.VANILLA, O.ROSENBROCK, 10, 20, 80),
    ("ActiveCMAES", O.ACTIVE, O.ROSENBROCK, 10, 20, 80),
    ("SepCMAES", O.SEP, O.ELLIPSOID, 10, 20, 80),
    ("CMAES", O.VANILLA, O.RASTRIGIN, 40, 15, 60),
    ("ActiveCMAES", O.ACTIVE, O.ACKLEY, 60, 128, 25),
    ("SepCMAES", O.SEP, O.ELLIPSOID, 1000, 256, 25),
])
def test_trajectory_replay(pkg, port_oracle, cls, kind, obj, n, lam, gens):
    """T2: replayed draws + replayed eigenbasis; per-generation state within 1e-10, ranking bit-exact.
    Trajectory lengths are bounded: two fp64 implementations that sum in different orders drift apart
    slowly (the separable run on the 10-D ellipsoid crosses 1e-10 of max|m| after ~107 generations, as m
    itself converges to 0); the whole-optimize anchor below checks the long run at the north star's 1e-8."""
    P = port_oracle
    box = {O.ROSENBROCK: 10.0, O.ELLIPSOID: 10.0, O.RASTRIGIN: 5.12, O.ACKLEY: 32.768}[obj]
    lo, hi = -box * np.ones(n), box * np.ones(n)
    guess = np.random.default_rng(17).uniform(-box, box, n)
    P.seed(42)
    h = P.cma(kind, 10 ** 9, 1e-12, lam, adjustlr=True)
    h.init(obj, lo, hi, guess)
    alg = getattr(pkg, cls)(mfev=10 ** 9, tol=1e-12, np=lam)
    alg.set_draw_replay(42)
    if kind != O.SEP:
        def basis(fev, n_):
            assert h.geti("eigenlastev") == fev, "the engine decomposes exactly when the reference does"
            return h.mat("B"), h.getv("D")
        alg.set_eigen_replay(basis)
    alg.initialize(pkg.Objective.builtin(obj), lo, hi, guess)
    for g in range(gens):
        h.iterate()
        alg.iterate()
        assert np.array_equal(alg.state("rank"), h.getv("fit_idx")), f"gen {g} ranking"
        assert_close(alg.state("arx"), h.getv("arx"), what=f"gen {g} arx")
        assert_close(alg.state("xmean"), h.getv("xmean"), what=f"gen {g} xmean")
        assert_close(alg.sigma, h.getd("sigma"), what=f"gen {g} sigma")
        assert_close(alg.state("pc"), h.getv("pc"), what=f"gen {g} pc")
        assert_close(alg.state("ps"), h.getv("ps"), what=f"gen {g} ps")
        if kind == O.SEP:
            assert_close(alg.state("C"), h.getv("C"), what=f"gen {g} C")
        else:
            lower = np.tril_indices(n)
            assert_close(alg.state("C").reshape(n, n)[lower], h.mat("C")[lower], what=f"gen {g} C")
            assert_close(alg.state("invsqrtC"), h.getv("invsqrtC"), what=f"gen {g} invsqrtC")
        assert alg.converged() == h.converged()
    assert alg.n_evals == h.geti("fev")
    if kind != O.SEP:
        assert alg.n_decompositions > 0


def test_readme_anchor_full_optimize_replay(pkg, port_oracle):
    """Whole optimize() under replay: same evaluation count, same stop, final best within 1e-8."""
    P = port_oracle
    n = 10
    P.seed(42)
    guess = P.uniform(-10.0, 10.0, n)
    h = P.cma(O.ACTIVE, 10000, 1e-4, 20)
    lo, hi = -10 * np.ones(n), 10 * np.ones(n)
    h.init(O.ROSENBROCK, lo, hi, guess)
    alg = pkg.ActiveCMAES(mfev=10000, tol=1e-4, np=20)
    alg.set_draw_replay(42)
    alg.burn_uniforms(n)  # the oracle spent n uniform_real draws on the guess before the first normal
    alg.set_eigen_replay(lambda fev, n_: (h.mat("B"), h.getv("D")))
    alg.initialize(pkg.Rosenbrock, lo, hi, guess)
    converged = False
    while alg.n_evals < 10000:
        h.iterate()
        alg.iterate()
        co, cg = h.converged(), alg.converged()
        assert co == cg
        if cg:
            converged = True
            break
    assert converged and alg.n_evals == h.geti("fev") == 6020
    sol = alg.solution()
    assert abs(P.eval(O.ROSENBROCK, sol.x) - 7.8273165905451219e-07) <= 1e-8


@pytest.mark.parametrize("driver,rk", [("IPopCMAES", 0), ("BiPopCMAES", 1)])
def test_restart_drivers_replay_with_separable_base(pkg, port_oracle, driver, rk):
    """IPOP / BIPOP control flow (restart schedule, budgets, +1 evaluation, regime choice) against the
    oracle, with the separable base so that no eigenbasis is involved: the whole run (8 IPOP runs /
    24 BIPOP runs, lambda from 9 to 288) replays with identical evaluation counts.  adjustlr=False: with
    the learning-rate adjustment this configuration drives the REFERENCE itself to NaN fitness at
    lambda >= 144 (c_i turns negative), where std::sort's comparator is undefined."""
    P = port_oracle
    n = 6
    lo, hi = -5.12 * np.ones(n), 5.12 * np.ones(n)
    P.seed(5)
    guess = P.uniform(-5, 5, n)
    base_o = P.cma(O.SEP, 1000, 1e-8, 10, adjustlr=False)
    x_o, fev_o, trace = P.restart_optimize(rk, base_o, 30000, O.RASTRIGIN, lo, hi, guess)
    pkg.seed(5)
    np_guess = pkg_uniform(pkg, -5, 5, n)
    assert np.array_equal(np_guess, guess)
    base = pkg.SepCMAES(mfev=1000, tol=1e-8, np=10, adjustlr=False)
    base.set_draw_replay_global()
    alg = getattr(pkg, driver)(base, mfev=30000)
    sol = alg.optimize(pkg.Rastrigin, lo, hi, guess)
    assert sol.n_evals == fev_o
    assert_close(sol.x, x_o, rel=1e-8, what="restart best x")


@pytest.mark.parametrize("driver,rk,base_cls,base_kind", [("IPopCMAES", 0, "CMAES", O.VANILLA), ("BiPopCMAES", 1, "CMAES", O.VANILLA),
                                                        ("BiPopCMAES", 1, "ActiveCMAES", O.ACTIVE)])
def test_restart_drivers_replay_with_full_covariance_base(pkg, port_oracle, driver, rk, base_cls, base_kind):
    """IPOP / BIPOP with a FULL-covariance base (cma/ipop_cmaes.cpp:65-162, cma/bipop_cmaes.cpp:61-267) under draw replay
    + eigenbasis replay: the oracle logs (fev, B, D) at every decomposition of the whole restart sequence and the engine
    adopts them event by event.  Exercises the re-init quirk on the device (cmaes.cpp:53-60: B and C keep their
    off-diagonal entries across restarts of the same object, `bbgpu_cma_reset(keep_offdiag)`): the first generation of
    every restart samples from the carried-over B before its first decomposition."""
    P = port_oracle
    n, budget = 6, 20000
    lo, hi = -5.12 * np.ones(n), 5.12 * np.ones(n)
    P.seed(5)
    guess = P.uniform(-5, 5, n)
    base_o = P.cma(base_kind, 1000, 1e-8, 10)
    base_o.eig_log(True)
    x_o, fev_o, trace = P.restart_optimize(rk, base_o, budget, O.RASTRIGIN, lo, hi, guess)
    events = base_o.eig_events()
    assert len(trace) >= 4 and len(events) > 100, "at least 3 restarts, each with its decompositions"
    pkg.seed(5)
    assert np.array_equal(pkg_uniform(pkg, -5, 5, n), guess)
    base = getattr(pkg, base_cls)(mfev=1000, tol=1e-8, np=10)
    base.set_draw_replay_global()
    used = [0]

    def basis(fev, n_):
        k = used[0]
        assert k < len(events), "the engine decomposes more often than the reference"
        assert events[k][0] == fev, f"decomposition {k}: engine at fev {fev}, reference at {events[k][0]}"
        used[0] += 1
        return events[k][1], events[k][2]

    base.set_eigen_replay(basis)
    alg = getattr(pkg, driver)(base, mfev=budget)
    sol = alg.optimize(pkg.Rastrigin, lo, hi, guess)
    assert used[0] == len(events), "every decomposition of the reference happened on the device too"
    assert sol.n_evals == fev_o
    assert_close(sol.x, x_o, rel=1e-8, what="restart best x")


def pkg_uniform(pkg, a, b, n):
    return np.array([pkg.uniform(a, b) for _ in range(n)])


@pytest.mark.parametrize("cls,kind", [("JADE", O.JADE), ("SHADE", O.SHADE)])
def test_de_classes_replay(pkg, port_oracle, cls, kind):
    P = port_oracle
    n, np_, box, mfev = 20, 40, 5.12, 30000
    lo, hi = -box * np.ones(n), box * np.ones(n)
    P.seed(7)
    h = P.de(kind, mfev, np_, 1e-12, h=10)
    h.init(O.RASTRIGIN, lo, hi)
    alg = getattr(pkg, cls)(mfev=mfev, np=np_, tol=1e-12) if kind == O.JADE else pkg.SHADE(mfev=mfev, npinit=np_, tol=1e-12, h=10)
    alg.set_initial_population(h.swarm()[0])
    alg.set_draw_replay(lambda gen: h.trace())
    alg.initialize(pkg.Rastrigin, lo, hi, lo)
    for g in range(80):
        h.iterate()
        alg.iterate()
        xo, fo = h.swarm()
        assert alg.popsize == h.geti("np") and alg.n_evals == h.geti("fev")
        assert_close(alg.fitness(), fo, what=f"gen {g} fitness")
        assert_close(alg.swarm(), xo, what=f"gen {g} swarm")


@pytest.mark.parametrize("cls", ["JADE", "SHADE"])
def test_de_free_running_optimize(pkg, cls):
    n = 10
    pkg.seed(11)
    alg = pkg.JADE(mfev=60000, np=50, tol=1e-7) if cls == "JADE" else pkg.SHADE(mfev=60000, npinit=50, tol=1e-7)
    sol = alg.optimize(pkg.Sphere, -5 * np.ones(n), 5 * np.ones(n), np.zeros(n))
    assert float((sol.x ** 2).sum()) < 1e-8
    assert sol.n_evals <= 60000 + 50


@pytest.mark.parametrize("cls", ["JADE", "SHADE"])
def test_de_repeated_optimize_calls_are_independent_runs(pkg, cls):
    """The reference's process-global engine keeps advancing, so a second optimize() on the same object is a new run
    (multi-start loops rely on it); a key fixed with set_stream_key() makes the run reproducible instead."""
    n = 6
    lo, hi = -5.12 * np.ones(n), 5.12 * np.ones(n)
    make = (lambda: pkg.JADE(mfev=600, np=20, tol=0.0)) if cls == "JADE" else (lambda: pkg.SHADE(mfev=600, npinit=20, tol=0.0))
    alg = make()
    a = alg.optimize(pkg.Rastrigin, lo, hi, lo).x
    b = alg.optimize(pkg.Rastrigin, lo, hi, lo).x
    assert not np.array_equal(a, b)
    fixed = make()
    fixed.set_stream_key(12345)
    c = fixed.optimize(pkg.Rastrigin, lo, hi, lo).x
    d = fixed.optimize(pkg.Rastrigin, lo, hi, lo).x
    assert np.array_equal(c, d)


def test_python_exception_in_the_objective_propagates(pkg):
    """py/multivariate_py.cpp:385-388: an exception raised by the user's callable surfaces as itself, not as a generic
    engine error."""
    class Boom(Exception):
        pass

    calls = [0]

    def f(x):
        calls[0] += 1
        if calls[0] > 30:
            raise Boom("objective failed")
        return float(np.sum(x * x))

    alg = pkg.CMAES(mfev=1000, tol=1e-8, np=10)
    with pytest.raises(Boom):
        alg.optimize(f, -np.ones(4), np.ones(4), 0.5 * np.ones(4))


def test_bipop_free_running(pkg):
    n = 8
    pkg.seed(2)
    base = pkg.CMAES(mfev=1000, tol=1e-8, np=10)
    alg = pkg.BiPopCMAES(base, mfev=40000)
    sol = alg.optimize(pkg.Rastrigin, -5.12 * np.ones(n), 5.12 * np.ones(n), 3 * np.ones(n))
    assert sol.n_evals >= 1000
    x = sol.x
    assert 10 * n + float(np.sum(x * x - 10 * np.cos(2 * np.pi * x))) < 20.0
