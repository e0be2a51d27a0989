"""The CPU restatement (oracle/port) against golden vectors generated from the compiled
reference (tests/golden/make_golden.py).  Bit-exact: every comparison is array_equal."""
import glob
import os

import numpy as np
import pytest

from oracle import oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return np.load(os.path.join(GOLD, name))


def test_rng_stack_matches_libstdcxx(port_oracle):
    P, g = port_oracle, load("rng.npz")
    P.seed(123)
    assert np.array_equal(P.draw(3, 64), g["rng_u32"])
    P.seed(123)
    assert np.array_equal(P.draw(0, 65), g["rng_normal"])
    P.seed(123)
    assert np.array_equal(P.draw(1, 32, -10.0, 10.0), g["rng_real"])
    P.seed(123)
    assert np.array_equal(P.draw(2, 64, 0, 999), g["rng_int1000"])
    P.seed(123)
    # a 1-element range still consumes one engine draw (uniform_int_dist.h, Lemire path)
    assert np.array_equal(P.draw(2, 8, 0, 0), g["rng_int1"])
    P.seed(123)
    P.draw(2, 8, 0, 0)
    nxt = P.draw(3, 1)[0]
    P.seed(123)
    assert P.draw(3, 9)[8] == nxt


def test_objectives_match(port_oracle):
    g = load("objectives.npz")
    for obj in range(5):
        for n in (1, 2, 10, 33, 1000):
            assert port_oracle.eval(obj, g[f"x_{obj}_{n}"]) == g[f"f_{obj}_{n}"][0]


def test_objective_values_are_the_textbook_functions(port_oracle):
    rng = np.random.default_rng(1)
    x = rng.uniform(-5, 5, 257)
    n = x.size
    ref = {
        O.SPHERE: np.sum(x * x),
        O.ELLIPSOID: np.sum(10.0 ** (6.0 * np.arange(n) / (n - 1.0)) * x * x),
        O.ROSENBROCK: np.sum(100 * (x[1:] - x[:-1] ** 2) ** 2 + (1 - x[:-1]) ** 2),
        O.RASTRIGIN: 10 * n + np.sum(x * x - 10 * np.cos(2 * np.pi * x)),
        O.ACKLEY: -20 * np.exp(-0.2 * np.sqrt(np.mean(x * x))) - np.exp(np.mean(np.cos(2 * np.pi * x))) + 20 + np.e,
    }
    for obj, val in ref.items():
        assert port_oracle.eval(obj, x) == pytest.approx(val, rel=1e-13)


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "cma_*.npz"))))
def test_cma_trajectory_matches_reference(port_oracle, path):
    P, g = port_oracle, np.load(path)
    kind, obj, n, lam, box, bound = g["meta"]
    kind, obj, n, lam = int(kind), int(obj), int(n), int(lam)
    gens = sorted({int(k.split("_")[0][1:]) for k in g.files if k.startswith("g") and k != "guess"})
    P.seed(42)
    guess = P.uniform(-box, box, n)
    assert np.array_equal(guess, g["guess"])
    h = P.cma(kind, 10 ** 9, 1e-12, lam, bound=bool(bound))
    h.init(obj, -box * np.ones(n), box * np.ones(n), guess)
    for it in range(1, max(gens) + 1):
        h.iterate()
        if it in gens:
            s = h.snapshot()
            for k in ("xmean", "pc", "ps", "D", "C", "fit_val", "fit_idx"):
                assert np.array_equal(np.asarray(s[k]), g[f"g{it}_{k}"]), (it, k)
            assert s["sigma"] == g[f"g{it}_sigma"]
            if kind != O.SEP:
                assert np.array_equal(s["B"], g[f"g{it}_B"])
                assert np.array_equal(s["invsqrtC"], g[f"g{it}_invsqrtC"])
            if it == 1:
                assert np.array_equal(s["arx"], g["g1_arx"])


def test_readme_example_anchor(port_oracle):
    """examples/multivariate.py: ActiveCMAES(mfev=10000, tol=1e-4, np=20) on 10-D Rosenbrock."""
    P, g = port_oracle, load("readme_active_rosen10.npz")
    P.seed(42)
    guess = P.uniform(-10.0, 10.0, 10)
    h = P.cma(O.ACTIVE, 10000, 1e-4, 20)
    x, fev, conv = h.optimize(O.ROSENBROCK, -10 * np.ones(10), 10 * np.ones(10), guess)
    assert fev == int(g["fev"][0]) == 6020
    assert conv and int(g["conv"][0]) == 1
    assert np.array_equal(x, g["x"])
    assert P.eval(O.ROSENBROCK, x) == g["f"][0] == 7.8273165905451219e-07


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "de_*.npz"))))
def test_de_trajectory_matches_reference(port_oracle, path):
    P, g = port_oracle, np.load(path)
    kind, obj, n, np_, box, mfev = g["meta"]
    kind, obj, n, np_, mfev = int(kind), int(obj), int(n), int(np_), int(mfev)
    gens = sorted({int(k.split("_")[0][1:]) for k in g.files if k.startswith("g")})
    P.seed(7)
    h = P.de(kind, mfev, np_, 1e-12, h=10)
    h.init(obj, -box * np.ones(n), box * np.ones(n))
    x0, f0 = h.swarm()
    assert np.array_equal(x0, g["init_x"]) and np.array_equal(f0, g["init_f"])
    for it in range(1, max(gens) + 1):
        h.iterate()
        if it in gens:
            x, f = h.swarm()
            assert np.array_equal(x, g[f"g{it}_x"]) and np.array_equal(f, g[f"g{it}_f"])
            assert np.array_equal(h.archive(), g[f"g{it}_arch"])
            if kind == O.JADE:
                assert np.array_equal(np.array([h.getd("mucr"), h.getd("muf")]), g[f"g{it}_mu"])
            else:
                assert np.array_equal(h.getv("MCR"), g[f"g{it}_MCR"])
                assert np.array_equal(h.getv("MF"), g[f"g{it}_MF"])
            assert [h.geti("np"), h.geti("fev"), h.geti("larch")] == list(g[f"g{it}_counts"])


def test_restart_drivers_match_reference(port_oracle):
    P, g = port_oracle, load("restarts.npz")
    n = 6
    for rk, rname in ((0, "ipop"), (1, "bipop")):
        for bk in (O.VANILLA, O.ACTIVE, O.SEP):
            P.seed(5)
            guess = P.uniform(-5, 5, n)
            assert np.array_equal(guess, g[f"{rname}_{bk}_guess"])
            base = P.cma(bk, 1000, 1e-8, 10)
            x, fev, trace = P.restart_optimize(rk, base, 30000, O.RASTRIGIN, -5.12 * np.ones(n),
                                               5.12 * np.ones(n), guess)
            assert fev == int(g[f"{rname}_{bk}_fev"][0])
            assert np.array_equal(x, g[f"{rname}_{bk}_x"])
            assert len(trace) >= 2 and trace[0][0] == 4 + int(3 * np.log(n))


def test_de_trace_is_consistent_with_the_generation(port_oracle):
    """The tagged draws the port records must reproduce its own generation when replayed."""
    P = port_oracle
    n, np_, box = 12, 24, 5.12
    P.seed(3)
    h = P.de(O.JADE, 10 ** 6, np_, 1e-12)
    h.init(O.RASTRIGIN, -box * np.ones(n), box * np.ones(n))
    for _ in range(6):
        x_prev, f_prev = h.swarm()
        order = np.argsort(f_prev, kind="stable")
        x_prev, f_prev = x_prev[order], f_prev[order]
        arch_prev = h.archive()
        mucr, muf = h.getd("mucr"), h.getd("muf")
        h.iterate()
        t = h.trace()
        x, f = x_prev.copy(), f_prev.copy()
        arch = [r.copy() for r in arch_prev]
        for i in range(np_):
            CR = min(max(t["zcr"][i] * 0.1 + mucr, 0.0), 1.0)
            F = min(t["cauchy"][i] * 0.1 + muf, 1.0)
            assert CR == t["CR"][i] and F == t["F"][i]
            r2 = int(t["r2"][i])
            x2 = arch[r2 - np_] if r2 >= np_ else x[r2]
            cross = t["u"][i] < CR
            cross[int(t["jrand"][i])] = True
            trial = np.where(cross, x[i] + F * (x[int(t["ibest"][i])] - x[i]) + F * (x[int(t["r1"][i])] - x2), x[i])
            lo, hi = -box, box
            trial = np.where(trial < lo, (lo + x[i]) / 2.0, np.where(trial > hi, (hi + x[i]) / 2.0, trial))
            ft = P.eval(O.RASTRIGIN, trial)
            assert ft == t["ftrial"][i]
            if ft <= f[i]:
                assert t["succ"][i] >= 1
                if len(arch) >= np_:
                    arch[int(t["slot"][i])] = x[i].copy()
                else:
                    arch.append(x[i].copy())
                x[i], f[i] = trial, ft
            else:
                assert t["succ"][i] == 0
        xa, fa = h.swarm()
        assert np.array_equal(xa, x) and np.array_equal(fa, f)
