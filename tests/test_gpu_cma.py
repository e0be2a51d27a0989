"""Teacher-forced per-generation parity of the CMA-ES device phases against the oracle
(tier T1), through the C ABI.

Each generation: load the oracle's state -> replay the oracle's own N(0,1) draws -> compare
X (1e-10), fitness bits and ranking (bit-exact), then m, p_sigma, p_c, C (1e-10), sigma, and the
eigen path (own solver: invariants; replayed basis: C^-1/2 at 1e-10)."""
import numpy as np
import pytest

from bboptpy_b200 import _cabi
from oracle import oracle as O
from util import REL, assert_close, oracle_constants, update_sigma_host

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    return _cabi.Context(0)


def load_state(g, s, kind):
    g.set(_cabi.F_XMEAN, s["xmean"])
    g.set(_cabi.F_PC, s["pc"])
    g.set(_cabi.F_PS, s["ps"])
    g.set(_cabi.F_C, s["C"])
    if kind != O.SEP:
        g.set(_cabi.F_INVSQRTC, s["invsqrtC"])
        g.set(_cabi.F_D, s["D"])
        g.set(_cabi.F_B, s["B"])
    else:
        g.set(_cabi.F_D, s["D"])


CASES = [
    # kind, obj, n, lambda, gens, bound, box
    (O.VANILLA, O.ROSENBROCK, 10, 20, 40, False, 10.0),
    (O.ACTIVE, O.ROSENBROCK, 10, 20, 40, False, 10.0),
    (O.SEP, O.ELLIPSOID, 10, 20, 40, False, 10.0),
    (O.VANILLA, O.RASTRIGIN, 40, 15, 25, False, 5.12),
    (O.ACTIVE, O.ACKLEY, 33, 64, 15, True, 32.768),
    (O.SEP, O.SPHERE, 100, 32, 15, True, 10.0),
    (O.VANILLA, O.ROSENBROCK, 200, 300, 4, False, 10.0),
    (O.ACTIVE, O.RASTRIGIN, 130, 257, 4, True, 5.12),
    (O.SEP, O.ELLIPSOID, 1000, 512, 4, False, 10.0),
    # the BASELINE shapes bench.py times (configs A and S): the 128x112 sampling tile over 288 CTAs, the 64x64 SYRK at
    # K = 2048 (4096 signed rows for Active), the ROWSQ norm GEMM, the own eigensolver at n = 1000 inside a real
    # generation (the reference decomposes after EVERY generation here) and sep_sample at n = 10^4 -- against the oracle
    (O.VANILLA, O.ROSENBROCK, 1000, 4096, 2, False, 10.0),
    (O.ACTIVE, O.ROSENBROCK, 1000, 4096, 2, False, 10.0),
    (O.VANILLA, O.ROSENBROCK, 1000, 4096, 1, True, 10.0),
    (O.SEP, O.ELLIPSOID, 10000, 4096, 2, False, 10.0),
]


@pytest.mark.parametrize("kind,obj,n,lam,gens,bound,box", CASES)
def test_teacher_forced_generations(ctx, port_oracle, kind, obj, n, lam, gens, bound, box):
    P = port_oracle
    P.seed(42)
    lo, hi = -box * np.ones(n), box * np.ones(n)
    guess = P.uniform(-box, box, n)
    h = P.cma(kind, 10 ** 9, 1e-12, lam, bound=bound)
    h.init(obj, lo, hi, guess)
    consts = oracle_constants(h, kind)
    damps = h.getd("damps")
    g = _cabi.Cma(ctx, kind, n, lam, h.getv("weights"), consts, bound=bound, lower=lo, upper=hi)
    g.reset(guess)
    objective = _cabi.Objective(ctx, n, builtin=obj)
    ik, hlen = h.geti("ik"), h.geti("hlen")
    for gen in range(gens):
        s0 = h.snapshot()
        load_state(g, s0, kind)
        # ---- sample with replayed draws
        h.sample()
        z = h.getv("z")
        g.sample(s0["sigma"], gen, z=z)
        X = h.mat("arx")
        assert_close(g.get(_cabi.F_ARX), X, what=f"gen {gen} arx")
        assert np.array_equal(g.get(_cabi.F_Z).ravel(), z)
        # ---- evaluate + rank on the oracle's X (teacher forced): bit-exact
        g.set(_cabi.F_ARX, X)
        h.evalsort()
        g.evaluate(objective)
        g.rank(ik)
        fit_sorted, fit_idx = h.getv("fit_val"), h.getv("fit_idx").astype(np.int32)
        assert len(set(fit_sorted.tolist())) == lam, "parity cases must be tie-free"
        fit = g.get(_cabi.F_FITNESS)
        assert np.array_equal(fit[fit_idx], fit_sorted), f"gen {gen} fitness bits"
        assert np.array_equal(g.rank_idx(), fit_idx), f"gen {gen} ranking"
        assert np.array_equal(g.get(_cabi.F_RANKVAL), fit_sorted)
        # ---- update
        fev_after = h.geti("fev")
        eig_before = h.geti("eigenlastev") if kind != O.SEP else 0
        fworst, fbest, it = h.getd("fworst"), h.getd("fbest"), h.geti("it")
        h.update()
        g.update(s0["sigma"], fev_after)
        rep = g.fetch()
        s1 = h.snapshot()
        assert list(rep.ibw) == [int(v) for v in h.getv("ibw")]
        assert list(rep.ybw) == h.getv("ybw").tolist()
        assert rep.f_kth == fit_sorted[ik]
        assert_close(g.get(_cabi.F_XMEAN), s1["xmean"], what=f"gen {gen} xmean")
        assert_close(g.get(_cabi.F_XOLD), s1["xold"], what=f"gen {gen} xold")
        assert_close(g.get(_cabi.F_PS), s1["ps"], what=f"gen {gen} ps")
        assert_close(g.get(_cabi.F_PC), s1["pc"], what=f"gen {gen} pc")
        assert_close(rep.pslen, np.linalg.norm(s1["ps"]), what="pslen")
        sigma = update_sigma_host(s0["sigma"], rep.pslen, consts["cs"], damps, consts["chi"], fit_sorted[0],
                                  fit_sorted[ik], it, hlen, fworst, fbest)
        assert_close(sigma, s1["sigma"], rel=1e-12, what=f"gen {gen} sigma")
        if kind == O.SEP:
            assert_close(g.get(_cabi.F_C), s1["C"], what=f"gen {gen} C")
            assert_close(g.get(_cabi.F_D), s1["D"], what=f"gen {gen} D")
        else:
            lower = np.tril_indices(n)
            Cg = g.get(_cabi.F_C)
            eig_fired = s1["eigenlastev"] != eig_before
            if kind == O.ACTIVE:
                assert_close(g.get(_cabi.F_YCOEFF), s1["ycoeff"], rel=1e-9, what=f"gen {gen} ycoeff")
            if not eig_fired:
                assert_close(Cg[lower], s1["C"][lower], what=f"gen {gen} C")
            else:
                # the decomposition may shift the diagonal (cmaes.cpp:249-266); compare after our own eig
                g.eig()
                rep = g.fetch()
                assert rep.eig_info == 0
                Cg = g.get(_cabi.F_C)
                assert_close(Cg[lower], s1["C"][lower], what=f"gen {gen} C after eig")
                Bg, Dg = g.get(_cabi.F_B), g.get(_cabi.F_D)
                Cfull = np.tril(Cg) + np.tril(Cg, -1).T
                cs = np.abs(Cfull).max()
                assert np.abs(Bg.T @ Bg - np.eye(n)).max() < 1e-12 * n
                assert np.abs((Bg * Dg ** 2) @ Bg.T - Cfull).max() < 1e-12 * n * cs
                assert np.all(np.diff(Dg) >= 0)
                assert_close(Dg ** 2, s1["D"] ** 2, rel=1e-10, what=f"gen {gen} eigenvalues")
                assert rep.d_min == Dg[0] and rep.d_max == Dg[-1]
                # C^-1/2 is basis independent, but ill-conditioned by the small eigenvalues
                cond = (Dg[-1] / Dg[0]) ** 2
                assert_close(g.get(_cabi.F_INVSQRTC), s1["invsqrtC"], rel=1e-12 * n * max(cond, 1.0),
                             what=f"gen {gen} invsqrtC (own basis)")
                # eigenbasis replay: adopt the oracle's (B, D)
                g.eig(s1["B"], s1["D"])
                assert_close(g.get(_cabi.F_INVSQRTC), s1["invsqrtC"], what=f"gen {gen} invsqrtC (replayed basis)")
                assert np.array_equal(g.get(_cabi.F_B), s1["B"])
        # ---- convergence probe quantities
        iaxis = n - 1 - (it % n)  # converged() runs after _it++ (base_cmaes.cpp:155): n-1-((it+1-1) % n)
        g.probe(s1["sigma"], iaxis)
        rep = g.fetch()
        spread = s1["D"] if kind == O.SEP else np.sqrt(np.diag(g.get(_cabi.F_C)))
        assert_close(rep.tolx_max, np.max(np.maximum(g.get(_cabi.F_PC), spread)), rel=1e-14, what="tolx")
        h.history()


def test_philox_generation_is_a_valid_cma_step(ctx, port_oracle):
    """Free-running mode (device Philox): one generation from the identity state moves the mean
    toward lower fitness and keeps C symmetric positive definite on its lower triangle."""
    n, lam = 64, 128
    P = port_oracle
    h = P.cma(O.VANILLA, 10 ** 9, 1e-12, lam)
    guess = 3.0 * np.ones(n)
    h.init(O.SPHERE, -10 * np.ones(n), 10 * np.ones(n), guess)
    g = _cabi.Cma(ctx, O.VANILLA, n, lam, h.getv("weights"), oracle_constants(h, O.VANILLA), seed=99)
    g.reset(guess)
    obj = _cabi.Objective(ctx, n, builtin=O.SPHERE)
    sigma, fev = 2.0, 0
    f_first = None
    for gen in range(30):
        fev += lam
        g.generation(obj, sigma, gen, h.geti("ik"), fev)
        rep = g.fetch()
        f_first = rep.ybw[0] if f_first is None else f_first
        sigma = update_sigma_host(sigma, rep.pslen, h.getd("cs"), h.getd("damps"), h.getd("chi"), rep.ybw[0], rep.f_kth,
                                  gen, 10 ** 9, 0.0, 1.0)
        g.eig()
    rep = g.fetch()
    assert rep.ybw[0] < 0.45 * f_first  # the oracle reaches 0.20-0.30 on this configuration
    C = g.get(_cabi.F_C)
    Cf = np.tril(C) + np.tril(C, -1).T
    assert np.linalg.eigvalsh(Cf).min() > 0


@pytest.mark.parametrize("obj,n,lam,bound", [(O.ELLIPSOID, 1000, 64, False), (O.RASTRIGIN, 77, 40, True),
                                             (O.ACKLEY, 10, 16, True), (O.SPHERE, 129, 24, False), (O.ROSENBROCK, 50, 16, False)])
def test_sep_fused_generation_equals_phases(ctx, obj, n, lam, bound):
    """SepCmaes free-running: the fused Philox -> sample -> evaluate pass (bbgpu_cma_generation) writes the same X and
    the same fitness bits as the separate phases, and the generation ends in the same state (Rosenbrock: unfused path)."""
    from bboptpy_b200.strategy import cma_constants
    w, c = cma_constants(2, n, lam)
    box = 5.0
    lo, hi = -box * np.ones(n), box * np.ones(n)
    guess = np.linspace(-4.0, 4.5, n)
    objective = _cabi.Objective(ctx, n, builtin=obj)
    a = _cabi.Cma(ctx, _cabi.CMA_SEP, n, lam, w, c, bound=bound, lower=lo, upper=hi, seed=31)
    b = _cabi.Cma(ctx, _cabi.CMA_SEP, n, lam, w, c, bound=bound, lower=lo, upper=hi, seed=31)
    a.reset(guess)
    b.reset(guess)
    sigma, fev = 3.0, 0
    for gen in range(4):
        fev += lam
        a.sample(sigma, gen)
        a.evaluate(objective)
        a.rank(c["ik"])
        a.update(sigma, fev)
        b.generation(objective, sigma, gen, c["ik"], fev)
        ra, rb = a.fetch(), b.fetch()
        assert np.array_equal(a.get(_cabi.F_ARX), b.get(_cabi.F_ARX)), f"gen {gen} X"
        assert np.array_equal(a.get(_cabi.F_FITNESS), b.get(_cabi.F_FITNESS)), f"gen {gen} fitness bits"
        assert np.array_equal(a.rank_idx(), b.rank_idx())
        for f in (_cabi.F_XMEAN, _cabi.F_C, _cabi.F_D, _cabi.F_PS, _cabi.F_PC):
            assert np.array_equal(a.get(f), b.get(f)), f"gen {gen} field {f}"
        assert ra.pslen == rb.pslen and list(ra.ibw) == list(rb.ibw)
        if bound:
            X = b.get(_cabi.F_ARX)
            assert (X >= -box).all() and (X <= box).all()
        sigma *= 0.9


@pytest.mark.parametrize("n,cond", [(256, 1e2), (300, 1e4), (777, 1e6), (1000, 1.0 + 1e-3), (1000, 1e4), (1500, 1e3)])
def test_own_eigensolver_inside_the_handle(ctx, port_oracle, n, cond):
    """`bbgpu_cma_eig` at n >= 256 runs our own decomposition (cooperative sytrd + tridiagonal divide & conquer + ormtr,
    cmaes.cpp:229-283 is the contract): C = B diag(D^2) B^T with B orthogonal, D ascending, C^-1/2 C C^-1/2 = I, on
    covariance matrices with a prescribed spectrum (nearly isotropic like early CMA generations, and graded)."""
    P = port_oracle
    lam = 16
    h = P.cma(O.VANILLA, 10 ** 9, 1e-12, lam)
    h.init(O.SPHERE, -10 * np.ones(n), 10 * np.ones(n), np.ones(n))
    g = _cabi.Cma(ctx, O.VANILLA, n, lam, h.getv("weights"), oracle_constants(h, O.VANILLA), seed=1)
    g.reset(np.ones(n))
    rng = np.random.default_rng(n)
    Q, _ = np.linalg.qr(rng.standard_normal((n, n)))
    spectrum = np.sort(np.exp(rng.uniform(0.0, np.log(cond), n)))
    C = (Q * spectrum) @ Q.T
    C = 0.5 * (C + C.T)
    g.set(_cabi.F_C, C)
    g.eig()
    rep = g.fetch()
    assert rep.eig_info == 0
    B, D, inv = g.get(_cabi.F_B), g.get(_cabi.F_D), g.get(_cabi.F_INVSQRTC)
    assert np.all(np.diff(D) >= 0)
    assert_close(D ** 2, np.linalg.eigvalsh(C), rel=1e-10, what="eigenvalues")
    assert np.abs(B.T @ B - np.eye(n)).max() < 1e-12
    assert np.abs((B * D ** 2) @ B.T - C).max() < 1e-10 * spectrum[-1]
    invf = np.tril(inv) + np.tril(inv, -1).T
    assert np.abs(invf @ C @ invf - np.eye(n)).max() < 1e-9 * cond
