"""Batched small-n CMA-ES engine (bbgpu_cmab_*, SURVEY.md 8(e) row 1 / BASELINE config 5) against the oracle.

Every run of a batch is one reference `Cmaes` (cma/cmaes.cpp) whose whole generation -- sampling, evaluation, ranking,
update, updateSigma, the lazy decomposition, the history rings and converged() -- runs inside one CTA.  Parity tiers as
for the one-handle engine: T2 trajectories under draw replay + eigenbasis replay (state within 1e-10,
ranking bit-exact, sigma / history / stop flags equal), T4 the in-CTA tred2/tql2 against the reference's own basis on the
same matrix, plus free-running behaviour against free-running oracle runs."""
import numpy as np
import pytest

from oracle import oracle as O
from util import assert_close

pytestmark = pytest.mark.gpu

BOX = {O.ROSENBROCK: 10.0, O.ELLIPSOID: 10.0, O.RASTRIGIN: 5.12, O.ACKLEY: 32.768, O.SPHERE: 10.0}


@pytest.fixture(scope="module")
def ctx():
    from bboptpy_b200 import _cabi
    return _cabi.Context(0)


def _objective(ctx, obj, n):
    from bboptpy_b200 import _cabi
    return _cabi.Objective(ctx, n, builtin=obj)


@pytest.mark.parametrize("n", [2, 5, 16, 40, 64])
def test_in_cta_eigensolver_invariants(ctx, n):
    """T4: B orthonormal, B diag(w) B^T = A, ascending, eigenvalues equal LAPACK's."""
    rng = np.random.default_rng(n)
    count = 12
    A = np.empty((count, n, n))
    for c in range(count):
        M = rng.standard_normal((n, n))
        S = M @ M.T / n + np.diag(rng.uniform(0.1, 3.0, n))
        if c % 3 == 0:
            S = np.eye(n) + 1e-3 * (S + S.T)  # nearly isotropic, the early-CMA regime
        if c % 4 == 1:
            S *= 10.0 ** rng.integers(-8, 8)
        A[c] = 0.5 * (S + S.T)
    B, w, info = ctx.eig_small(A)
    assert (info >= 0).all()
    for c in range(count):
        scale = np.abs(A[c]).max()
        assert np.all(np.diff(w[c]) >= 0)
        assert np.abs(B[c].T @ B[c] - np.eye(n)).max() < 1e-13 * n
        assert np.abs(B[c] @ np.diag(w[c]) @ B[c].T - A[c]).max() < 1e-13 * n * scale
        assert_close(w[c], np.linalg.eigvalsh(A[c]), rel=1e-12, what="eigenvalues")


def test_in_cta_eigensolver_matches_the_reference_basis(ctx, port_oracle):
    """Same matrix in, same algorithm (cmaes.cpp:285-478): the basis agrees with the reference's own, sign included,
    whenever the spectrum is separated.  The matrices are the covariance matrices of a real run."""
    P = port_oracle
    n, lam = 12, 24
    P.seed(9)
    h = P.cma(O.VANILLA, 10 ** 9, 1e-12, lam)
    h.init(O.ROSENBROCK, -10 * np.ones(n), 10 * np.ones(n), P.uniform(-10, 10, n))
    Cs, Bs, Ds = [], [], []
    last = 0
    for g in range(150):
        h.iterate()
        if h.geti("eigenlastev") != last and g > 40:
            last = h.geti("eigenlastev")
            Cs.append(h.mat("C").copy())
            Bs.append(h.mat("B").copy())
            Ds.append(h.getv("D").copy())
        last = h.geti("eigenlastev")
    assert len(Cs) > 20
    B, w, info = ctx.eig_small(np.array(Cs))
    assert (info >= 0).all()
    for c in range(len(Cs)):
        assert_close(np.sqrt(w[c]), Ds[c], rel=1e-11, what="D")
        gaps = np.diff(Ds[c] ** 2).min() / (Ds[c] ** 2).max()
        assert np.abs(B[c] - Bs[c]).max() < 1e-12 / max(gaps, 1e-6), (c, gaps)


CASES = [  # (objective, n, lambdas of the runs of one batch, generations, clip to the box)
    (O.RASTRIGIN, 40, [15, 15, 30, 120, 480], 40, False),
    (O.ROSENBROCK, 10, [20, 7, 64], 120, False),
    (O.ACKLEY, 64, [17, 300], 25, False),
    (O.SPHERE, 5, [8, 9, 4096], 12, False),      # odd lambda; the largest class of the ranking network (short: the mean
                                                 # of the lambda = 4096 run collapses onto 0, where 1e-10 RELATIVE tightens fast)
    (O.RASTRIGIN, 12, [10, 40, 300], 40, True),  # samples and mean clipped to the box (cmaes.cpp:74-77, 92-94)
]


@pytest.mark.parametrize("obj,n,lams,gens,bound", CASES)
def test_batch_trajectory_replay(ctx, port_oracle, obj, n, lams, gens, bound):
    """T2 for every run of a batch at once: replayed draws (the oracle's z of that generation) and replayed eigenbases;
    everything the reference holds after a generation is compared, including what the one-handle engine leaves to the
    host (sigma, the history extremes, the stop flag)."""
    from bboptpy_b200 import _cabi
    P = port_oracle
    box = BOX[obj]
    lo, hi = -box * np.ones(n), box * np.ones(n)
    if bound:  # a tight, off-centre box so that clipping is active every generation
        lo, hi = -0.4 * box * np.ones(n), 0.7 * box * np.ones(n)
    R = len(lams)
    P.seed(1234)
    guesses = [P.uniform(lo[0], hi[0], n) for _ in range(R)]
    hs = []
    for r, lam in enumerate(lams):
        h = P.cma(O.VANILLA, 10 ** 9, 1e-12, lam, sigma0=0.3 * box, bound=bound)
        h.init(obj, lo, hi, guesses[r])
        hs.append(h)
    batch = _cabi.CmaBatch(ctx, n, R, bound=bound, lower=lo, upper=hi)
    params = [_cabi.CmaBatch.run_params(n, lam, 10 ** 9, 0.3 * box, 1e-12, seed=r) for r, lam in enumerate(lams)]
    for r, h in enumerate(hs):  # the host derives the same constants as the reference
        for k in ("cc", "cs", "c1", "cmu", "mueff", "damps", "chi", "eigenfreq"):
            assert getattr(params[r], k) == h.getd(k), k
        assert (params[r].hlen, params[r].ik, params[r].mu) == (h.geti("hlen"), h.geti("ik"), h.geti("mu"))
    batch.start(range(R), params, np.array(guesses))
    objective = _objective(ctx, obj, n)
    lastev = [0] * R
    for g in range(gens):
        zs = []
        for h in hs:
            h.iterate()
            zs.append(h.getv("z"))
        batch.step(objective, range(R), max_gens=1, z=np.concatenate(zs), eig_replay=True)
        st = batch.fetch()
        for r, h in enumerate(hs):
            lam = lams[r]
            what = f"gen {g} run {r} (lambda {lam})"
            fired = h.geti("eigenlastev") != lastev[r]
            lastev[r] = h.geti("eigenlastev")
            assert bool(st[r].eig_due) == fired, what + ": the gate fires exactly when the reference's does"
            X, f, rank = batch.population(r, lam)
            assert np.array_equal(rank, h.getv("fit_idx").astype(np.int32)), what + " ranking"
            assert_close(X, h.mat("arx"), what=what + " arx")
            assert_close(batch.get(r, _cabi.F_XMEAN), h.getv("xmean"), what=what + " xmean")
            assert_close(batch.get(r, _cabi.F_PS), h.getv("ps"), what=what + " ps")
            assert_close(batch.get(r, _cabi.F_PC), h.getv("pc"), what=what + " pc")
            low = np.tril_indices(n)
            assert_close(batch.get(r, _cabi.F_C)[low], h.mat("C")[low], what=what + " C")
            assert_close(st[r].sigma, h.getd("sigma"), what=what + " sigma")
            assert (st[r].it, st[r].fev) == (h.geti("it"), h.geti("fev"))
            if h.geti("best_len") == h.geti("hlen"):
                assert_close([st[r].hist_best, st[r].hist_worst], [h.getd("fbest"), h.getd("fworst")], what=what + " history")
            assert list(st[r].ibw) == [int(v) for v in h.getv("ibw")]
            if fired:
                batch.adopt_basis(r, h.mat("B"), h.getv("D"))
                assert_close(batch.get(r, _cabi.F_INVSQRTC), h.mat("invsqrtC"), what=what + " invsqrtC")
            assert (st[r].status != 0) == bool(h.converged()), what + " converged()"
            if st[r].status != 0:
                assert st[r].flag == h.geti("flag")
        if any(s.status != 0 for s in st):
            break


def test_stop_tests_fire_like_the_reference(ctx, port_oracle):
    """Run to the reference's own stop (TolHistFun / TolX / ...) under replay: same generation, same flag, and the
    returned solution / evaluation count of optimize() (base_cmaes.cpp:162-174)."""
    from bboptpy_b200 import _cabi
    P = port_oracle
    n, lam, obj = 8, 12, O.SPHERE
    lo, hi = -10 * np.ones(n), 10 * np.ones(n)
    P.seed(77)
    guess = P.uniform(-10, 10, n)
    h = P.cma(O.VANILLA, 100000, 1e-6, lam)
    h.init(obj, lo, hi, guess)
    batch = _cabi.CmaBatch(ctx, n, 1)
    batch.start([0], [_cabi.CmaBatch.run_params(n, lam, 100000, 2.0, 1e-6, seed=1)], guess[None, :])
    objective = _objective(ctx, obj, n)
    lastev = 0
    for g in range(2000):
        h.iterate()
        batch.step(objective, [0], max_gens=1, z=h.getv("z"), eig_replay=True)
        st = batch.fetch()[0]
        if h.geti("eigenlastev") != lastev:
            lastev = h.geti("eigenlastev")
            batch.adopt_basis(0, h.mat("B"), h.getv("D"))
        stop = bool(h.converged())
        assert (st.status != 0) == stop, g
        if stop:
            break
    assert stop and st.status == 1 and st.flag == h.geti("flag") and st.fev == h.geti("fev")
    x = batch.best(0, 1)[0]
    assert_close(x, h.mat("arx")[int(h.getv("ibw")[0])], rel=1e-8, what="bestSolution")
    assert abs(st.fx_final - P.eval(obj, x)) <= 1e-8


def test_launch_granularity_does_not_change_a_run(ctx):
    """Whole generations per launch: one launch of many generations == many launches of one generation, bit for bit
    (free-running Philox draws, in-CTA decomposition)."""
    from bboptpy_b200 import _cabi
    n, R = 40, 6
    lams = [15, 15, 30, 60, 15, 300]
    objective = _objective(ctx, O.RASTRIGIN, n)
    rng = np.random.default_rng(5)
    guesses = rng.uniform(-5, 5, (R, n))
    out = []
    for mode in range(2):
        batch = _cabi.CmaBatch(ctx, n, R)
        params = [_cabi.CmaBatch.run_params(n, lam, 40 * lam, 2.0, 1e-12, seed=100 + r) for r, lam in enumerate(lams)]
        batch.start(range(R), params, guesses)
        if mode == 0:
            for _ in range(40):
                batch.step(objective, range(R), max_gens=1)
        else:
            batch.step(objective, range(R), quantum=10 ** 9, max_gens=10 ** 6)
        st = batch.fetch()
        assert all(s.it == 40 and s.fev == 40 * lam and s.status in (1, 2) for s, lam in zip(st, lams))
        out.append((np.array([batch.get(r, _cabi.F_C) for r in range(R)]), np.array([batch.get(r, _cabi.F_XMEAN) for r in range(R)]),
                    np.array([s.sigma for s in st]), batch.best()))
        assert all(s.eig_count > 5 for s in st)
    for a, b in zip(out[0], out[1]):
        assert np.array_equal(a, b)


def test_free_running_batch_solves_like_the_reference(ctx, port_oracle):
    """No replay at all: 96 runs on the 10-D Rosenbrock function converge like the reference's own free runs
    (success rate and evaluations to converge), every decomposition done by the in-CTA solver."""
    from bboptpy_b200 import _cabi
    P = port_oracle
    n, lam, R, mfev = 10, 20, 96, 40000
    lo, hi = -10 * np.ones(n), 10 * np.ones(n)
    rng = np.random.default_rng(3)
    guesses = rng.uniform(-10, 10, (R, n))
    batch = _cabi.CmaBatch(ctx, n, R)
    batch.start(range(R), [_cabi.CmaBatch.run_params(n, lam, mfev, 2.0, 1e-9, seed=r) for r in range(R)], guesses)
    objective = _objective(ctx, O.ROSENBROCK, n)
    active = list(range(R))
    while active:
        batch.step(objective, active, quantum=4000, max_gens=10 ** 6)
        st = batch.fetch()
        active = [r for r in active if st[r].status == 0]
    fev = np.array([s.fev for s in st])
    fx = np.array([s.fx_final for s in st])
    assert all(s.status in (1, 2) for s in st)
    xs = batch.best()
    for r in range(0, R, 7):
        assert abs(P.eval(O.ROSENBROCK, xs[r]) - fx[r]) <= 1e-8 * max(1.0, abs(fx[r]))
    ok = fx < 1e-8
    # the reference, free running, same guesses
    fev_o, ok_o = [], []
    P.seed(11)
    for r in range(24):
        h = P.cma(O.VANILLA, mfev, 1e-9, lam)
        x, fe, conv = h.optimize(O.ROSENBROCK, lo, hi, guesses[r])
        fev_o.append(fe)
        ok_o.append(P.eval(O.ROSENBROCK, x) < 1e-8)
    assert ok.mean() >= np.mean(ok_o) - 0.15
    assert abs(np.median(fev[ok]) / np.median(np.array(fev_o)[np.array(ok_o)]) - 1.0) < 0.2


# ---- batches of IPOP / BIPOP runs through the Python API (host/batch.cpp over the batched engine) -----------------

@pytest.fixture(scope="module")
def pkg():
    import bboptpy_b200
    return bboptpy_b200


def _rastrigin(x):
    return 10 * x.size + float(np.sum(x * x - 10 * np.cos(2 * np.pi * x)))


@pytest.mark.parametrize("driver", ["BiPopCMAESBatch", "IPopCMAESBatch"])
def test_restart_batch_free_running(pkg, port_oracle, driver):
    """64 independent restart sequences on the 8-D Rastrigin function: budgets respected, every returned point is what
    it claims to be, and the batch solves the problem as well as the reference's own driver does (same budget)."""
    P = port_oracle
    n, R, mfev = 8, 64, 40000
    lo, hi = -5.12 * np.ones(n), 5.12 * np.ones(n)
    pkg.seed(21)
    guesses = np.random.default_rng(0).uniform(-5, 5, (R, n))
    batch = getattr(pkg, driver)(runs=R, mfev=mfev, tol=1e-8)
    sols = batch.optimize(pkg.Rastrigin, lo, hi, guesses)
    assert len(sols) == R and list(batch.run_ids) == list(range(R))
    st = batch.stats
    fvals = np.array([_rastrigin(s.x) for s in sols])
    assert_close(batch.best_values, fvals, rel=1e-9, what="best f per run")
    evals = np.array([s.n_evals for s in sols])
    assert st.evaluations == evals.sum() and st.base_runs >= 2 * R and st.generations > 0
    # a base run may overshoot its budget by less than one generation (base_cmaes.cpp:166-173), the sequence by one run
    assert (evals >= mfev * 0.5).all() and (evals <= mfev + 10 * n * n + 1).all()
    fb, owner, xb = batch.best_of()
    assert fb == batch.best_values.min() and owner == int(np.argmin(batch.best_values)) and np.array_equal(xb, sols[owner].x)
    # the reference's own driver, free running, same budget: 12 sequences
    rk = 1 if driver.startswith("BiPop") else 0
    ref = []
    P.seed(5)
    for r in range(12):
        base_o = P.cma(O.VANILLA, 1000, 1e-8, 10)
        x_o, fev_o, trace = P.restart_optimize(rk, base_o, mfev, O.RASTRIGIN, lo, hi, guesses[r])
        ref.append(_rastrigin(x_o))
    ref = np.array(ref)
    # Rastrigin's local optima sit at integer multiples of ~0.995 per coordinate: compare the typical depth reached
    assert np.median(fvals) <= np.median(ref) + 2.0, (np.median(fvals), np.median(ref))
    assert (fvals < 30).all()


def test_restart_batch_is_reproducible(pkg):
    n, R = 6, 16
    lo, hi = -5.12 * np.ones(n), 5.12 * np.ones(n)
    guesses = np.random.default_rng(1).uniform(-5, 5, (R, n))
    out = []
    for quantum in (500, 5000):
        pkg.seed(33)
        batch = pkg.BiPopCMAESBatch(runs=R, mfev=15000, tol=1e-8)
        batch.set_quantum(quantum)  # launch granularity must not change any run
        sols = batch.optimize(pkg.Rastrigin, lo, hi, guesses)
        out.append((np.array([s.x for s in sols]), [s.n_evals for s in sols]))
    assert np.array_equal(out[0][0], out[1][0]) and out[0][1] == out[1][1]


# ---- objectives outside the registry: one evaluate call per generation over the candidates of all listed runs ---------

def test_external_objective_equals_the_fused_kernel(ctx):
    """A batched host callback that computes the registry objective's own values must reproduce the fused path bit for
    bit (same draws, same fitness bits -> same ranking, same state), generation by generation and as a multi-generation
    step; the callback sees each generation's candidates of ALL runs in one call."""
    from bboptpy_b200 import _cabi
    n, lams = 12, [9, 20, 300]
    R = len(lams)
    reg = _objective(ctx, O.RASTRIGIN, n)
    calls = []

    def cb(X):
        calls.append(X.shape[0])
        return reg.eval(X)

    ext = _cabi.Objective.from_python(ctx, n, cb, batched=True)
    guesses = np.random.default_rng(8).uniform(-5, 5, (R, n))
    out = []
    for objective in (reg, ext):
        batch = _cabi.CmaBatch(ctx, n, R)
        batch.start(range(R), [_cabi.CmaBatch.run_params(n, lam, 30 * lam, 2.0, 1e-12, seed=50 + r) for r, lam in enumerate(lams)], guesses)
        for _ in range(5):
            batch.step(objective, range(R), max_gens=1)
        batch.step(objective, range(R), quantum=10 ** 9, max_gens=10 ** 6)  # to the end of the budgets: 30 generations each
        st = batch.fetch()
        assert all(s.it == 30 and s.fev == 30 * lam and s.status in (1, 2) for s, lam in zip(st, lams))
        out.append((np.array([batch.get(r, _cabi.F_C) for r in range(R)]), np.array([batch.get(r, _cabi.F_XMEAN) for r in range(R)]),
                    np.array([s.sigma for s in st]), batch.best(), np.array([s.fx_final for s in st]),
                    [s.generations_this_step for s in st]))
    for a, b in zip(out[0][:5], out[1][:5]):
        assert np.array_equal(a, b)
    assert out[1][5] == [25] * R  # generations_this_step accumulates over the launches of one step call
    assert calls[:5] == [sum(lams)] * 5 and len(calls) == 30


def test_restart_batch_with_a_python_callable(pkg):
    """The reference's own calling convention (a Python callable f(x) -> float) through the batched driver."""
    n, R = 5, 6
    lo, hi = -5.12 * np.ones(n), 5.12 * np.ones(n)
    guesses = np.random.default_rng(2).uniform(-5, 5, (R, n))
    evals = [0]

    def f(x):
        evals[0] += 1
        return float(np.sum(x * x))

    pkg.seed(4)
    batch = pkg.IPopCMAESBatch(runs=R, mfev=3000, tol=1e-10)
    sols = batch.optimize(f, lo, hi, guesses)
    assert len(sols) == R and all(float(np.sum(s.x ** 2)) < 1e-8 for s in sols)
    # the engine calls f once per candidate; the reference counts one more evaluation per base run (ipop_cmaes.cpp:94)
    assert evals[0] + batch.stats.base_runs == batch.stats.evaluations
