"""GPU parity of the standalone kernels, through the C ABI (ctypes on libbbgpu.so)."""
import numpy as np
import pytest

from bboptpy_b200 import _cabi
from oracle import oracle as O
from util import assert_close

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    return _cabi.Context(0)


@pytest.mark.parametrize("obj", range(5))
@pytest.mark.parametrize("n", [1, 2, 10, 31, 32, 33, 100, 1000, 4099])
def test_objective_bits_match_oracle(ctx, port_oracle, obj, n):
    rng = np.random.default_rng(100 * obj + n)
    scale = {0: 10.0, 1: 10.0, 2: 10.0, 3: 5.12, 4: 32.768}[obj]
    X = rng.uniform(-scale, scale, size=(67, n))
    X[0] = 0.0
    X[1] = np.round(X[1])  # integers and half-integers stress the cos(2 pi x) folding
    X[2] = np.round(X[2] * 2) / 2
    X[3] = np.round(X[3] * 4) / 4
    X[4] = np.round(X[4] * 8) / 8
    o = _cabi.Objective(ctx, n, builtin=obj)
    got = o.eval(X)
    want = port_oracle.eval_rows(obj, X)
    assert np.array_equal(got, want), np.max(np.abs(got - want))


def test_objective_full_size_linearity_properties(ctx):
    # size-independent properties at the BASELINE sizes: sphere(a x) = a^2 sphere(x) for a power of two (exact),
    # rastrigin/ackley at the origin, rosenbrock at the all-ones optimum
    n, lam = 1000, 4096
    rng = np.random.default_rng(5)
    X = rng.uniform(-5, 5, size=(lam, n))
    sp = _cabi.Objective(ctx, n, builtin=_cabi.OBJ_SPHERE)
    assert np.array_equal(sp.eval(4.0 * X), 16.0 * sp.eval(X))
    assert np.array_equal(sp.eval(-X), sp.eval(X))
    assert _cabi.Objective(ctx, n, builtin=_cabi.OBJ_RASTRIGIN).eval(np.zeros((1, n)))[0] == 0.0
    assert abs(_cabi.Objective(ctx, n, builtin=_cabi.OBJ_ACKLEY).eval(np.zeros((1, n)))[0]) < 1e-15
    assert _cabi.Objective(ctx, n, builtin=_cabi.OBJ_ROSENBROCK).eval(np.ones((3, n))).tolist() == [0.0, 0.0, 0.0]


def test_host_callback_objective(ctx):
    n = 7
    o = _cabi.Objective.from_python(ctx, n, lambda x: float(np.sum(x * x)))
    X = np.arange(21, dtype=np.float64).reshape(3, n)
    assert np.allclose(o.eval(X), (X * X).sum(1))


@pytest.mark.parametrize("M,N,K", [(1, 1, 1), (8, 8, 4), (20, 10, 10), (129, 113, 17), (130, 257, 33), (256, 1000, 1000),
                                   (4096, 1000, 1000), (300, 40, 40)])
def test_dgemm_matches_numpy(ctx, M, N, K):
    rng = np.random.default_rng(M + N + K)
    A, B = rng.standard_normal((M, K)), rng.standard_normal((K, N))
    C = ctx.dgemm(A, B)
    want = A @ B
    # fp64 accumulate in a different order than numpy: 1e-13 of the row/column scale is ample
    assert np.max(np.abs(C - want)) <= 1e-13 * np.sqrt(K) * max(1.0, np.abs(want).max())


@pytest.mark.parametrize("n,K", [(1, 1), (10, 10), (33, 7), (129, 64), (1000, 2048), (40, 7)])
def test_dsyrk_matches_numpy(ctx, n, K):
    rng = np.random.default_rng(n + K)
    Y, coef = rng.standard_normal((K, n)), rng.uniform(-1, 1, K)
    C0 = rng.standard_normal((n, n))
    C0 = C0 + C0.T
    alpha = 0.7
    C = ctx.dsyrk(Y, coef, C0, alpha)
    want = alpha * C0 + (Y.T * coef) @ Y
    lo = np.tril_indices(n)
    assert np.max(np.abs(C[lo] - want[lo])) <= 1e-13 * np.sqrt(K) * max(1.0, np.abs(want).max())
    up = np.triu_indices(n, 1)
    assert np.array_equal(C[up], C0[up])  # the strict upper triangle is never written (cma/cmaes.cpp:127)


@pytest.mark.parametrize("count", [2, 3, 20, 1000, 4096, 8191, 8192, 8193, 10000, 16384, 20000, 65536])
def test_rank_is_the_sorted_permutation(ctx, count):
    rng = np.random.default_rng(count)
    f = rng.standard_normal(count) * 10.0 ** rng.integers(-3, 4, count)
    f[rng.integers(0, count)] = -0.0
    idx = ctx.rank(f)
    assert sorted(idx.tolist()) == list(range(count))
    assert np.array_equal(f[idx], np.sort(f))
    assert np.array_equal(idx, np.argsort(f, kind="stable"))  # tie-free input -> unique permutation


def test_rank_orders_ties_by_index_and_nan_last(ctx):
    f = np.array([3.0, 1.0, 3.0, np.nan, 1.0, -np.inf, np.inf, 2.0])
    idx = ctx.rank(f)
    assert idx.tolist() == [5, 1, 4, 7, 0, 2, 6, 3]
    # a NaN with the sign bit set (x86's default NaN, e.g. inf - inf in a host callback) must not become the best
    # candidate, and -0.0 / +0.0 are one key (ties by index)
    neg_nan = np.frombuffer(np.uint64(0xFFF8000000000000).tobytes(), dtype=np.float64)[0]
    f = np.array([0.0, neg_nan, -0.0, -1.0, np.nan, 5.0])
    assert ctx.rank(f).tolist() == [3, 0, 2, 5, 1, 4]


@pytest.mark.parametrize("count", [8192, 9000, 16384])
def test_rank_chunk_path_with_heavy_ties_and_nans(ctx, count):
    """The chunk-sort path (count >= 8192): many equal keys across chunks (ties by index), both zeros, NaNs of both signs."""
    rng = np.random.default_rng(count)
    f = rng.integers(0, 7, count).astype(np.float64)
    f[rng.integers(0, count, 20)] = np.nan
    f[rng.integers(0, count, 20)] = -0.0
    neg_nan = np.frombuffer(np.uint64(0xFFF8000000000000).tobytes(), dtype=np.float64)[0]
    f[rng.integers(0, count, 5)] = neg_nan
    idx = ctx.rank(f)
    key = np.where(np.isnan(f), np.inf, f + 0.0)
    want = np.lexsort((np.arange(count), np.isnan(f), key))  # value, NaN last, then index
    assert np.array_equal(idx, want)


def test_philox_normals_are_standard_normal_and_placement_invariant(ctx):
    z = ctx.philox_normal(1234, 7, 512, 1001)
    assert np.isfinite(z).all()
    assert abs(z.mean()) < 5e-3 and abs(z.std() - 1.0) < 5e-3
    k = ((z - z.mean()) ** 4).mean() / z.var() ** 2
    assert abs(k - 3.0) < 0.05
    # Kolmogorov-Smirnov against N(0,1), the mass in the far tail, symmetry, and no correlation between the two normals of
    # a Philox block
    from scipy import stats
    big = ctx.philox_normal(99, 3, 2048, 2000)
    assert stats.kstest(big.ravel(), "norm").pvalue > 1e-4
    R = 3.654152885361009
    tail = (np.abs(big) > R).mean()
    assert abs(tail - 2 * stats.norm.sf(R)) < 4 * np.sqrt(2 * stats.norm.sf(R) / big.size)
    assert abs((big > 0).mean() - 0.5) < 2e-3
    assert abs(np.corrcoef(big[:, 0::2].ravel(), big[:, 1::2].ravel())[0, 1]) < 3e-3
    # same (seed, generation) -> same draws; different generation -> different draws
    assert np.array_equal(z, ctx.philox_normal(1234, 7, 512, 1001))
    assert not np.array_equal(z, ctx.philox_normal(1234, 8, 512, 1001))
    # the first rows do not depend on how many rows are drawn (candidate-keyed counters)
    assert np.array_equal(z[:64], ctx.philox_normal(1234, 7, 64, 1001))


@pytest.mark.parametrize("n", [3, 10, 149, 300, 600, 1000, 1024, 1025, 1036, 1700])
def test_cooperative_sytrd_matches_lapack(ctx, n):
    """The persistent-kernel tridiagonalisation against LAPACK dsytrd('L') on the same matrix: same convention, so d, e, tau
    agree to rounding, and the reflectors reproduce Q^T A Q = T."""
    from scipy.linalg import lapack
    rng = np.random.default_rng(n)
    M = rng.standard_normal((n, n))
    A = np.eye(n) + 1e-2 * (M + M.T) + np.diag(rng.uniform(0, 1, n))  # a CMA-like covariance: clustered spectrum
    d, e, tau, V, ms = ctx.sytrd(A)
    c, d0, e0, tau0, info = lapack.dsytrd(A, lower=1)
    assert info == 0
    scale = np.abs(A).max()
    assert np.abs(d - d0).max() <= 1e-11 * scale and np.abs(np.abs(e) - np.abs(e0)).max() <= 1e-11 * scale
    # Q = H_0 H_1 ... ; H_i = I - tau_i v_i v_i^T, v_i = (0,...,0, 1, V[i+2:, i])
    Q = np.eye(n)
    for i in range(n - 1):
        v = np.zeros(n)
        v[i + 1] = 1.0
        v[i + 2:] = V[i + 2:, i]
        Q = Q - tau[i] * np.outer(Q @ v, v)
    T = np.diag(d) + np.diag(e, 1) + np.diag(e, -1)
    assert np.abs(Q.T @ Q - np.eye(n)).max() <= 1e-12 * n
    assert np.abs(Q.T @ A @ Q - T).max() <= 1e-12 * n * scale


@pytest.mark.parametrize("n,kind", [(5, 1), (64, 1), (65, 0), (200, 0), (200, 3), (333, 2), (1000, 0), (1000, 1), (1000, 4), (257, 1), (777, 3), (1700, 2)])
def test_tridiagonal_divide_and_conquer(ctx, n, kind):
    """T = Z diag(w) Z^T on the device against LAPACK: eigenvalues, orthogonality and residual at working precision, also
    for the spectra that stress deflation (nearly isotropic early-CMA covariance, graded, exactly degenerate)."""
    from scipy.linalg import lapack
    rng = np.random.default_rng(100 * n + kind)
    M = rng.standard_normal((n, n))
    if kind == 0:
        A = np.eye(n) + 7e-5 * (M + M.T)
    elif kind == 1:
        A = np.eye(n) + 1e-2 * (M + M.T) + np.diag(rng.uniform(0, 1, n))
    elif kind == 2:
        A = M @ M.T / n
    elif kind == 3:
        A = np.diag(10.0 ** rng.uniform(-6, 6, n)) + 1e-3 * (M + M.T)
    else:
        A = np.eye(n)
    _, d, e, _, info = lapack.dsytrd(A, lower=1)
    assert info == 0
    w, Z, ms = ctx.stedc(d, e)
    T = np.diag(d) + np.diag(e, 1) + np.diag(e, -1)
    scale = np.abs(T).max()
    assert np.all(np.diff(w) >= 0)
    assert np.abs(w - np.linalg.eigvalsh(T)).max() <= 1e-13 * scale * max(1, n / 100)
    assert np.abs(Z.T @ Z - np.eye(n)).max() <= 1e-13 * max(1, n / 100)
    assert np.abs(T @ Z - Z * w).max() <= 1e-13 * scale * max(1, n / 100)
