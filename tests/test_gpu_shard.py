"""Candidate-sharded CMA-ES run: G handles that each own lambda/G candidates
must reproduce the unsharded run -- and, teacher-forced, the oracle.

On one GPU the G shards are G handles on the same device and the two exchange steps (fitness
all-gather, partial all-reduce) are done by the test through the C ABI's exchange fields; the NCCL
transport itself is covered by tests/multi_gpu_shard.py (needs >= 2 GPUs, run by bench/driver calls)."""
import numpy as np
import pytest

from bboptpy_b200 import _cabi
from oracle import oracle as O
from util import assert_close, oracle_constants

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    return _cabi.Context(0)


def exchange_fitness(shards, active):
    fit = np.concatenate([s.get(_cabi.F_FITNESS)[s.shard_rank * s.lam_loc:(s.shard_rank + 1) * s.lam_loc] for s in shards])
    for s in shards:
        s.set(_cabi.F_FITNESS, fit)
    if active:
        nrm = np.concatenate([s.get(_cabi.F_NORMS)[s.shard_rank * s.lam_loc:(s.shard_rank + 1) * s.lam_loc] for s in shards])
        for s in shards:
            s.set(_cabi.F_NORMS, nrm)
    return fit


def exchange_partials(shards):
    total = np.zeros_like(shards[0].get(_cabi.F_SHARD_RED))
    for s in shards:
        total += s.get(_cabi.F_SHARD_RED)
    for s in shards:
        s.set(_cabi.F_SHARD_RED, total)


def load_state(g, s):
    g.set(_cabi.F_XMEAN, s["xmean"])
    g.set(_cabi.F_PC, s["pc"])
    g.set(_cabi.F_PS, s["ps"])
    g.set(_cabi.F_C, s["C"])
    g.set(_cabi.F_INVSQRTC, s["invsqrtC"])
    g.set(_cabi.F_D, s["D"])
    g.set(_cabi.F_B, s["B"])


@pytest.mark.parametrize("kind,obj,n,lam,G,bound,gens", [
    (O.VANILLA, O.ROSENBROCK, 24, 48, 2, False, 6),
    (O.ACTIVE, O.ROSENBROCK, 24, 48, 4, False, 6),
    (O.ACTIVE, O.RASTRIGIN, 130, 256, 8, True, 6),
    (O.VANILLA, O.ACKLEY, 200, 320, 4, True, 6),
    # BASELINE config A8 (n = 1000, lambda = 16384 over 8 GPUs): the shapes bench.py's sharded_run times, 8 emulated shards
    (O.VANILLA, O.ROSENBROCK, 1000, 16384, 8, False, 1),
    (O.ACTIVE, O.ROSENBROCK, 1000, 16384, 8, False, 1),
])
def test_sharded_generation_matches_oracle(ctx, port_oracle, kind, obj, n, lam, G, bound, gens):
    """Teacher-forced against the oracle: replayed draws, oracle state in, sharded phases, exchange, state out."""
    P = port_oracle
    P.seed(11)
    box = 5.0
    lo, hi = -box * np.ones(n), box * np.ones(n)
    guess = P.uniform(-box, box, n)
    h = P.cma(kind, 10 ** 9, 1e-12, lam, bound=bound)
    h.init(obj, lo, hi, guess)
    consts = oracle_constants(h, kind)
    objective = _cabi.Objective(ctx, n, builtin=obj)
    shards = [_cabi.Cma(ctx, kind, n, lam, h.getv("weights"), consts, bound=bound, lower=lo, upper=hi, shard=(r, G))
              for r in range(G)]
    for s in shards:
        s.reset(guess)
    ik = h.geti("ik")
    ll = lam // G
    for gen in range(gens):
        s0 = h.snapshot()
        h.sample()
        z = h.getv("z")
        X = h.mat("arx")
        for r, s in enumerate(shards):
            load_state(s, s0)
            s.sample(s0["sigma"], gen, z=z)           # every shard consumes its own rows of the draw stream
            assert_close(s.get(_cabi.F_ARX), X[r * ll:(r + 1) * ll], what=f"gen {gen} shard {r} arx")
            s.set(_cabi.F_ARX, X[r * ll:(r + 1) * ll])  # teacher forced
            s.evaluate(objective)
        h.evalsort()
        fit = exchange_fitness(shards, kind == O.ACTIVE)
        fit_idx = h.getv("fit_idx").astype(np.int32)
        for s in shards:
            s.rank(ik)
            assert np.array_equal(s.rank_idx(), fit_idx), f"gen {gen} ranking"
        assert np.array_equal(fit[fit_idx], h.getv("fit_val"))
        for s in shards:
            s.shard_partial()
        exchange_partials(shards)
        eig_before = h.geti("eigenlastev")
        h.update()
        eig_fired = h.geti("eigenlastev") != eig_before
        lower = np.tril_indices(n)
        for r, s in enumerate(shards):
            s.shard_finish(s0["sigma"], h.geti("fev"))
            if eig_fired:  # the decomposition may shift the diagonal of C (cmaes.cpp:249-266)
                s.eig()
            rep = s.fetch()
            assert_close(s.get(_cabi.F_XMEAN), h.getv("xmean"), what=f"gen {gen} shard {r} xmean")
            assert_close(s.get(_cabi.F_PS), h.getv("ps"), what=f"gen {gen} ps")
            assert_close(s.get(_cabi.F_PC), h.getv("pc"), what=f"gen {gen} pc")
            assert_close(s.get(_cabi.F_C)[lower], h.mat("C")[lower], what=f"gen {gen} C")
            if kind == O.ACTIVE:
                assert_close(s.get(_cabi.F_YCOEFF), h.getv("ycoeff"), what=f"gen {gen} ycoeff")
            assert list(rep.ibw) == [int(v) for v in h.getv("ibw")]
        h.history()


@pytest.mark.parametrize("kind", [O.VANILLA, O.ACTIVE])
def test_sharded_philox_run_equals_unsharded(ctx, port_oracle, kind):
    """Free running: Philox counters carry the global candidate id, so G shards draw, evaluate and rank exactly what
    the unsharded handle does; the update agrees to 1e-10 (different summation order)."""
    P = port_oracle
    P.seed(5)
    n, lam, G = 64, 128, 4
    lo, hi = -10.0 * np.ones(n), 10.0 * np.ones(n)
    guess = P.uniform(-10, 10, n)
    h = P.cma(kind, 10 ** 9, 1e-12, lam)
    h.init(O.ROSENBROCK, lo, hi, guess)
    consts = oracle_constants(h, kind)
    w = h.getv("weights")
    objective = _cabi.Objective(ctx, n, builtin=O.ROSENBROCK)
    one = _cabi.Cma(ctx, kind, n, lam, w, consts, seed=99)
    shards = [_cabi.Cma(ctx, kind, n, lam, w, consts, seed=99, shard=(r, G)) for r in range(G)]
    for g in [one] + shards:
        g.reset(guess)
    sigma, ik, fev, ll = 2.0, h.geti("ik"), 0, lam // G
    lower = np.tril_indices(n)
    for gen in range(5):
        fev += lam
        one.sample(sigma, gen)
        one.evaluate(objective)
        one.rank(ik)
        one.update(sigma, fev)
        X = one.get(_cabi.F_ARX)
        for r, s in enumerate(shards):
            s.sample(sigma, gen)
            assert np.array_equal(s.get(_cabi.F_ARX), X[r * ll:(r + 1) * ll]), "same draws, same GEMM rows"
            s.evaluate(objective)
        fit = exchange_fitness(shards, kind == O.ACTIVE)
        assert np.array_equal(fit, one.get(_cabi.F_FITNESS))
        for s in shards:
            s.rank(ik)
            assert np.array_equal(s.rank_idx(), one.rank_idx())
            s.shard_partial()
        exchange_partials(shards)
        for s in shards:
            s.shard_finish(sigma, fev)
            assert_close(s.get(_cabi.F_XMEAN), one.get(_cabi.F_XMEAN), what="xmean")
            assert_close(s.get(_cabi.F_C)[lower], one.get(_cabi.F_C)[lower], what="C")
            assert_close(s.get(_cabi.F_PS), one.get(_cabi.F_PS), what="ps")
            # keep the shards' replicated state bit-identical to each other, as an all-reduce does
        ref = shards[0]
        for s in shards[1:]:
            assert np.array_equal(s.get(_cabi.F_C), ref.get(_cabi.F_C)) and np.array_equal(s.get(_cabi.F_XMEAN), ref.get(_cabi.F_XMEAN))
        best = int(one.rank_idx()[0])
        owner = best // ll
        assert np.array_equal(shards[owner].candidate(best), X[best])
        # continue both from the unsharded state so rounding differences do not accumulate into the ranking
        for s in shards:
            for f in (_cabi.F_XMEAN, _cabi.F_PC, _cabi.F_PS, _cabi.F_C):
                s.set(f, one.get(f))


def test_shard_config_errors(ctx):
    w = np.ones(4) / 4
    consts = dict(cc=0.1, cs=0.1, c1=0.01, cmu=0.01, mueff=4.0, chi=3.0)
    with pytest.raises(_cabi.BbgpuError):
        _cabi.Cma(ctx, O.VANILLA, 8, 9, w, consts, shard=(0, 2))      # lambda not divisible
    with pytest.raises(_cabi.BbgpuError):
        _cabi.Cma(ctx, O.SEP, 8, 8, w, dict(consts, ccov=0.1), shard=(0, 2))
    g = _cabi.Cma(ctx, O.VANILLA, 8, 8, w, consts, shard=(1, 2))
    g.reset(np.zeros(8))
    with pytest.raises(_cabi.BbgpuError):
        g.update(1.0, 8)  # sharded update without a communicator must not silently run unsharded
