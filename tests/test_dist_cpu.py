"""world_size-2 CPU (`gloo`) tests of the multi-GPU host logic: the candidate-sharded update algebra against the
oracle, and the best-of reduce over independent runs.  No GPU: the device engine's NCCL path is covered by
tests/multi_gpu_shard.py on >= 2 GPUs, its per-shard kernels by tests/test_gpu_shard.py."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, kind, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    from bboptpy_b200 import sharding
    from oracle import oracle as O
    from util import assert_close, oracle_constants
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        P = O.port()
        P.seed(17)                       # same seed in every process: the replicated state
        n, lam = 12, 32
        box = 5.0
        h = P.cma(kind, 10 ** 9, 1e-12, lam)
        h.init(O.ROSENBROCK, -box * np.ones(n), box * np.ones(n), P.uniform(-box, box, n))
        consts = oracle_constants(h, kind)
        consts["weights"] = h.getv("weights")
        lo, hi = sharding.shard_bounds(lam, world, rank)
        for gen in range(5):
            s0 = h.snapshot()
            h.sample()
            h.evalsort()
            X = h.mat("arx")
            rank_idx = h.getv("fit_idx").astype(np.int64)
            norms_loc = None
            if kind == O.ACTIVE:
                d = (X[lo:hi] - s0["xmean"]) @ s0["invsqrtC"].T
                norms_loc = (d * d).sum(1)
            eig_before = h.geti("eigenlastev")
            h.update()
            state = {k: s0[k] for k in ("xmean", "pc", "ps", "invsqrtC")}
            state["C"] = np.tril(s0["C"]) + np.tril(s0["C"], -1).T
            xmean, ps, pc, Cl, yc = sharding.sharded_update(dist, X[lo:hi], lo, hi, rank_idx, state, consts, kind,
                                                            s0["sigma"], h.geti("fev"), lam, norms_loc)
            assert_close(xmean, h.getv("xmean"), what=f"gen {gen} xmean")
            assert_close(ps, h.getv("ps"), what=f"gen {gen} ps")
            assert_close(pc, h.getv("pc"), what=f"gen {gen} pc")
            if h.geti("eigenlastev") == eig_before:
                assert_close(Cl[np.tril_indices(n)], h.mat("C")[np.tril_indices(n)], what=f"gen {gen} C")
            if kind == O.ACTIVE:
                assert_close(yc, h.getv("ycoeff"), rel=1e-9, what=f"gen {gen} ycoeff")
            h.history()
        # best-of reduce: rank 1 holds the better run
        fb, owner, xb = sharding.best_of(dist, 3.0 - 2.0 * rank, np.full(4, float(rank)))
        assert (fb, owner) == (1.0, 1) and np.array_equal(xb, np.ones(4))
        q.put((rank, "ok"))
    except Exception as e:  # surfaced by the parent
        import traceback
        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("kind", [0, 1])
def test_sharded_update_algebra_gloo_world2(kind):
    import socket
    import torch.multiprocessing as mp
    pytest.importorskip("torch")
    from oracle import oracle as O
    O.build(("port",))
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, kind, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, msg in results:
        assert msg == "ok", f"rank {rank}: {msg}"


def test_shard_bounds_and_coefficients():
    sys.path.insert(0, ROOT)
    from bboptpy_b200 import sharding
    assert sharding.shard_bounds(16, 4, 2) == (8, 12)
    with pytest.raises(ValueError):
        sharding.shard_bounds(10, 4, 0)
    lam, mu = 8, 4
    w = np.array([0.4, 0.3, 0.2, 0.1])
    rank_idx = np.array([5, 2, 7, 0, 1, 3, 4, 6])
    tot_c, tot_w = np.zeros(lam), np.zeros(lam)
    for r in range(2):
        lo, hi = sharding.shard_bounds(lam, 2, r)
        c, wm, _ = sharding.shard_coefficients(rank_idx, lam, mu, w, lo, hi, cpos=2.0)
        tot_c[lo:hi], tot_w[lo:hi] = c, wm
    assert np.array_equal(tot_w[rank_idx[:mu]], w) and np.all(tot_w[rank_idx[mu:]] == 0)
    assert np.array_equal(tot_c[rank_idx[:mu]], 2.0 * w)
    assert sharding.best_of(None, 2.5, np.arange(3.0))[:2] == (2.5, 0)
