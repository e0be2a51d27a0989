"""NCCL transport of the candidate-sharded CMA-ES run: launched with one process per GPU
(`python -m torch.distributed.run --nproc-per-node G tests/multi_gpu_shard.py`), needs G >= 2 GPUs.

Every rank also steps the UNSHARDED handle with the same Philox seed: draws, fitness and ranking must be
bit-identical, mean / paths / covariance within 1e-10, and the replicated state bit-identical across ranks."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from bboptpy_b200 import _cabi, strategy  # noqa: E402
from util import assert_close, update_sigma_host  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = _cabi.Context(local)
    ids = [_cabi.Comm.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    comm = _cabi.Comm(ctx, world, rank, ids[0])
    for kind in (_cabi.CMA_FULL, _cabi.CMA_ACTIVE):
        n, lam = 96, 64 * world
        w, consts = strategy.cma_constants(kind, n, lam)
        st = consts
        obj = _cabi.Objective(ctx, n, builtin=_cabi.OBJ_ROSENBROCK)
        guess = np.linspace(-3, 3, n)
        one = _cabi.Cma(ctx, kind, n, lam, w, consts, seed=2024)
        sh = _cabi.Cma(ctx, kind, n, lam, w, consts, seed=2024, shard=(rank, world))
        sh.attach_comm(comm)
        one.reset(guess)
        sh.reset(guess)
        sigma1 = sigma2 = 2.0
        fev, ik, ll = 0, st["ik"], lam // world
        lower = np.tril_indices(n)
        for gen in range(8):
            fev += lam
            one.generation(obj, sigma1, gen, ik, fev)
            sh.generation(obj, sigma2, gen, ik, fev)
            r1, r2 = one.fetch(), sh.fetch()
            assert np.array_equal(sh.get(_cabi.F_ARX), one.get(_cabi.F_ARX)[rank * ll:(rank + 1) * ll]), "draws / sample rows"
            assert np.array_equal(sh.get(_cabi.F_FITNESS), one.get(_cabi.F_FITNESS)), "all-gathered fitness"
            assert np.array_equal(sh.rank_idx(), one.rank_idx()), "ranking"
            assert list(r1.ibw) == list(r2.ibw) and r1.hsig == r2.hsig
            assert_close(sh.get(_cabi.F_XMEAN), one.get(_cabi.F_XMEAN), what="xmean")
            assert_close(sh.get(_cabi.F_PS), one.get(_cabi.F_PS), what="ps")
            assert_close(sh.get(_cabi.F_PC), one.get(_cabi.F_PC), what="pc")
            assert_close(sh.get(_cabi.F_C)[lower], one.get(_cabi.F_C)[lower], what="C")
            # replicated state must be bit-identical on every rank
            mine = torch.from_numpy(np.concatenate([sh.get(_cabi.F_C).ravel(), sh.get(_cabi.F_XMEAN)])).cuda()
            ref = mine.clone()
            dist.broadcast(ref, src=0)
            assert torch.equal(mine, ref), "replicated state diverged between ranks"
            best = sh.candidate(int(r2.ibw[0]))
            assert np.array_equal(best, one.get(_cabi.F_ARX)[int(r1.ibw[0])])
            sigma1 = update_sigma_host(sigma1, r1.pslen, consts["cs"], st["damps"], consts["chi"], r1.ybw[0], r1.f_kth, gen, 10 ** 9, 0.0, 1.0)
            sigma2 = update_sigma_host(sigma2, r2.pslen, consts["cs"], st["damps"], consts["chi"], r2.ybw[0], r2.f_kth, gen, 10 ** 9, 0.0, 1.0)
            one.eig()
            sh.eig()
            # continue the sharded run from the unsharded state: rounding differences must not reach the ranking
            for f in (_cabi.F_XMEAN, _cabi.F_PC, _cabi.F_PS, _cabi.F_C):
                sh.set(f, one.get(f))
            sh.eig()
            sigma2 = sigma1
        if rank == 0:
            print("sharded", "active" if kind == _cabi.CMA_ACTIVE else "vanilla", "over", world, "GPUs: ok; rank-mu exchange over",
                  "NVLink peer memory (one kernel)" if sh.uses_peer_memory() else "ncclAllReduce", flush=True)
        # the peer-memory all-reduce against NCCL's on the same partials: same sums up to the order of addition, and the
        # peer-memory result bit-identical on every rank
        peer = sh.uses_peer_memory()
        del sh
        os.environ["BBGPU_SHARD_P2P"] = "0"
        nc = _cabi.Cma(ctx, kind, n, lam, w, consts, seed=2024, shard=(rank, world))
        nc.attach_comm(comm)
        del os.environ["BBGPU_SHARD_P2P"]
        pm = _cabi.Cma(ctx, kind, n, lam, w, consts, seed=2024, shard=(rank, world))
        pm.attach_comm(comm)
        assert not nc.uses_peer_memory() and pm.uses_peer_memory() == peer
        for hnd in (nc, pm):
            hnd.reset(guess)
            hnd.generation(obj, 2.0, 0, ik, lam)
            hnd.fetch()
        assert_close(pm.get(_cabi.F_C)[lower], nc.get(_cabi.F_C)[lower], what="C: peer-memory vs NCCL exchange")
        assert_close(pm.get(_cabi.F_XMEAN), nc.get(_cabi.F_XMEAN), what="xmean: peer-memory vs NCCL exchange")
        mine = torch.from_numpy(np.concatenate([pm.get(_cabi.F_C).ravel(), pm.get(_cabi.F_XMEAN)])).cuda()
        ref = mine.clone()
        dist.broadcast(ref, src=0)
        assert torch.equal(mine, ref), "peer-memory exchange: ranks disagree"
        del nc, pm
    # ---- the same through the reference-facing Python API: CMAES.set_communicator + optimize
    os.environ["BBOPT_DEVICE"] = str(local)
    import bboptpy_b200 as pkg
    n, lam, gens = 64, 32 * world, 12
    lo, hi, guess = -10.0 * np.ones(n), 10.0 * np.ones(n), np.linspace(-2, 2, n)
    ids2 = [pkg.Communicator.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids2, src=0)
    pcomm = pkg.Communicator(world, rank, ids2[0])
    rosen = lambda x: float(np.sum(100 * (x[1:] - x[:-1] ** 2) ** 2 + (1 - x[:-1]) ** 2))
    for cls in (pkg.CMAES, pkg.ActiveCMAES):
        # (1) one generation: same Philox key, B = I -> the sharded run must reproduce the unsharded one.  Later generations
        # are NOT comparable run-to-run: the eigenbasis of a nearly isotropic C is not a function of C, so a
        # 1e-16 difference in C rotates B and with it every sample.
        single = cls(mfev=lam, tol=0.0, np=lam)
        single.set_stream_key(4242)
        s1 = single.optimize(pkg.Rosenbrock, lo, hi, guess)
        shard = cls(mfev=lam, tol=0.0, np=lam)
        shard.set_communicator(pcomm, 4242)
        s2 = shard.optimize(pkg.Rosenbrock, lo, hi, guess)
        assert s1.n_evals == s2.n_evals == lam
        assert np.array_equal(s2.x, s1.x), "best candidate of generation 0"
        assert_close(shard.state("xmean"), single.state("xmean"), what="xmean after one sharded generation")
        tl = np.tril_indices(n)
        assert_close(shard.state("C").reshape(n, n)[tl], single.state("C").reshape(n, n)[tl], what="C after one sharded generation")
        # (2) a longer sharded run: every rank must hold bit-identical replicated state and return the same solution
        shard = cls(mfev=gens * lam, tol=0.0, np=lam)
        shard.set_communicator(pcomm, 777)
        s3 = shard.optimize(pkg.Rosenbrock, lo, hi, guess)
        assert s3.n_evals == gens * lam and np.isfinite(rosen(s3.x))
        mine = torch.from_numpy(np.concatenate([s3.x, shard.state("C"), shard.state("xmean"), shard.state("B")])).cuda()
        ref = mine.clone()
        dist.broadcast(ref, src=0)
        assert torch.equal(mine, ref), "ranks disagree after the sharded optimize()"
        if rank == 0:
            print(cls.__name__, "optimize() sharded over", world, "GPUs: ok", rosen(guess), "->", rosen(s3.x), flush=True)
    # ---- a batch of independent BIPOP runs sharded r % world == rank: every run must come out
    # bit-identical to the same run of the unsharded batch (run-keyed Philox streams and restart draws), and the best-of
    # reduce must name the same winner on every rank
    n, R = 8, 24
    lo, hi = -5.12 * np.ones(n), 5.12 * np.ones(n)
    guesses = np.random.default_rng(4).uniform(-5, 5, (R, n))
    pkg.seed(99)
    whole = pkg.BiPopCMAESBatch(runs=R, mfev=12000, tol=1e-8)
    sols_w = whole.optimize(pkg.Rastrigin, lo, hi, guesses)
    fw, ow, xw = whole.best_of()
    pkg.seed(99)
    part = pkg.BiPopCMAESBatch(runs=R, mfev=12000, tol=1e-8)
    part.set_communicator(pcomm)
    sols_p = part.optimize(pkg.Rastrigin, lo, hi, guesses)
    assert list(part.run_ids) == list(range(rank, R, world))
    for s, r in zip(sols_p, part.run_ids):
        assert np.array_equal(s.x, sols_w[r].x) and s.n_evals == sols_w[r].n_evals, ("run", r)
    fp, op, xp = part.best_of()
    assert (fp, op) == (fw, ow) and np.array_equal(xp, xw), (fp, op, fw, ow)
    if rank == 0:
        print("BIPOP batch of", R, "runs sharded over", world, "GPUs: ok, best run", op, "f", fp, flush=True)
    del part, whole
    del pcomm
    # best-of reduce
    f_dev = torch.tensor([10.0 - rank if rank != 1 else -5.0], dtype=torch.float64, device="cuda")
    x_dev = torch.full((7,), float(rank), dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    fb, owner = comm.best_of(f_dev.data_ptr(), x_dev.data_ptr(), 7)
    exp_owner = 1 if world > 1 else 0
    assert owner == exp_owner and fb == (-5.0 if world > 1 else 10.0), (fb, owner)
    ctx.sync()
    assert torch.equal(x_dev.cpu(), torch.full((7,), float(exp_owner), dtype=torch.float64))
    if rank == 0:
        print("best-of reduce: ok", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
