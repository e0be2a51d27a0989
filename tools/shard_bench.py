"""Sharded large-lambda run only (BASELINE config 4), peer-memory exchange vs ncclAllReduce:
    python -m torch.distributed.run --nproc-per-node G tools/shard_bench.py [lambda] [steps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

lam = int(sys.argv[1]) if len(sys.argv) > 1 else bench.SHARD_LAMBDA
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
rank, world, local, dist = bench.dist_setup(0)
from bboptpy_b200 import _cabi  # noqa: E402
ctx = _cabi.Context(local)
ids = [_cabi.Comm.unique_id() if rank == 0 else None]
dist.broadcast_object_list(ids, src=0)
comm = _cabi.Comm(ctx, world, rank, ids[0])
for mode in ("peer", "nccl", "peer", "nccl"):
    if mode == "nccl":
        os.environ["BBGPU_SHARD_P2P"] = "0"
    else:
        os.environ.pop("BBGPU_SHARD_P2P", None)
    for kname, kind in (("cmaes", _cabi.CMA_FULL), ("active", _cabi.CMA_ACTIVE)):
        bench.barrier(dist, local)
        r = bench.bench_sharded(ctx, comm, kind, lam, steps, 3, seed=77)
        t = bench.all_max(dist, local, sum(r["times_ms"]) / 1e3)
        if rank == 0:
            print(f"{mode:5s} {kname:7s} lambda {lam} over {world} GPUs: {len(r['times_ms']) / t:8.1f} generations/s, comm share "
                  f"{r['shares'].get('comm', 0):.3f}, syrk {r['shares'].get('cov_syrk', 0):.3f}", flush=True)
dist.barrier()
dist.destroy_process_group()
