"""Collects tests/multi_gpu_shard.py under pytest: the NCCL / NVLink-peer-memory transport of the candidate-sharded
CMA-ES run, the sharded BIPOP batch and the best-of reduce, one process per GPU.  Skipped on boxes with fewer than two
GPUs (the single-GPU round-end run); `gpurun --gpus N -- python -m pytest tests -m gpu` runs it with every GPU of the box."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpu_count():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.skipif(_gpu_count() < 2, reason="needs >= 2 GPUs on one node")
def test_nccl_sharded_run_over_all_gpus():
    world = min(_gpu_count(), 8)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29731", os.path.join(ROOT, "tests", "multi_gpu_shard.py")]
    r = subprocess.run(cmd, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-4000:]
    assert "best-of reduce: ok" in r.stdout
