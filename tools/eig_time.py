"""Headline generation on the device: ms per generation and the share of the eigendecomposition (BBGPU_SYTRD=..., BBGPU_DC_GEMM=...,
BBGPU_DC_LEAF=... select the comparison kernels).   python tools/eig_time.py [steps]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from bboptpy_b200 import _cabi
ctx = _cabi.Context(0)
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
res = bench.bench_cma_device(ctx, steps, 3, seed=1)
print(f"generation {np.mean(res['times_ms']):.3f} ms, eig {res['eig_ms']:.3f} ms, rest {res['excl_eig_ms']:.3f} ms, fbest {res['fbest']:.6e}", flush=True)
