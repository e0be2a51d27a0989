"""Cholesky-CMA-ES at the headline shape on the device: ms per generation and the factorisation's share.
    python tools/chol_time.py [steps]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from bboptpy_b200 import _cabi
ctx = _cabi.Context(0)
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
r = bench.bench_chol_device(ctx, steps, 3, seed=1)
print(f"generation {r['ms_per_generation']:.3f} ms ({r['gens_per_s']:.0f}/s), factorisation {r['factorisation_ms']:.3f} ms, fbest {r['fbest']:.9e}", flush=True)
