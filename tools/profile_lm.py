"""LM-CMA-ES at the bench shape (n = 10^4, lambda = 64, memory 200) through the Python API: generations/s, for ncu launch lists.
    python tools/profile_lm.py [generations]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, bboptpy_b200 as pkg
gens = int(sys.argv[1]) if len(sys.argv) > 1 else 400
bench.bench_lm_e2e(pkg, 20, 1)
for _ in range(3):
    print("LM-CMA-ES generations/s:", bench.bench_lm_e2e(pkg, gens, 1), flush=True)
