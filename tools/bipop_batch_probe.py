"""Timing probe of the batched BIPOP driver (BASELINE config 5): python tools/bipop_batch_probe.py [runs] [n] [mfev] [quantum]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bboptpy_b200 as pkg  # noqa: E402

runs = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
mfev = int(sys.argv[3]) if len(sys.argv) > 3 else 100000
quantum = int(sys.argv[4]) if len(sys.argv) > 4 else 3000
obj = {"rastrigin": pkg.Rastrigin, "ackley": pkg.Ackley}[sys.argv[5] if len(sys.argv) > 5 else "rastrigin"]
box = 5.12 if obj is pkg.Rastrigin else 32.768
pkg.seed(1)
guesses = np.random.default_rng(0).uniform(-box, box, (runs, n))
for rep in range(int(os.environ.get('REPS', '2'))):
    b = pkg.BiPopCMAESBatch(runs=runs, mfev=mfev, tol=1e-8)
    b.set_quantum(quantum)
    b.set_timing(True)
    b.set_streams(int(os.environ.get('STREAMS', '0')))
    t0 = time.perf_counter()
    sols = b.optimize(obj, -box * np.ones(n), box * np.ones(n), guesses)
    dt = time.perf_counter() - t0
    st = b.stats
    f = b.best_values
    print(f"rep {rep}: runs {runs} n {n} mfev {mfev} quantum {quantum}: wall {dt:.3f} s, device {st.device_ms / 1e3:.3f} s, launches {st.launches}, "
          f"base runs {st.base_runs}, generations {st.generations} ({st.generations / dt / 1e6:.2f} M run-gens/s), evals {st.evaluations} "
          f"({st.evaluations / dt / 1e6:.1f} M evals/s), decompositions {st.decompositions}, runs/s {runs / dt:.1f}; "
          f"f median {np.median(f):.3f} min {f.min():.3f}", flush=True)
