"""SaNSDE / SSDE at the bench shape (n = np = 1000, Rastrigin) through the Python API: evaluations/s, for ncu launch lists.
    python tools/profile_de2.py sansde|ssde|ssde2 [generations]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, bboptpy_b200 as pkg
which = sys.argv[1] if len(sys.argv) > 1 else "ssde"
gens = int(sys.argv[2]) if len(sys.argv) > 2 else 200
fn = {"sansde": lambda g: bench.bench_sansde_e2e(pkg, g, 1), "ssde": lambda g: bench.bench_ssde_e2e(pkg, g, 1),
      "ssde2": lambda g: bench.bench_ssde_e2e(pkg, g, 1, usede=True)}[which]
fn(5)
for _ in range(2):
    print(which, "evaluations/s, rounds:", fn(gens), flush=True)
