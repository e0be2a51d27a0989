"""time the sampling-GEMM shape for the tile variant selected by BBGPU_GEMM_VARIANT (one process per variant)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bboptpy_b200 import _cabi
ctx = _cabi.Context(0)
for (M, N, K) in [(4096, 1000, 1000), (2048, 1000, 1000)]:
    ms = ctx.bench_dgemm(M, N, K, 30)
    print("variant", os.environ.get("BBGPU_GEMM_VARIANT", "auto"), (M, N, K), "ms %.4f TF %.2f" % (ms, 2.0 * M * N * K / ms / 1e9))
