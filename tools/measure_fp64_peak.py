"""Measure the fp64 denominators MEASURED_PEAKS.json lacks: cuBLAS DGEMM
8192^3 burst + sustained, and our own DMMA GEMM at the headline shape.  Run under gpurun;
writes gpurun_out/fp64_peak.json (copied to profiles/ by hand)."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bboptpy_b200 import _cabi  # noqa: E402


def time_matmul(n, reps):
    a = torch.randn(n, n, dtype=torch.float64, device="cuda")
    b = torch.randn(n, n, dtype=torch.float64, device="cuda")
    for _ in range(2):
        torch.matmul(a, b)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    # sustained: back to back for ~3 s
    t0 = time.time()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    k = 0
    while time.time() - t0 < 3.0:
        for _ in range(4):
            torch.matmul(a, b)
        k += 4
        torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    sustained = e0.elapsed_time(e1) / k
    flop = 2.0 * n ** 3
    return flop / best / 1e9, flop / sustained / 1e9


def main():
    out = {"gpu": torch.cuda.get_device_name(0)}
    burst, sus = time_matmul(8192, 5)
    out["cublas_dgemm_8192_tflops_burst"] = burst
    out["cublas_dgemm_8192_tflops_sustained"] = sus
    burst4, sus4 = time_matmul(4096, 5)
    out["cublas_dgemm_4096_tflops_burst"] = burst4
    # cuBLAS at the headline shape (lambda x n x n)
    a = torch.randn(4096, 1000, dtype=torch.float64, device="cuda")
    b = torch.randn(1000, 1000, dtype=torch.float64, device="cuda")
    for _ in range(3):
        torch.matmul(a, b)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        torch.matmul(a, b)
    e1.record()
    torch.cuda.synchronize()
    out["cublas_dgemm_4096x1000x1000_tflops"] = 2.0 * 4096 * 1000 * 1000 / (e0.elapsed_time(e1) / 20) / 1e9
    ctx = _cabi.Context(0)
    for (M, N, K) in [(4096, 1000, 1000), (8192, 8192, 8192), (16384, 1000, 1000), (2048, 1000, 1000)]:
        reps = 3 if M == 8192 else 20
        ms = ctx.bench_dgemm(M, N, K, reps)
        out[f"bbgpu_dmma_{M}x{N}x{K}_ms"] = ms
        out[f"bbgpu_dmma_{M}x{N}x{K}_tflops"] = 2.0 * M * N * K / ms / 1e9
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/fp64_peak.json", "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
