#!/usr/bin/env python
"""bench.py -- the reference's headline metric on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl own|reference]

Workload (BASELINE.json `metric`): full-covariance CMA-ES (reference class `Cmaes`), n = 1000,
lambda = 4096, Rosenbrock on [-10,10]^n, synthetic data.  A step = one whole generation of the
reference's loop at this configuration: sample -> evaluate -> rank -> update -> sigma -> the
eigendecomposition (the reference's gate fires after EVERY generation here: eigenfreq = 465
evaluations < lambda, cma/cmaes.cpp:48,233).  `value` and `e2e` therefore both INCLUDE the
decomposition, in both arms (own and --impl reference); the decomposition is also timed
separately (`eig.ms`) and `gens_per_s_excl_eig` is reported beside it in both arms, as the north
star asks.  The same JSON line also carries the DE half of the metric (JADE / L-SHADE
candidate-evals/s, n = 1000, np = 1000, Rastrigin) under "de".

N > 1 (torchrun, one rank per GPU): independent runs sharded per GPU (weak scaling), max over
ranks of the device time, final best-of reduce over NCCL.

--impl reference: the reference's own CPU implementation (oracle/_ref, the unmodified sources
compiled in place; else the pinned port) on the host cores, same metric/config: every step is ONE
WHOLE GENERATION at n = 1000, lambda = 4096 through the reference's own phases (no scaling, no
subtraction); the run is bounded by wall clock (the steps actually timed are reported).
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_DIM, LAMBDA, BOX = 1000, 4096, 10.0
DE_N, DE_NP, DE_BOX = 1000, 1000, 5.12
DE_BATCH = 148         # independent DE runs per launch for the HBM-resident figure (one per SM: no grid-wide barrier)
SEP_N = 10000          # BASELINE config 3
SHARD_LAMBDA = 16384   # BASELINE config 4
BIPOP_RUNS, BIPOP_N, BIPOP_MFEV = 4096, 40, 100000  # BASELINE config 5
METRIC = "CMA-ES generations/s (n=1000, lambda=4096)"
UNIT = "generations/s"
WORKLOAD = "Cmaes n=1000 lambda=4096 mu=2048 Rosenbrock box=[-10,10]^n fp64 (BASELINE.json metric config)"
STEP_DEF = ("one whole generation: sample -> evaluate -> rank -> update -> sigma -> eigendecomposition (the reference's gate "
            "fires every generation at this configuration); gens_per_s_excl_eig reported beside it")


def fp64_peak():
    """fp64 denominator: MEASURED_PEAKS.json has no fp64 entry and the profiling guide states no fp64
    fallback, so the denominator is our own cuBLAS DGEMM 8192^3 measurement on this pool's B200
    (tools/measure_fp64_peak.py -> profiles/r01_fp64_peak.json)."""
    try:
        mp = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        if "fp64_tflops" in mp:
            return float(mp["fp64_tflops"]), "MEASURED_PEAKS.json fp64_tflops (of measured)"
    except Exception:
        pass
    try:
        pk = json.load(open(os.path.join(ROOT, "profiles", "r01_fp64_peak.json")))
        return float(pk["cublas_dgemm_8192_tflops_burst"]), ("profiles/r01_fp64_peak.json: cuBLAS DGEMM 8192^3 burst on "
                                                             "this pool's B200 (of measured; MEASURED_PEAKS.json has no fp64 entry)")
    except Exception:
        return 37.0, "nominal B200 fp64 (no measurement available)"


def hbm_peak():
    try:
        mp = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(mp["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
    except Exception:
        return 6650.0, "B200_PROFILING.md fallback 6.65 TB/s (of fallback)"


def ncu_kernel_block(kernel_prefix, name_glob="r0*_ncu_gemm*.txt"):
    """Metrics of one kernel from the newest committed ncu summary under profiles/ (written by tools/ncu_summary.py from a
    `ncu --set full` capture): {metric: (value, unit)}, plus the file it came from.  bench.py never runs under ncu; the
    roofline's `traffic` is read from the evidence file instead of being a literal."""
    import glob
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", name_glob)), reverse=True):
        block, out = False, {}
        for line in open(path):
            if line.startswith("## "):
                if block:
                    break
                block = kernel_prefix in line
                continue
            parts = line.split()
            if block and len(parts) >= 2:
                try:
                    out[parts[0]] = (float(parts[1]), parts[2] if len(parts) > 2 else "")
                except ValueError:
                    pass
        if out:
            return out, os.path.relpath(path, ROOT)
    return {}, None


def to_bytes(value_unit):
    if value_unit is None:
        return None
    v, u = value_unit
    return v * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for name, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def dist_setup(n_gpus):
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist_mod
        torch.cuda.set_device(local)
        dist_mod.init_process_group("nccl", device_id=torch.device("cuda", local))
        dist = dist_mod
    return rank, world, local, dist


def all_max(dist, local, value):
    if dist is None:
        return value
    import torch
    t = torch.tensor([value], dtype=torch.float64, device=f"cuda:{local}")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def all_min_with_owner(dist, local, rank, value):
    """final best-of reduce over NCCL: (min fitness, owning rank)"""
    if dist is None:
        return value, rank
    import torch
    t = torch.tensor([value], dtype=torch.float64, device=f"cuda:{local}")
    gathered = [torch.zeros_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(gathered, t)
    vals = [float(g.item()) for g in gathered]
    owner = int(np.argmin(vals))
    return vals[owner], owner


def all_sum(dist, local, value):
    if dist is None:
        return value
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64, device=f"cuda:{local}")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def barrier(dist, local):
    if dist is not None:
        import torch
        dist.barrier()
        torch.cuda.synchronize(local)


# ------------------------------------------------------------------------------------------------
# CPU side: the reference (or the pinned port) on the host cores, bounded sample
# ------------------------------------------------------------------------------------------------
def cpu_oracle():
    from oracle import oracle as O
    if O.have_ref():
        return O.ref(), "reference"
    O.build(("port",))
    return O.port(), "port"


class CpuRun:
    """The CPU implementation at the REAL configuration (n = 1000, lambda = 4096): every generation() is one whole
    generation through the reference's own phases -- samplePopulation, evaluateAndSortPopulation, updateDistribution
    (which ends with updateSigma and, at this configuration, the eigendecomposition every time), updateHistory -- timed
    by phase.  Nothing is scaled or subtracted."""

    def __init__(self, o, seed=1):
        from oracle import oracle as O
        o.seed(seed)
        n = N_DIM
        self.h = o.cma(O.VANILLA, 10 ** 9, 1e-12, LAMBDA)
        guess = o.uniform(-BOX, BOX, n)
        self.h.init(O.ROSENBROCK, -BOX * np.ones(n), BOX * np.ones(n), guess)

    def generation(self):
        h = self.h
        eig_before = h.geti("eigenlastev")
        t0 = time.perf_counter()
        h.sample()
        t1 = time.perf_counter()
        h.evalsort()
        t2 = time.perf_counter()
        h.update()
        t3 = time.perf_counter()
        h.history()
        t4 = time.perf_counter()
        return {"t_sample": t1 - t0, "t_evalsort": t2 - t1, "t_update_incl_eig": t3 - t2, "t_generation": t4 - t0,
                "eig_fired": h.geti("eigenlastev") != eig_before}

    def time_eig(self):
        """one more decomposition of the current C, forced past the gate, timed alone"""
        return self.h.time_eig()


def cpu_sample(o, gens=1, seed=1, budget_s=1e9):
    """`gens` whole generations of the CPU implementation (bounded by budget_s of wall clock, at least one) and one
    separately timed decomposition; gens_per_s_excl_eig = 1 / (t_generation - t_eig), both measured."""
    run = CpuRun(o, seed)
    t_start = time.perf_counter()
    rows = []
    for k in range(gens):
        rows.append(run.generation())
        if time.perf_counter() - t_start + rows[-1]["t_generation"] > budget_s:
            break
    t_eig = run.time_eig()
    t_gen = float(np.mean([r["t_generation"] for r in rows]))
    return {"generations": len(rows), "t_generation": t_gen, "t_sample": float(np.mean([r["t_sample"] for r in rows])),
            "t_evalsort": float(np.mean([r["t_evalsort"] for r in rows])),
            "t_update_incl_eig": float(np.mean([r["t_update_incl_eig"] for r in rows])), "t_eig": t_eig,
            "eig_fired_every_generation": all(r["eig_fired"] for r in rows),
            "gens_per_s": 1.0 / t_gen, "gens_per_s_excl_eig": 1.0 / max(t_gen - t_eig, 1e-9)}


def cpu_de_sample(o, kind, gens=3, seed=1):
    from oracle import oracle as O
    o.seed(seed)
    h = o.de(kind, 10 ** 7, DE_NP, 1e-12)
    h.init(O.RASTRIGIN, -DE_BOX * np.ones(DE_N), DE_BOX * np.ones(DE_N))
    t0 = time.perf_counter()
    for _ in range(gens):
        h.iterate()
    return DE_NP * gens / (time.perf_counter() - t0)


def _cpu_worker(seed):
    o, _ = cpu_oracle()
    return cpu_sample(o, gens=1, seed=seed)["gens_per_s"]


def cpu_all_cores(max_procs=32):
    """The reference is single-threaded (one process-global RNG, src/random.hpp:168): more cores only help as independent
    runs, one process each -- the CPU counterpart of our N-GPU independent-runs scaling."""
    import multiprocessing as mp
    procs = max(1, min(os.cpu_count() or 1, max_procs))
    with mp.get_context("spawn").Pool(procs) as pool:
        t0 = time.perf_counter()
        rates = pool.map(_cpu_worker, range(100, 100 + procs))
        wall = time.perf_counter() - t0
    return {"processes": procs, "aggregate_gens_per_s": float(sum(rates)), "per_process_gens_per_s": float(np.mean(rates)),
            "wall_s": wall, "note": "P independent runs, one single-threaded process per core, one whole generation each at the real configuration"}


def run_reference(args):
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    if rank != 0:
        return
    o, kind = cpu_oracle()
    budget = 150.0  # seconds of timed generations: the steps actually timed are what the line reports
    run = CpuRun(o, seed=1)
    warm = min(args.warmup, 1)
    for _ in range(warm):
        run.generation()
    t_start = time.perf_counter()
    rows = []
    for k in range(max(1, args.steps)):
        rows.append(run.generation())
        if time.perf_counter() - t_start + rows[-1]["t_generation"] > budget:
            break
    timed_s = time.perf_counter() - t_start
    t_eig = run.time_eig()
    t_gen = timed_s / len(rows)
    value = 1.0 / t_gen
    de = {"jade_evals_per_s": cpu_de_sample(o, 0), "shade_evals_per_s": cpu_de_sample(o, 1), "sansde_evals_per_s": cpu_sansde_sample(o),
          "ssde_evals_per_s": cpu_ssde_sample(o),
          "unit": "candidate-evals/s",
          "sample": "3 generations each, n=1000 np=1000 Rastrigin, 1 thread"}
    sample = (f"{len(rows)} whole generations at n=1000, lambda=4096 through the reference's own phases (sample, evaluate+sort, "
              "update incl. sigma and the eigendecomposition, which fires every generation here), after "
              f"{warm} warm-up generation(s); nothing scaled; the reference is single-threaded (process-global RNG)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": len(rows),
        "warmup": warm, "ms_per_step": t_gen * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "step": STEP_DEF},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": kind, "sample": sample,
                         "all_cores": cpu_all_cores()},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "eig": {"ms": t_eig * 1e3, "cadence_generations": 1,
                "fired_every_generation": all(r["eig_fired"] for r in rows),
                "gens_per_s_excl_eig": 1.0 / max(t_gen - t_eig, 1e-9),
                "note": "ms: one decomposition forced past the gate and timed alone; excl_eig = 1 / (t_generation - ms)"},
        "phases_ms": {k: float(np.mean([r[k] for r in rows])) * 1e3 for k in ("t_sample", "t_evalsort", "t_update_incl_eig")},
        "timed_region_s": timed_s,
        "de": de,
        "cholesky": {"gens_per_s": cpu_chol_sample(o), "sample": "1 whole generation of CholeskyCmaes, n=1000 lambda=4096, 1 thread"},
        "lm_cma": {"gens_per_s": cpu_lm_sample(o), "sample": f"400 generations of LmCmaes (the memory fills during the run), n={LM_N} lambda={LM_LAMBDA}, 1 thread"},
        "xnes": {"gens_per_s": cpu_xnes_sample(o), "sample": "1 generation of xNES, n=1000, 1 thread"},
        "bipop_batch": cpu_bipop_sample(o),
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# GPU side
# ------------------------------------------------------------------------------------------------
def bench_cma_device(ctx, steps, warmup, seed):
    """`value`: whole generations/s with everything resident in HBM: sample -> evaluate -> rank -> update -> sigma (host,
    from the 104-byte report) -> eigendecomposition, the reference's cadence at this configuration (every generation).
    The decomposition's share comes from the CUDA-event profile of the same timed region."""
    from bboptpy_b200 import _cabi
    from bboptpy_b200.strategy import cma_constants, flops_per_generation, update_sigma
    n, lam = N_DIM, LAMBDA
    w, c = cma_constants(0, n, lam)
    assert c["eigenfreq"] < lam, "the reference's gate fires every generation at this configuration"
    rng = np.random.default_rng(seed)
    guess = rng.uniform(-BOX, BOX, n)
    g = _cabi.Cma(ctx, _cabi.CMA_FULL, n, lam, w, c, seed=seed)
    g.reset(guess)
    obj = _cabi.Objective(ctx, n, builtin=_cabi.OBJ_ROSENBROCK)
    state = {"sigma": 2.0, "fev": 0, "gen": 0}

    def step():
        state["fev"] += lam
        g.generation(obj, state["sigma"], state["gen"], c["ik"], state["fev"])
        rep = g.fetch()
        state["sigma"] = update_sigma(state["sigma"], rep.pslen, c, rep.ybw[0], rep.f_kth)
        state["gen"] += 1
        g.eig()  # cma/cmaes.cpp:233: fev - eigenlastev = lambda > eigenfreq, every generation
        return rep

    for i in range(warmup):
        step()
    g.fetch()
    ctx.profile_reset()
    ctx.profile(True)
    launches0 = ctx.launches
    times = []
    rep = None
    for _ in range(steps):
        ctx.flush_l2()
        ctx.timer_start()
        rep = step()
        times.append(ctx.timer_stop())
    launches = ctx.launches - launches0
    prof = ctx.profile_read()
    ctx.profile(False)
    rep = g.fetch()
    assert rep.eig_info == 0
    eig_total = prof.get("eig", {}).get("ms_total", 0.0)
    return {"times_ms": times, "launches": launches, "prof": prof, "eig_ms": eig_total / max(steps, 1),
            "excl_eig_ms": (sum(times) - eig_total) / max(steps, 1), "fbest": rep.ybw[0],
            "flops": flops_per_generation(0, n, lam)}


def bench_cma_e2e(pkg, gens, seed):
    """Through the reference-facing API with host buffers: optimize(f, lower, upper, guess) -> solution.
    Includes init uploads, the eigendecomposition at the reference's cadence (every generation here), the per-generation
    report reads and the final solution download."""
    n, lam = N_DIM, LAMBDA
    rng = np.random.default_rng(seed)
    lower, upper, guess = -BOX * np.ones(n), BOX * np.ones(n), rng.uniform(-BOX, BOX, n)
    alg = pkg.CMAES(mfev=gens * lam, tol=0.0, np=lam)
    t0 = time.perf_counter()
    sol = alg.optimize(pkg.Rosenbrock, lower, upper, guess)
    dt = time.perf_counter() - t0
    done = sol.n_evals // lam
    h2d = (3 * n + lam // 2) * 8 / max(done, 1)           # lower, upper, guess, weights per optimize()
    d2h = 2 * 104 + n * 8 / max(done, 1)                  # report after update + after probe; final x once
    return done / dt, dt, done, h2d, d2h, alg.n_decompositions


def bench_chol_device(ctx, steps, warmup, seed):
    """Cholesky-CMA-ES (cma/cholesky_cmaes.cpp), the reference's eigendecomposition-free variant, at the headline shape
    n = 1000, lambda = 4096: whole generations/s resident in HBM (sample -> evaluate -> rank -> SYRK over [A^T; Y] -> blocked
    Cholesky factorisation -> sigma on the host)."""
    from bboptpy_b200 import _cabi
    from bboptpy_b200.strategy import cma_constants, update_sigma
    n, lam = N_DIM, LAMBDA
    w, c = cma_constants(3, n, lam)
    g = _cabi.Cma(ctx, _cabi.CMA_CHOLESKY, n, lam, w, c, seed=seed)
    g.reset(np.random.default_rng(seed).uniform(-BOX, BOX, n))
    obj = _cabi.Objective(ctx, n, builtin=_cabi.OBJ_ROSENBROCK)
    state = {"sigma": 2.0, "fev": 0, "gen": 0}

    def step():
        state["fev"] += lam
        g.generation(obj, state["sigma"], state["gen"], c["ik"], state["fev"])
        rep = g.fetch()
        state["sigma"] = update_sigma(state["sigma"], rep.pslen, c, rep.ybw[0], rep.f_kth)
        state["gen"] += 1
        return rep

    for _ in range(warmup):
        step()
    ctx.profile_reset()
    ctx.profile(True)
    times = []
    for _ in range(steps):
        ctx.flush_l2()
        ctx.timer_start()
        rep = step()
        times.append(ctx.timer_stop())
    prof = ctx.profile_read()
    ctx.profile(False)
    assert rep.eig_info == 0
    mu = lam // 2
    flops = 2.0 * n * n * lam + (n + mu) * n * (n + 1.0) + n ** 3 / 3.0  # sampling GEMM, SYRK over n + mu rows, factorisation
    ms = float(np.mean(times))
    return {"ms_per_generation": ms, "gens_per_s": 1e3 / ms, "fbest": rep.ybw[0], "algorithmic_flops_per_generation": flops,
            "kernel_share_of_step": {k: v["ms_total"] / sum(times) for k, v in prof.items()},
            "factorisation_ms": prof.get("eig", {}).get("ms_total", 0.0) / max(steps, 1)}


def bench_chol_e2e(pkg, gens, seed):
    n, lam = N_DIM, LAMBDA
    rng = np.random.default_rng(seed)
    lower, upper, guess = -BOX * np.ones(n), BOX * np.ones(n), rng.uniform(-BOX, BOX, n)
    alg = pkg.CholeskyCMAES(mfev=gens * lam, tol=0.0, stol=0.0, np=lam)
    t0 = time.perf_counter()
    sol = alg.optimize(pkg.Rosenbrock, lower, upper, guess)
    dt = time.perf_counter() - t0
    return (sol.n_evals // lam) / dt


def cpu_chol_sample(o, seed=1):
    """one whole generation of the reference's CholeskyCmaes at n = 1000, lambda = 4096 on one host core"""
    from oracle import oracle as O
    o.seed(seed)
    n = N_DIM
    h = o.cma(O.CHOLESKY, 10 ** 9, 1e-12, LAMBDA, stol=1e-12)
    h.init(O.ROSENBROCK, -BOX * np.ones(n), BOX * np.ones(n), o.uniform(-BOX, BOX, n))
    t0 = time.perf_counter()
    h.iterate()
    return 1.0 / (time.perf_counter() - t0)


LM_N, LM_LAMBDA = 10000, 64


def bench_lm_e2e(pkg, gens, seed):
    """LM-CMA-ES (cma/lm_cmaes.cpp), the reference's large-n variant: n = 10^4, lambda = 64, default memory (2 sqrt(n) = 200
    direction vectors), Ellipsoid, through the Python API (host buffers in, solution out)."""
    n, lam = LM_N, LM_LAMBDA
    lower, upper, guess = -BOX * np.ones(n), BOX * np.ones(n), np.random.default_rng(seed).uniform(-BOX, BOX, n)
    alg = pkg.LmCMAES(mfev=gens * lam, tol=0.0, np=lam)
    t0 = time.perf_counter()
    sol = alg.optimize(pkg.Ellipsoid, lower, upper, guess)
    dt = time.perf_counter() - t0
    return (sol.n_evals // lam) / dt


def cpu_lm_sample(o, gens=400, seed=1):  # the same 400 generations as bench_lm_e2e: the memory fills during the run
    from oracle import oracle as O
    o.seed(seed)
    n = LM_N
    h = o.cma(O.LM, 10 ** 9, 0.0, LM_LAMBDA)
    h.init(O.ELLIPSOID, -BOX * np.ones(n), BOX * np.ones(n), o.uniform(-BOX, BOX, n))
    t0 = time.perf_counter()
    for _ in range(gens):
        h.iterate()
    return gens / (time.perf_counter() - t0)


def bench_sansde_e2e(pkg, gens, seed):
    """SaNSDE (de/sansde.cpp), n = 1000, np = 1000, Rastrigin through the Python API: candidate-evaluations/s (the reference's
    count: one per candidate and generation)."""
    lower, upper = -DE_BOX * np.ones(DE_N), DE_BOX * np.ones(DE_N)
    alg = pkg.SANSDE(mfev=DE_NP * (gens + 1), np=DE_NP, tol=0.0)
    t0 = time.perf_counter()
    sol = alg.optimize(pkg.Rastrigin, lower, upper, lower)
    dt = time.perf_counter() - t0
    return sol.n_evals / dt, alg.rounds


def cpu_sansde_sample(o, gens=3, seed=1):
    from oracle import oracle as O
    o.seed(seed)
    h = o.sansde(10 ** 9, DE_NP, 0.0)
    h.init(O.RASTRIGIN, -DE_BOX * np.ones(DE_N), DE_BOX * np.ones(DE_N))
    t0 = time.perf_counter()
    for _ in range(gens):
        h.iterate()
    return gens * DE_NP / (time.perf_counter() - t0)


def bench_ssde_e2e(pkg, gens, seed, usede=False):
    """SSDE (de/ssde.cpp), n = 1000, npinit = 1000, Rastrigin through the Python API: objective evaluations/s (the reference's
    own count; with `usede` a failed first trial costs a second evaluation)."""
    lower, upper = -DE_BOX * np.ones(DE_N), DE_BOX * np.ones(DE_N)
    alg = pkg.SSDE(mfev=DE_NP * (gens + 1) * (2 if usede else 1), npinit=DE_NP, tol=0.0, patience=10 ** 9, usede=usede)
    t0 = time.perf_counter()
    sol = alg.optimize(pkg.Rastrigin, lower, upper, lower)
    dt = time.perf_counter() - t0
    return sol.n_evals / dt, alg.rounds


def cpu_ssde_sample(o, gens=2, seed=1, usede=False):
    from oracle import oracle as O
    o.seed(seed)
    h = o.ssde(10 ** 9, DE_NP, 0.0, usede=usede)
    h.init(O.RASTRIGIN, -DE_BOX * np.ones(DE_N), DE_BOX * np.ones(DE_N))
    f0 = h.geti("fev")
    t0 = time.perf_counter()
    for _ in range(gens):
        h.iterate()
    return (h.geti("fev") - f0) / (time.perf_counter() - t0)


def bench_xnes_e2e(pkg, gens, seed):
    """xNES (nes/xnes.cpp) at n = 1000 (np = 24 points), Rosenbrock, through the Python API: per generation one eigendecomposition
    (the reference's matrix exponential) and two n^3 tensor-core GEMMs."""
    n = N_DIM
    lower, upper = -BOX * np.ones(n), BOX * np.ones(n)
    alg = pkg.xNES(mfev=gens * (4 + int(3 * math.log(n))), tol=0.0)
    t0 = time.perf_counter()
    sol = alg.optimize(pkg.Rosenbrock, lower, upper, np.zeros(n))
    dt = time.perf_counter() - t0
    return (sol.n_evals // alg.np) / dt


def cpu_xnes_sample(o, seed=1):
    from oracle import oracle as O
    o.seed(seed)
    n = N_DIM
    h = o.xnes(10 ** 9, 0.0)
    h.init(O.ROSENBROCK, -BOX * np.ones(n), BOX * np.ones(n), np.zeros(n))
    t0 = time.perf_counter()
    h.iterate()
    return 1.0 / (time.perf_counter() - t0)


def bench_de_device(ctx, kind, steps, warmup, seed):
    from bboptpy_b200 import _cabi
    lo, hi = -DE_BOX * np.ones(DE_N), DE_BOX * np.ones(DE_N)
    g = _cabi.De(ctx, kind, DE_N, DE_NP, lo, hi, mfev=10 ** 7, seed=seed)
    obj = _cabi.Objective(ctx, DE_N, builtin=_cabi.OBJ_RASTRIGIN)
    g.init(obj)
    for _ in range(warmup):
        g.generation(obj)
    ctx.profile_reset()
    ctx.profile(True)
    times, rounds, evals = [], [], 0
    for _ in range(steps):
        npop = g.np
        ctx.flush_l2()
        ctx.timer_start()
        rep = g.generation(obj)
        times.append(ctx.timer_stop())
        rounds.append(rep.rounds)
        evals += npop
    prof = ctx.profile_read()
    ctx.profile(False)
    return {"evals_per_s": evals / (sum(times) / 1e3), "ms_per_generation": float(np.mean(times)), "rounds_mean": float(np.mean(rounds)),
            "round_kernel_ms_avg": prof.get("de_round", {}).get("ms_avg"), "fbest": rep.fbest}


def bench_de_batched(ctx, kind, runs, steps, warmup, seed):
    """R independent runs (same configuration, their own Philox seeds) per launch: the rows of all runs stream
    through one cooperative kernel, so the working set (R x 32 MB) is HBM resident, not L2 resident."""
    from bboptpy_b200 import _cabi
    lo, hi = -DE_BOX * np.ones(DE_N), DE_BOX * np.ones(DE_N)
    obj = _cabi.Objective(ctx, DE_N, builtin=_cabi.OBJ_RASTRIGIN)
    gs = [_cabi.De(ctx, kind, DE_N, DE_NP, lo, hi, mfev=10 ** 7, seed=seed + 1000 * r) for r in range(runs)]
    for g in gs:
        g.init(obj)
    for _ in range(warmup):
        _cabi.De.generation_batch(gs, obj, want_report=False)
    ctx.sync()
    ctx.timer_start()
    for _ in range(steps):
        _cabi.De.generation_batch(gs, obj, want_report=False)
    ms = ctx.timer_stop()
    reps = _cabi.De.generation_batch(gs, obj)
    return {"runs": runs, "evals_per_s": runs * DE_NP * steps / (ms * 1e-3), "ms_per_generation": ms / steps,
            "rounds_last": int(reps[0].rounds), "fbest_min": float(min(r.fbest for r in reps)),
            "working_set_mb": runs * 4 * DE_NP * DE_N * 8 / 1e6}


def bench_sep_device(ctx, steps, warmup, seed):
    """BASELINE config 3: separable CMA-ES, n = 10^4, lambda = 4096, Ellipsoid: the bandwidth-bound path."""
    from bboptpy_b200 import _cabi
    from bboptpy_b200.strategy import cma_constants, update_sigma
    n, lam = SEP_N, LAMBDA
    w, c = cma_constants(2, n, lam)
    rng = np.random.default_rng(seed)
    guess = rng.uniform(-BOX, BOX, n)
    g = _cabi.Cma(ctx, _cabi.CMA_SEP, n, lam, w, c, seed=seed)
    g.reset(guess)
    obj = _cabi.Objective(ctx, n, builtin=_cabi.OBJ_ELLIPSOID)
    state = {"sigma": 2.0, "fev": 0, "gen": 0}

    def step():
        state["fev"] += lam
        g.generation(obj, state["sigma"], state["gen"], c["ik"], state["fev"])
        rep = g.fetch()
        state["sigma"] = update_sigma(state["sigma"], rep.pslen, c, rep.ybw[0], rep.f_kth)
        state["gen"] += 1
        return rep

    for _ in range(warmup):
        step()
    ctx.profile_reset()
    ctx.profile(True)
    times = []
    for _ in range(steps):
        ctx.timer_start()
        rep = step()
        times.append(ctx.timer_stop())
    prof = ctx.profile_read()
    ctx.profile(False)
    ms = float(np.mean(times))
    return {"gens_per_s": 1e3 / ms, "ms_per_generation": ms, "fbest": rep.ybw[0], "bytes_per_generation": 8.0 * n * (lam + lam // 2),
            "kernel_share_of_step": {k: v["ms_total"] / sum(times) for k, v in prof.items()}}


def bench_sharded(ctx, comm, kind, lam, steps, warmup, seed):
    """BASELINE config 4: ONE run, candidates sharded over the GPUs of the communicator (fitness all-gather + rank-mu
    all-reduce per generation).  Strong scaling: lambda is fixed, every GPU owns lambda/world candidates."""
    from bboptpy_b200 import _cabi
    from bboptpy_b200.strategy import cma_constants, flops_per_generation, update_sigma
    n = N_DIM
    w, c = cma_constants(kind, n, lam)
    guess = np.random.default_rng(seed).uniform(-BOX, BOX, n)   # same seed on every rank: replicated state
    g = _cabi.Cma(ctx, kind, n, lam, w, c, seed=seed, shard=(comm.rank, comm.world))
    g.attach_comm(comm)
    g.reset(guess)
    obj = _cabi.Objective(ctx, n, builtin=_cabi.OBJ_ROSENBROCK)
    state = {"sigma": 2.0, "fev": 0, "gen": 0}

    def step():
        state["fev"] += lam
        g.generation(obj, state["sigma"], state["gen"], c["ik"], state["fev"])
        rep = g.fetch()
        state["sigma"] = update_sigma(state["sigma"], rep.pslen, c, rep.ybw[0], rep.f_kth)
        state["gen"] += 1
        return rep

    # ---- parity of the sharded exchange at THIS world size, before anything is timed: generation 0 of an unsharded handle
    # with the same Philox key must give bit-identical fitness and ranking (draws are keyed by the GLOBAL candidate id) and
    # mean / covariance within 1e-10; after the timed run the replicated state (m, C) must be bit-identical on every rank.
    parity = []
    one = _cabi.Cma(ctx, kind, n, lam, w, c, seed=seed)
    one.reset(guess)
    one.generation(obj, 2.0, 0, c["ik"], lam)
    r1 = one.fetch()
    g.generation(obj, 2.0, 0, c["ik"], lam)
    r2 = g.fetch()
    if not np.array_equal(one.get(_cabi.F_FITNESS), g.get(_cabi.F_FITNESS)):
        parity.append("generation-0 fitness differs from the unsharded handle")
    if not np.array_equal(one.rank_idx(), g.rank_idx()) or list(r1.ibw) != list(r2.ibw):
        parity.append("generation-0 ranking differs from the unsharded handle")
    tl = np.tril_indices(n)
    for name, f in (("xmean", _cabi.F_XMEAN), ("ps", _cabi.F_PS), ("pc", _cabi.F_PC)):
        a_, b_ = one.get(f), g.get(f)
        if np.abs(a_ - b_).max() > 1e-10 * np.abs(a_).max():
            parity.append(f"generation-0 {name} off by {np.abs(a_ - b_).max():.2e}")
    a_, b_ = one.get(_cabi.F_C)[tl], g.get(_cabi.F_C)[tl]
    if np.abs(a_ - b_).max() > 1e-10 * np.abs(a_).max():
        parity.append(f"generation-0 C off by {np.abs(a_ - b_).max():.2e}")
    del one
    g.reset(guess)
    for i in range(warmup):
        step()
        if i == 0:
            g.eig()
    g.fetch()
    ctx.profile_reset()
    ctx.profile(True)
    times = []
    for _ in range(steps):
        ctx.timer_start()
        rep = step()
        times.append(ctx.timer_stop())
    prof = ctx.profile_read()
    ctx.profile(False)
    import hashlib
    digest = hashlib.sha256(g.get(_cabi.F_XMEAN).tobytes() + g.get(_cabi.F_C).tobytes() + g.rank_idx().tobytes()).hexdigest()
    return {"times_ms": times, "fbest": rep.ybw[0], "flops": flops_per_generation(kind, n, lam),
            "shares": {k: v["ms_total"] / sum(times) for k, v in prof.items()}, "parity_issues": parity, "state_digest": digest}


def bench_bipop_batch(pkg, objective, box, comm, seed, runs=BIPOP_RUNS, n=BIPOP_N, mfev=BIPOP_MFEV):
    """BASELINE config 5: `runs` independent BIPOP-CMA-ES(Cmaes) sequences at n = 40, budget mfev each, through the Python
    API with host buffers (guesses in, solutions out).  With a communicator the runs are sharded r % world == rank and the
    batch ends with one best-of reduce; the caller takes the max of wall_s over ranks."""
    pkg.seed(seed)
    guesses = np.random.default_rng(seed).uniform(-box, box, (runs, n))
    batch = pkg.BiPopCMAESBatch(runs=runs, mfev=mfev, tol=1e-8)
    batch.set_timing(True)
    if comm is not None:
        batch.set_communicator(comm)
    t0 = time.perf_counter()
    sols = batch.optimize(objective, -box * np.ones(n), box * np.ones(n), guesses)
    fb, owner, xb = batch.best_of()
    wall = time.perf_counter() - t0
    st = batch.stats
    return {"local_runs": len(sols), "wall_s": wall, "device_s": st.device_ms / 1e3, "launches": int(st.launches),
            "base_runs": int(st.base_runs), "generations": int(st.generations), "evaluations": int(st.evaluations),
            "decompositions": int(st.decompositions), "f_median": float(np.median(batch.best_values)), "f_best": float(fb),
            "best_run": int(owner), "h2d_bytes": int(guesses.nbytes + 2 * n * 8), "d2h_bytes": int(len(sols) * n * 8)}


def cpu_bipop_sample(o, runs=2, seed=3):
    """The reference's own BiPopCmaes(Cmaes) on the same problem, `runs` sequences on one core."""
    from oracle import oracle as O
    n = BIPOP_N
    lo, hi = -5.12 * np.ones(n), 5.12 * np.ones(n)
    o.seed(seed)
    t0 = time.perf_counter()
    fev = 0
    for _ in range(runs):
        base = o.cma(O.VANILLA, 1000, 1e-8, 10)
        _, fe, _ = o.restart_optimize(1, base, BIPOP_MFEV, O.RASTRIGIN, lo, hi, o.uniform(-5.12, 5.12, n))
        fev += fe
    dt = time.perf_counter() - t0
    return {"runs_per_s": runs / dt, "evals_per_s": fev / dt, "runs": runs, "cores": 1}


def bench_de_e2e(pkg, cls, gens, seed):
    lo, hi = -DE_BOX * np.ones(DE_N), DE_BOX * np.ones(DE_N)
    if cls == "JADE":
        alg = pkg.JADE(mfev=DE_NP * (gens + 1), np=DE_NP, tol=0.0)
    else:
        alg = pkg.SHADE(mfev=DE_NP * (gens + 1), npinit=DE_NP, tol=0.0, npmin=DE_NP)
    t0 = time.perf_counter()
    sol = alg.optimize(pkg.Rastrigin, lo, hi, lo)
    dt = time.perf_counter() - t0
    return sol.n_evals / dt


def run_own(args):
    # libraries (NCCL's version banner) write to fd 1: keep the real stdout for the ONE JSON line
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    rank, world, local, dist = dist_setup(args.gpus)
    import __graft_entry__ as entry
    if rank == 0 and not os.path.exists(os.path.join(ROOT, "bboptpy_b200", "libbbgpu.so")):
        entry.build()
    barrier(dist, local)
    os.environ["BBOPT_DEVICE"] = str(local)
    import bboptpy_b200 as pkg
    from bboptpy_b200 import _cabi
    ctx = _cabi.Context(local)  # raises without a GPU: there is no CPU path to fall back to
    pkg.seed(1000 + rank)
    sampler = ClockSampler(local)
    barrier(dist, local)
    sampler.start()
    res = bench_cma_device(ctx, args.steps, args.warmup, seed=1 + rank)
    clocks = sampler.stop()
    barrier(dist, local)
    t_local = sum(res["times_ms"]) / 1e3
    t_max = all_max(dist, local, t_local)
    value = world * args.steps / t_max
    fbest, owner = all_min_with_owner(dist, local, rank, res["fbest"])
    # e2e through the public API
    e2e_gens = max(3, min(5 * args.steps, 100))  # enough generations to amortise optimize()'s allocations and uploads (~10 ms per call)
    bench_cma_e2e(pkg, 2, seed=1 + rank)  # first-call costs (cuSOLVER handle, workspace) are not part of a step
    barrier(dist, local)
    # the host side of optimize() (handle set-up, allocations, report reads) jitters by +-5 % from call to call and an
    # occasional call takes 30 % longer: five calls, the median one is reported, all five are listed
    e_runs = sorted((bench_cma_e2e(pkg, e2e_gens, seed=1 + rank) for _ in range(5)), key=lambda r: r[1])
    e_val, e_dt, e_done, h2d, d2h, ndec = e_runs[2]
    e_dt_max = all_max(dist, local, e_dt)
    e2e_value = world * e_done / e_dt_max
    # DE half of the metric
    de = {}
    de_steps = max(10, min(5 * args.steps, 100))
    for name, kind in (("jade", 0), ("shade", 1)):
        r = bench_de_device(ctx, kind, de_steps, max(args.warmup, 3), seed=7 + rank)
        t_de = all_max(dist, local, DE_NP * de_steps / r["evals_per_s"])
        r["evals_per_s"] = world * DE_NP * de_steps / t_de
        r["e2e_evals_per_s"] = bench_de_e2e(pkg, name.upper(), 400, seed=7 + rank) * world  # 400 generations: amortises the handle set-up
        b = bench_de_batched(ctx, kind, DE_BATCH, max(5, de_steps // 4), 3, seed=7 + rank)
        t_b = all_max(dist, local, b["ms_per_generation"])
        b["evals_per_s"] = world * DE_BATCH * DE_NP / (t_b * 1e-3)
        r["batched"] = b
        de[name] = r
    # separable CMA-ES on the bandwidth-bound configuration
    sep = bench_sep_device(ctx, max(5, args.steps // 2), 3, seed=1 + rank)
    t_sep = all_max(dist, local, sep["ms_per_generation"])
    sep["gens_per_s"] = world * 1e3 / t_sep
    # the reference's eigendecomposition-free variant at the headline shape
    chol = bench_chol_device(ctx, max(5, args.steps // 2), 3, seed=1 + rank)
    t_ch = all_max(dist, local, chol["ms_per_generation"])
    chol["gens_per_s"] = world * 1e3 / t_ch
    bench_chol_e2e(pkg, 2, seed=1 + rank)
    chol["e2e_gens_per_s"] = world * 1.0 / all_max(dist, local, 1.0 / bench_chol_e2e(pkg, e2e_gens, seed=1 + rank))
    bench_lm_e2e(pkg, 20, seed=1 + rank)
    # (these end-to-end legs are short wall-clock measurements: the median of three calls each)
    med3 = lambda fn: sorted((fn() for _ in range(3)), key=lambda r: r[0] if isinstance(r, tuple) else r)[1]
    lm = {"e2e_gens_per_s": world * 1.0 / all_max(dist, local, 1.0 / med3(lambda: bench_lm_e2e(pkg, 400, seed=1 + rank))),
          "workload": f"LmCMAES n={LM_N} lambda={LM_LAMBDA} memory=200 Ellipsoid, 400 generations through the Python API "
                      "(cma/lm_cmaes.cpp; the memory fills during the run)"}
    bench_sansde_e2e(pkg, 5, seed=1 + rank)
    sd_rate, sd_rounds = med3(lambda: bench_sansde_e2e(pkg, 200, seed=1 + rank))
    de["sansde"] = {"e2e_evals_per_s": world * DE_NP * 1.0 / all_max(dist, local, DE_NP / sd_rate), "rounds_last": sd_rounds,
                    "workload": "SANSDE n=1000 np=1000 Rastrigin, 200 generations through the Python API (de/sansde.cpp)"}
    bench_ssde_e2e(pkg, 5, seed=1 + rank)
    ss_rate, ss_rounds = med3(lambda: bench_ssde_e2e(pkg, 200, seed=1 + rank))
    ss2_rate, ss2_rounds = med3(lambda: bench_ssde_e2e(pkg, 100, seed=1 + rank, usede=True))
    de["ssde"] = {"e2e_evals_per_s": world * DE_NP * 1.0 / all_max(dist, local, DE_NP / ss_rate), "rounds_last": ss_rounds,
                  "usede_e2e_evals_per_s": world * DE_NP * 1.0 / all_max(dist, local, DE_NP / ss2_rate), "usede_rounds_last": ss2_rounds,
                  "workload": "SSDE n=1000 npinit=1000 Rastrigin, 200 generations (100 with usede) through the Python API "
                              "(de/ssde.cpp; linear population reduction: the run ends in thousands of generations of a handful of candidates)"}
    bench_xnes_e2e(pkg, 3, seed=1 + rank)
    xn = {"e2e_gens_per_s": world * 1.0 / all_max(dist, local, 1.0 / bench_xnes_e2e(pkg, e2e_gens, seed=1 + rank)),
          "workload": "xNES n=1000 np=24 Rosenbrock through the Python API (nes/xnes.cpp; one eigendecomposition + two n^3 GEMMs per generation)"}
    # one large-lambda run sharded by candidates over all GPUs of the job (N > 1 only)
    sharded = None
    pycomm = None
    if world > 1:
        import torch
        import torch.distributed as tdist
        ids = [_cabi.Comm.unique_id() if rank == 0 else None, pkg.Communicator.unique_id() if rank == 0 else None]
        tdist.broadcast_object_list(ids, src=0)
        comm = _cabi.Comm(ctx, world, rank, ids[0])
        pycomm = pkg.Communicator(world, rank, ids[1])
        sharded = {}
        for kname, kind in (("cmaes", _cabi.CMA_FULL), ("active_cmaes", _cabi.CMA_ACTIVE)):
            barrier(dist, local)
            sr = bench_sharded(ctx, comm, kind, SHARD_LAMBDA, max(5, args.steps // 2), 3, seed=77)
            t_sh = all_max(dist, local, sum(sr["times_ms"]) / 1e3)
            gps = len(sr["times_ms"]) / t_sh
            digests = [None] * world
            tdist.all_gather_object(digests, (sr["state_digest"], sr["parity_issues"]))
            issues = [f"rank {r}: {m}" for r, (_, ms_) in enumerate(digests) for m in ms_]
            if len({d for d, _ in digests}) != 1:
                issues.append("replicated state (m, C, ranking) differs between ranks after the timed run")
            sharded[kname] = {"parity": "ok" if not issues else "; ".join(issues),
                              "parity_checks": "generation 0 vs an unsharded handle with the same Philox key on every rank (fitness and "
                                               "ranking bit-identical, m / p_sigma / p_c / C at 1e-10); sha256 of (m, C, ranking) equal on "
                                               f"all {world} ranks after the timed generations",
                              "lambda": SHARD_LAMBDA, "gpus": world, "gens_per_s": gps, "ms_per_generation": 1e3 / gps,
                              "scaling": "strong", "tflops": sr["flops"] * gps / 1e12, "share_of_step": sr["shares"],
                              "collectives": "ncclAllGather(fitness) + one peer-memory all-reduce kernel over NVLink (n*ld + ld doubles; "
                                             "ncclAllReduce when peer mapping is unavailable) per generation"}
        del comm
    # BASELINE config 5: a batch of independent BIPOP runs, sharded over the ranks (strong scaling: the batch is fixed)
    bipop = {}
    bench_bipop_batch(pkg, pkg.Rastrigin, 5.12, None, seed=5, runs=256, mfev=20000)  # first-call costs
    for oname, objective, box in (("rastrigin", pkg.Rastrigin, 5.12), ("ackley", pkg.Ackley, 32.768)):
        barrier(dist, local)
        b = bench_bipop_batch(pkg, objective, box, pycomm, seed=11)
        wall = all_max(dist, local, b["wall_s"])
        dev_s = all_max(dist, local, b["device_s"])
        tot = {k: int(all_sum(dist, local, b[k])) for k in ("base_runs", "generations", "evaluations", "decompositions", "launches")}
        b.update(tot)
        b.update({"runs": BIPOP_RUNS, "n": BIPOP_N, "mfev_per_run": BIPOP_MFEV, "gpus": world, "scaling": "strong", "wall_s": wall,
                  "device_s": dev_s, "runs_per_s": BIPOP_RUNS / wall, "evals_per_s": tot["evaluations"] / wall,
                  "run_generations_per_s": tot["generations"] / wall, "device_run_generations_per_s": tot["generations"] / dev_s,
                  "note": "BiPopCMAESBatch.optimize(): host guesses in, solutions out; restart schedule on the host, base runs as "
                          "CTAs of the batched engine (in-CTA tred2/tql2 at the reference's cadence); wall clock includes every "
                          "host round trip; bound: per-run dependency chains (latency / issue), not HBM or tensor peak"})
        bipop[oname] = b
    excl_ms = all_max(dist, local, res["excl_eig_ms"])  # every rank: a collective after the return below would never complete
    if rank != 0:
        return
    peak, peak_src = fp64_peak()
    hbm, hbm_src = hbm_peak()
    sg = res["prof"].get("sample_gemm", {})
    sg_ms = sg.get("ms_avg")
    achieved = 2.0 * LAMBDA * N_DIM * N_DIM / (sg_ms * 1e-3) / 1e12 if sg_ms else None
    ms_step = t_max * 1e3 / args.steps
    step_flops = res["flops"]
    shares = {k: v["ms_total"] / sum(res["times_ms"]) for k, v in res["prof"].items()}
    for name in ("jade", "shade"):
        de[name]["roofline"] = {"bound": "hbm", "achieved": de[name]["evals_per_s"] / world * 5 * 8 * DE_N / 1e9, "peak": hbm,
                                "unit": "GB/s", "frac": de[name]["evals_per_s"] / world * 5 * 8 * DE_N / 1e9 / hbm, "traffic": None,
                                "peak_source": hbm_src, "algorithmic_bytes_per_eval": 5 * 8 * DE_N,
                                "note": "single run: 8 MB swarm (L2 resident), one cooperative launch per generation, bounded by the "
                                        "sequential dependency rounds of the in-place update, not by bandwidth"}
        bb = de[name]["batched"]
        bb["roofline"] = {"bound": "hbm", "achieved": bb["evals_per_s"] / world * 5 * 8 * DE_N / 1e9, "peak": hbm, "unit": "GB/s",
                          "frac": bb["evals_per_s"] / world * 5 * 8 * DE_N / 1e9 / hbm, "traffic": None, "peak_source": hbm_src,
                          "algorithmic_bytes_per_eval": 5 * 8 * DE_N}
    sep["roofline"] = {"bound": "hbm", "achieved": sep["bytes_per_generation"] * sep["gens_per_s"] / world / 1e9, "peak": hbm,
                       "unit": "GB/s", "frac": sep["bytes_per_generation"] * sep["gens_per_s"] / world / 1e9 / hbm, "traffic": None,
                       "peak_source": hbm_src, "algorithmic_bytes_per_generation": sep["bytes_per_generation"]}
    sep["workload"] = f"SepCmaes n={SEP_N} lambda={LAMBDA} Ellipsoid adjustlr=True (BASELINE config 3); working set > L2"
    ncu, ncu_src = ncu_kernel_block("gemm_tma_kernel<1, 4, 7")
    rd, wr = to_bytes(ncu.get("dram__bytes_read.sum")), to_bytes(ncu.get("dram__bytes_write.sum"))
    dmma = ncu.get("sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active", (None, ""))[0]
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "step": STEP_DEF + "; sigma on the host from the 104-byte report",
                   "draws": "device Philox4x32-10 + Box-Muller", "parallelism": f"independent runs x{world}" if world > 1 else "single run",
                   "l2": "L2 flushed (256 MB memset) before every timed generation, outside the event bracket"},
        "clocks": clocks, "gpu_launches": int(res["launches"]),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "note": f"optimize(f, lower, upper, guess) through the pybind API, {e_done} generations, "
                        f"{ndec} eigendecompositions inside the timed region (reference cadence); median of five calls",
                "calls_gens_per_s": [round(r[2] / r[1], 2) for r in e_runs]},
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                     "frac": (achieved / peak) if achieved else None,
                     "traffic": (rd + wr) if rd is not None and wr is not None else None,
                     "traffic_source": (f"{ncu_src}: dram__bytes_read.sum + dram__bytes_write.sum of this kernel, one launch "
                                        "(ncu --set full)") if ncu_src else None,
                     "dmma_pipe_pct": dmma,
                     "kernel": "gemm_tma_kernel<FORM1,SAMPLE> (TMA-fed DMMA.8x8x4): X = m + sigma Z (B D)^T", "kernel_ms_avg": sg_ms,
                     "algorithmic_flops_per_launch": 2.0 * LAMBDA * N_DIM * N_DIM, "peak_source": peak_src,
                     "generation_excl_eig_tflops": step_flops / (excl_ms * 1e-3) / 1e12 * world,
                     "generation_excl_eig_frac_of_peak": step_flops / (excl_ms * 1e-3) / 1e12 / peak,
                     "kernel_share_of_step": shares},
        "eig": {"ms": res["eig_ms"], "cadence_generations": 1, "inside_value": True,
                "solver": "own: cooperative sytrd kernel + tridiagonal divide & conquer with the row-local Q^T kernel on the side stream + one DMMA GEMM, then the DMMA C^-1/2 rebuild (cuSOLVER syevd outside 256 <= n <= ~1700 or with BBGPU_EIG=cusolver)",
                "ms_per_generation_excl_eig": excl_ms,
                "gens_per_s_excl_eig": world / (excl_ms * 1e-3),
                "note": "value and e2e include one decomposition per generation (the reference's cadence here); ms is its "
                        "CUDA-event time inside the same timed region"},
        "best_of": {"f": fbest, "rank": owner},
        "de": de,
        "sep": sep,
        "cholesky": chol,
        "lm_cma": lm,
        "xnes": xn,
        "bipop_batch": bipop,
    }
    chol["workload"] = "CholeskyCmaes n=1000 lambda=4096 Rosenbrock (cma/cholesky_cmaes.cpp; no eigendecomposition: SYRK + blocked Cholesky)"
    chol["frac_of_fp64_peak"] = chol["algorithmic_flops_per_generation"] * chol["gens_per_s"] / world / 1e12 / peak
    if sharded is not None:
        line["sharded_run"] = sharded
    if world == 1:
        o, kind = cpu_oracle()
        cs = cpu_sample(o, gens=2, budget_s=30.0)
        line["cpu_baseline"] = {"value": cs["gens_per_s"], "unit": UNIT, "cores": 1, "kind": kind,
                                "sample": f"{cs['generations']} whole generation(s) of the CPU implementation at n=1000, lambda=4096 "
                                          "through the reference's own phases (incl. sigma and the eigendecomposition, which fires "
                                          "every generation here); nothing scaled",
                                "gens_per_s_excl_eig": cs["gens_per_s_excl_eig"], "t_eig_s": cs["t_eig"],
                                "t_generation_s": cs["t_generation"],
                                "de_jade_evals_per_s": cpu_de_sample(o, 0), "de_shade_evals_per_s": cpu_de_sample(o, 1),
                                "cholesky_gens_per_s": cpu_chol_sample(o),
                                "lm_cma_gens_per_s": cpu_lm_sample(o),
                                "xnes_gens_per_s": cpu_xnes_sample(o),
                                "de_sansde_evals_per_s": cpu_sansde_sample(o),
                                "de_ssde_evals_per_s": cpu_ssde_sample(o),
                                "de_ssde_usede_evals_per_s": cpu_ssde_sample(o, gens=1, usede=True),
                                "bipop_batch": cpu_bipop_sample(o)}
    os.write(real_stdout, (json.dumps(line) + "\n").encode())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "own" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_own(args)


if __name__ == "__main__":
    main()
