"""ctypes view of the C ABI in include/bbgpu.h (libbbgpu.so, built in-tree by build.py).

This is the thinnest possible binding: argument types, status checking, numpy in/out.  It is
what the parity tests and bench.py call so that they exercise exactly the symbols a host
language binding would.  There is no fallback of any kind: if the library is missing or no CUDA
device is usable, construction raises.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("BBGPU_LIB_PATH") or os.path.join(_HERE, "libbbgpu.so")  # override: A/B builds of the engine

OBJ_SPHERE, OBJ_ELLIPSOID, OBJ_ROSENBROCK, OBJ_RASTRIGIN, OBJ_ACKLEY = range(5)
CMA_FULL, CMA_ACTIVE, CMA_SEP, CMA_CHOLESKY, CMA_LM, CMA_XNES = range(6)
DE_JADE, DE_SHADE = range(2)
(F_XMEAN, F_XOLD, F_PC, F_PS, F_C, F_B, F_D, F_INVSQRTC, F_ARX, F_Z, F_FITNESS, F_RANKVAL, F_YCOEFF,
 F_WEIGHTS, F_NORMS, F_SHARD_RED, F_A, F_PCMAT, F_VMAT, F_LMB, F_LMD) = range(21)
COMM_ID_BYTES = 128

c_dp = ctypes.c_void_p


class BbgpuError(RuntimeError):
    pass


class CmaConfig(ctypes.Structure):
    _fields_ = [("kind", ctypes.c_int), ("n", ctypes.c_int), ("lambda_", ctypes.c_int), ("mu", ctypes.c_int),
                ("bound", ctypes.c_int), ("shard_rank", ctypes.c_int), ("shard_count", ctypes.c_int),
                ("seed", ctypes.c_uint64),
                ("cc", ctypes.c_double), ("cs", ctypes.c_double), ("c1", ctypes.c_double), ("cmu", ctypes.c_double),
                ("mueff", ctypes.c_double), ("chi", ctypes.c_double),
                ("cneg", ctypes.c_double), ("alphaold", ctypes.c_double), ("cm", ctypes.c_double),
                ("ccov", ctypes.c_double),
                ("lm_memory", ctypes.c_int), ("lm_rademacher", ctypes.c_int), ("lm_usenew", ctypes.c_int)]


class CmaReport(ctypes.Structure):
    _fields_ = [("ibw", ctypes.c_int32 * 4), ("ybw", ctypes.c_double * 4), ("f_kth", ctypes.c_double),
                ("pslen", ctypes.c_double), ("hsig", ctypes.c_int32), ("eig_info", ctypes.c_int32),
                ("d_min", ctypes.c_double), ("d_max", ctypes.c_double), ("tolx_max", ctypes.c_double),
                ("noeffect_axis", ctypes.c_int32), ("noeffect_coor", ctypes.c_int32), ("eig_resid", ctypes.c_double)]


class DeConfig(ctypes.Structure):
    _fields_ = [("kind", ctypes.c_int), ("n", ctypes.c_int), ("np", ctypes.c_int), ("archive", ctypes.c_int),
                ("repaircr", ctypes.c_int), ("h", ctypes.c_int), ("npmin", ctypes.c_int), ("mfev", ctypes.c_int),
                ("pelite", ctypes.c_double), ("cdamp", ctypes.c_double), ("sigma", ctypes.c_double),
                ("seed", ctypes.c_uint64)]


class DeDraws(ctypes.Structure):
    _fields_ = [("zcr", c_dp), ("cauchy", c_dp), ("ibest", c_dp), ("r1", c_dp), ("r2", c_dp), ("jrand", c_dp),
                ("slot", c_dp), ("u", c_dp), ("ri", c_dp), ("pi", c_dp), ("trim", c_dp), ("ntrim", ctypes.c_int)]


class DeReport(ctypes.Structure):
    _fields_ = [("np", ctypes.c_int32), ("larch", ctypes.c_int32), ("nsucc", ctypes.c_int32), ("fev", ctypes.c_int32),
                ("rounds", ctypes.c_int32), ("kslot", ctypes.c_int32), ("mucr", ctypes.c_double),
                ("muf", ctypes.c_double), ("fbest", ctypes.c_double), ("radius_m2", ctypes.c_double)]


class CmabRunParams(ctypes.Structure):
    _fields_ = [("lambda_", ctypes.c_int32), ("mu", ctypes.c_int32), ("mfev", ctypes.c_int32), ("mit", ctypes.c_int32),
                ("hlen", ctypes.c_int32), ("ik", ctypes.c_int32), ("sigma0", ctypes.c_double), ("tol", ctypes.c_double),
                ("cc", ctypes.c_double), ("cs", ctypes.c_double), ("c1", ctypes.c_double), ("cmu", ctypes.c_double),
                ("mueff", ctypes.c_double), ("damps", ctypes.c_double), ("chi", ctypes.c_double),
                ("eigenfreq", ctypes.c_double), ("seed", ctypes.c_uint64), ("weights", c_dp)]


class CmabRunStatus(ctypes.Structure):
    _fields_ = [("status", ctypes.c_int32), ("flag", ctypes.c_int32), ("it", ctypes.c_int32), ("fev", ctypes.c_int32),
                ("lambda_", ctypes.c_int32), ("eig_count", ctypes.c_int32), ("eig_due", ctypes.c_int32),
                ("hsig", ctypes.c_int32), ("ql_iters", ctypes.c_int32), ("generations_this_step", ctypes.c_int32),
                ("ibw", ctypes.c_int32 * 4), ("ybw", ctypes.c_double * 4), ("sigma", ctypes.c_double),
                ("fx_final", ctypes.c_double), ("pslen", ctypes.c_double), ("fkth", ctypes.c_double),
                ("hist_best", ctypes.c_double), ("hist_worst", ctypes.c_double)]


DEVICE_OBJECTIVE_FN = ctypes.CFUNCTYPE(ctypes.c_int, c_dp, ctypes.c_int64, ctypes.c_int, ctypes.c_int, c_dp, c_dp, c_dp)
HOST_OBJECTIVE_FN = ctypes.CFUNCTYPE(ctypes.c_int, c_dp, ctypes.c_int64, ctypes.c_int, ctypes.c_int, c_dp, c_dp)

_lib = None


def symbols_in_header():
    """Every function name declared in include/bbgpu.h (used by the CPU export test)."""
    import re
    hdr = open(os.path.join(os.path.dirname(_HERE), "include", "bbgpu.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(bbgpu_[a-z0-9_]+)\s*\(", hdr)))


def lib():
    """Load libbbgpu.so (no GPU needed to load; every compute entry point needs one)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise BbgpuError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`. "
                         "There is no CPU fallback.")
    L = ctypes.CDLL(LIB_PATH)
    i, i64, u64, d, vp = ctypes.c_int, ctypes.c_int64, ctypes.c_uint64, ctypes.c_double, ctypes.c_void_p
    P = ctypes.POINTER
    L.bbgpu_version.restype = i
    L.bbgpu_last_error.restype = ctypes.c_char_p
    L.bbgpu_ctx_create.argtypes = [i, vp, P(vp)]
    L.bbgpu_ctx_destroy.argtypes = [vp]
    L.bbgpu_ctx_sync.argtypes = [vp]
    L.bbgpu_ctx_stream.argtypes = [vp]
    L.bbgpu_ctx_stream.restype = vp
    L.bbgpu_ctx_device.argtypes = [vp]
    L.bbgpu_ctx_launch_count.argtypes = [vp]
    L.bbgpu_ctx_launch_count.restype = i64
    L.bbgpu_ctx_profile_enable.argtypes = [vp, i]
    L.bbgpu_ctx_profile_read.argtypes = [vp, i, P(d), P(i64)]
    L.bbgpu_ctx_profile_reset.argtypes = [vp]
    L.bbgpu_ctx_timer_start.argtypes = [vp]
    L.bbgpu_ctx_timer_stop.argtypes = [vp, P(ctypes.c_float)]
    L.bbgpu_ctx_flush_l2.argtypes = [vp]
    L.bbgpu_objective_builtin.argtypes = [vp, i, i, P(vp)]
    L.bbgpu_objective_device.argtypes = [vp, DEVICE_OBJECTIVE_FN, vp, i, P(vp)]
    L.bbgpu_objective_host.argtypes = [vp, HOST_OBJECTIVE_FN, vp, i, P(vp)]
    L.bbgpu_objective_destroy.argtypes = [vp]
    L.bbgpu_eval.argtypes = [vp, vp, vp, i64, i, vp]
    L.bbgpu_eval_host.argtypes = [vp, vp, vp, i, vp]
    L.bbgpu_cma_create.argtypes = [vp, P(CmaConfig), vp, vp, vp, P(vp)]
    L.bbgpu_cma_destroy.argtypes = [vp]
    L.bbgpu_cma_reset.argtypes = [vp, vp, i]
    L.bbgpu_cma_set.argtypes = [vp, i, vp, i64]
    L.bbgpu_cma_get.argtypes = [vp, i, vp, i64]
    L.bbgpu_cma_get_rank.argtypes = [vp, vp, i]
    L.bbgpu_cma_get_rows.argtypes = [vp, i, i64, i64, vp]
    L.bbgpu_cma_device_ptr.argtypes = [vp, i, P(vp), P(i64)]
    L.bbgpu_cma_sample.argtypes = [vp, d, u64, vp]
    L.bbgpu_cma_sample_dev.argtypes = [vp, d, vp]
    L.bbgpu_cma_evaluate.argtypes = [vp, vp]
    L.bbgpu_cma_rank.argtypes = [vp, i]
    L.bbgpu_cma_update.argtypes = [vp, d, i]
    L.bbgpu_cma_eig.argtypes = [vp, vp, vp]
    L.bbgpu_cma_probe.argtypes = [vp, d, i]
    L.bbgpu_cma_swarm_radius.argtypes = [vp, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double)]
    L.bbgpu_cma_xnes_update.argtypes = [vp, d, d, d, ctypes.POINTER(ctypes.c_double)]
    L.bbgpu_cma_lm_sample.argtypes = [vp, d, ctypes.c_uint64, i, c_dp, c_dp, c_dp]
    L.bbgpu_cma_lm_update.argtypes = [vp, d, i, i, i, c_dp]
    L.bbgpu_cma_fetch.argtypes = [vp, P(CmaReport)]
    L.bbgpu_cma_generation.argtypes = [vp, vp, d, u64, i, i]
    L.bbgpu_cma_attach_comm.argtypes = [vp, vp]
    L.bbgpu_cma_shard_uses_peer_memory.argtypes = [vp]
    L.bbgpu_cma_shard_partial.argtypes = [vp]
    L.bbgpu_cma_shard_finish.argtypes = [vp, d, i]
    L.bbgpu_cma_get_candidate.argtypes = [vp, i, vp]
    L.bbgpu_comm_unique_id.argtypes = [vp]
    L.bbgpu_comm_create.argtypes = [vp, i, i, vp, P(vp)]
    L.bbgpu_comm_destroy.argtypes = [vp]
    L.bbgpu_comm_rank.argtypes = [vp]
    L.bbgpu_comm_world.argtypes = [vp]
    L.bbgpu_comm_allreduce_sum.argtypes = [vp, vp, i64]
    L.bbgpu_comm_best_of.argtypes = [vp, vp, vp, i, P(d), P(i)]
    L.bbgpu_comm_best_of_host.argtypes = [vp, d, vp, i, P(d), P(i)]
    L.bbgpu_dgemm_host.argtypes = [vp, i, i, i, vp, vp, vp]
    L.bbgpu_dsyrk_host.argtypes = [vp, i, i, d, vp, vp, vp]
    L.bbgpu_rank_host.argtypes = [vp, i, vp, vp]
    L.bbgpu_philox_normal_host.argtypes = [vp, u64, u64, i, i, vp]
    L.bbgpu_bench_dgemm.argtypes = [vp, i, i, i, i, P(ctypes.c_float)]
    L.bbgpu_de_create.argtypes = [vp, P(DeConfig), vp, vp, P(vp)]
    L.bbgpu_de_destroy.argtypes = [vp]
    L.bbgpu_de_init.argtypes = [vp, vp, vp]
    L.bbgpu_de_generation.argtypes = [vp, vp, P(DeDraws), P(DeReport)]
    L.bbgpu_de_get_swarm.argtypes = [vp, vp, vp]
    L.bbgpu_de_get_archive.argtypes = [vp, vp, P(i)]
    L.bbgpu_de_get_memory.argtypes = [vp, vp, vp]
    L.bbgpu_de_get_trace.argtypes = [vp, vp, vp, vp, vp]
    L.bbgpu_de_phase_times.argtypes = [vp, vp]
    L.bbgpu_de_generation_batch.argtypes = [vp, i, vp, vp]
    L.bbgpu_cmab_create.argtypes = [vp, i, i, i, vp, vp, P(vp)]
    L.bbgpu_cmab_destroy.argtypes = [vp]
    L.bbgpu_cmab_limits.argtypes = [i, P(i), P(i), P(i)]
    L.bbgpu_cmab_start.argtypes = [vp, i, vp, P(CmabRunParams), vp, i]
    L.bbgpu_cmab_step.argtypes = [vp, vp, i, vp, i, i, vp, i]
    L.bbgpu_cmab_fetch.argtypes = [vp, P(CmabRunStatus)]
    L.bbgpu_cmab_get_best.argtypes = [vp, i, i, vp]
    L.bbgpu_cmab_get.argtypes = [vp, i, i, vp, i64]
    L.bbgpu_cmab_set.argtypes = [vp, i, i, vp, i64]
    L.bbgpu_cmab_set_sigma.argtypes = [vp, i, d]
    L.bbgpu_cmab_get_population.argtypes = [vp, i, vp, vp, vp]
    L.bbgpu_cmab_adopt_basis.argtypes = [vp, i, vp, vp]
    L.bbgpu_eig_small_host.argtypes = [vp, i, i, vp, vp, vp, vp]
    L.bbgpu_sytrd_host.argtypes = [vp, i, vp, vp, vp, vp, vp, i, P(ctypes.c_float)]
    L.bbgpu_stedc_host.argtypes = [vp, i, vp, vp, vp, vp, P(ctypes.c_float)]
    _lib = L
    return L


def check(status):
    if status != 0:
        raise BbgpuError(f"bbgpu status {status}: {lib().bbgpu_last_error().decode()}")


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


class Context:
    def __init__(self, device=0, stream=None):
        self.L = lib()
        self.h = ctypes.c_void_p()
        check(self.L.bbgpu_ctx_create(device, stream, ctypes.byref(self.h)))

    def close(self):
        if self.h:
            self.L.bbgpu_ctx_destroy(self.h)
            self.h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        check(self.L.bbgpu_ctx_sync(self.h))

    @property
    def stream(self):
        return self.L.bbgpu_ctx_stream(self.h)

    @property
    def launches(self):
        return self.L.bbgpu_ctx_launch_count(self.h)

    # ---- measurement helpers
    PROF_TAGS = ("sample_gemm", "cov_syrk", "norm_gemm", "eval", "rank", "philox", "eig", "update_misc", "sep_sample",
                 "de_round", "de_misc", "comm", "cmab_step")

    def profile(self, on=True):
        check(self.L.bbgpu_ctx_profile_enable(self.h, int(on)))

    def profile_reset(self):
        check(self.L.bbgpu_ctx_profile_reset(self.h))

    def profile_read(self):
        out = {}
        for tag, name in enumerate(self.PROF_TAGS):
            ms, cnt = ctypes.c_double(0), ctypes.c_int64(0)
            check(self.L.bbgpu_ctx_profile_read(self.h, tag, ctypes.byref(ms), ctypes.byref(cnt)))
            if cnt.value:
                out[name] = {"ms_total": ms.value, "launches": cnt.value, "ms_avg": ms.value / cnt.value}
        return out

    def timer_start(self):
        check(self.L.bbgpu_ctx_timer_start(self.h))

    def timer_stop(self):
        ms = ctypes.c_float(0)
        check(self.L.bbgpu_ctx_timer_stop(self.h, ctypes.byref(ms)))
        return ms.value

    def flush_l2(self):
        check(self.L.bbgpu_ctx_flush_l2(self.h))

    # ---- standalone kernels
    def dgemm(self, A, B):
        A, B = f64(A), f64(B)
        M, K = A.shape
        K2, N = B.shape
        assert K == K2
        C = np.empty((M, N))
        check(self.L.bbgpu_dgemm_host(self.h, M, N, K, A.ctypes.data, B.ctypes.data, C.ctypes.data))
        return C

    def dsyrk(self, Y, coef, C, alpha):
        Y, coef = f64(Y), f64(coef)
        C = f64(C).copy()
        K, n = Y.shape
        check(self.L.bbgpu_dsyrk_host(self.h, n, K, alpha, Y.ctypes.data, coef.ctypes.data, C.ctypes.data))
        return C

    def rank(self, f):
        f = f64(f)
        idx = np.empty(f.size, dtype=np.int32)
        check(self.L.bbgpu_rank_host(self.h, f.size, f.ctypes.data, idx.ctypes.data))
        return idx

    def philox_normal(self, seed, generation, rows, cols):
        z = np.empty((rows, cols))
        check(self.L.bbgpu_philox_normal_host(self.h, seed, generation, rows, cols, z.ctypes.data))
        return z

    def eig_small(self, A):
        """The batched engine's in-CTA tred2/tql2 on a stack of symmetric matrices [count, n, n]."""
        A = f64(A)
        count, n, _ = A.shape
        B, D, info = np.empty_like(A), np.empty((count, n)), np.empty(count, dtype=np.int32)
        check(self.L.bbgpu_eig_small_host(self.h, n, count, A.ctypes.data, B.ctypes.data, D.ctypes.data, info.ctypes.data))
        return B, D, info

    def sytrd(self, A, reps=1):
        """Cooperative-kernel tridiagonalisation: returns d, e, tau, V (column-major reflectors as an (n, n) array whose
        [c, i] entry is component c of reflector i) and the kernel time in ms."""
        A = f64(A)
        n = A.shape[0]
        d, e, tau, V = np.empty(n), np.empty(n - 1), np.empty(n - 1), np.empty((n, n))
        ms = ctypes.c_float(0)
        check(self.L.bbgpu_sytrd_host(self.h, n, A.ctypes.data, d.ctypes.data, e.ctypes.data, tau.ctypes.data, V.ctypes.data, reps,
                                      ctypes.byref(ms)))
        return d, e, tau, V.T.copy(), ms.value

    def stedc(self, d, e):
        """Tridiagonal divide & conquer on the device: eigenvalues ascending, Z with eigenvectors as COLUMNS, time in ms."""
        d, e = f64(d), f64(e)
        n = d.size
        w, Zt = np.empty(n), np.empty((n, n))
        ms = ctypes.c_float(0)
        check(self.L.bbgpu_stedc_host(self.h, n, d.ctypes.data, e.ctypes.data if n > 1 else None, w.ctypes.data, Zt.ctypes.data,
                                      ctypes.byref(ms)))
        return w, Zt.T.copy(), ms.value

    def bench_dgemm(self, M, N, K, reps=10):
        ms = ctypes.c_float(0)
        check(self.L.bbgpu_bench_dgemm(self.h, M, N, K, reps, ctypes.byref(ms)))
        return ms.value


class Objective:
    def __init__(self, ctx, n, builtin=None, device_fn=None, host_fn=None, user=None):
        self.ctx, self.L, self.n = ctx, ctx.L, n
        self.h = ctypes.c_void_p()
        self._keep = None
        if builtin is not None:
            check(self.L.bbgpu_objective_builtin(ctx.h, builtin, n, ctypes.byref(self.h)))
        elif device_fn is not None:
            self._keep = device_fn if isinstance(device_fn, DEVICE_OBJECTIVE_FN) else DEVICE_OBJECTIVE_FN(device_fn)
            check(self.L.bbgpu_objective_device(ctx.h, self._keep, user, n, ctypes.byref(self.h)))
        elif host_fn is not None:
            self._keep = host_fn if isinstance(host_fn, HOST_OBJECTIVE_FN) else HOST_OBJECTIVE_FN(host_fn)
            check(self.L.bbgpu_objective_host(ctx.h, self._keep, user, n, ctypes.byref(self.h)))
        else:
            raise ValueError("one of builtin / device_fn / host_fn is required")

    @classmethod
    def from_python(cls, ctx, n, fn, batched=False):
        """Batched host-callback bridge around a Python callable f(x)->float (or f(X)->array)."""
        def cb(h_X, ld, n_, count, h_f, _user):
            try:
                X = np.ctypeslib.as_array(ctypes.cast(h_X, ctypes.POINTER(ctypes.c_double)), shape=(count, ld))[:, :n_]
                out = np.ctypeslib.as_array(ctypes.cast(h_f, ctypes.POINTER(ctypes.c_double)), shape=(count,))
                if batched:
                    out[:] = np.asarray(fn(X), dtype=np.float64)
                else:
                    for k in range(count):
                        out[k] = fn(X[k].copy())
                return 0
            except Exception:  # surfaced as BBGPU_ERR_OBJECTIVE
                import traceback
                traceback.print_exc()
                return 1
        return cls(ctx, n, host_fn=cb)

    def __del__(self):
        try:
            if self.h:
                self.L.bbgpu_objective_destroy(self.h)
        except Exception:
            pass

    def eval(self, X):
        X = f64(X)
        X2 = X.reshape(-1, self.n)
        f = np.empty(X2.shape[0])
        check(self.L.bbgpu_eval_host(self.ctx.h, self.h, X2.ctypes.data, X2.shape[0], f.ctypes.data))
        return f


class Comm:
    """One rank of an NCCL communicator (include/bbgpu.h, bbgpu_comm_*)."""

    @staticmethod
    def unique_id(L=None):
        L = L or lib()
        buf = ctypes.create_string_buffer(COMM_ID_BYTES)
        check(L.bbgpu_comm_unique_id(buf))
        return buf.raw

    def __init__(self, ctx, world, rank, unique_id):
        self.ctx, self.L, self.world, self.rank = ctx, ctx.L, world, rank
        assert len(unique_id) == COMM_ID_BYTES
        self.h = ctypes.c_void_p()
        check(self.L.bbgpu_comm_create(ctx.h, world, rank, ctypes.create_string_buffer(unique_id, COMM_ID_BYTES),
                                       ctypes.byref(self.h)))

    def __del__(self):
        try:
            if self.h:
                self.L.bbgpu_comm_destroy(self.h)
        except Exception:
            pass

    def best_of(self, d_f_ptr, d_x_ptr, n):
        f, owner = ctypes.c_double(0), ctypes.c_int(0)
        check(self.L.bbgpu_comm_best_of(self.h, d_f_ptr, d_x_ptr, n, ctypes.byref(f), ctypes.byref(owner)))
        return f.value, owner.value


class Cma:
    """Device state + phases of one CMA-ES run (see include/bbgpu.h)."""

    def __init__(self, ctx, kind, n, lam, weights, consts, bound=False, lower=None, upper=None, seed=0, shard=(0, 1)):
        self.ctx, self.L = ctx, ctx.L
        self.kind, self.n, self.lam, self.mu = kind, n, lam, len(weights)
        self.shard_rank, self.shards = shard
        self.lam_loc = lam // self.shards
        cfg = CmaConfig()
        cfg.kind, cfg.n, cfg.lambda_, cfg.mu, cfg.bound = kind, n, lam, self.mu, int(bound)
        cfg.shard_rank, cfg.shard_count, cfg.seed = self.shard_rank, self.shards, seed
        for k in ("cc", "cs", "c1", "cmu", "mueff", "chi", "cneg", "alphaold", "cm", "ccov"):
            setattr(cfg, k, float(consts.get(k, 0.0)))
        if kind == CMA_ACTIVE and "cm" not in consts:
            cfg.cm = 1.0
        cfg.lm_memory, cfg.lm_rademacher, cfg.lm_usenew = (int(consts.get(k, 0)) for k in ("lm_memory", "lm_rademacher", "lm_usenew"))
        self.lm_memory = cfg.lm_memory
        self.cfg = cfg
        w = f64(weights)
        lo = f64(lower) if lower is not None else None
        hi = f64(upper) if upper is not None else None
        self.h = ctypes.c_void_p()
        check(self.L.bbgpu_cma_create(ctx.h, ctypes.byref(cfg), w.ctypes.data, lo.ctypes.data if lo is not None else None,
                                      hi.ctypes.data if hi is not None else None, ctypes.byref(self.h)))

    def __del__(self):
        try:
            if self.h:
                self.L.bbgpu_cma_destroy(self.h)
        except Exception:
            pass

    def _len(self, field):
        n, lam = self.n, self.lam
        if field == F_Z and self.kind == CMA_LM:
            return ((lam + 1) // 2) * n
        full = self.kind != CMA_SEP
        ld = (n + 7) // 8 * 8
        return {F_XMEAN: n, F_XOLD: n, F_PC: n, F_PS: n, F_D: n, F_C: n * n if full else n, F_B: n * n,
                F_INVSQRTC: n * n, F_ARX: self.lam_loc * n, F_Z: self.lam_loc * n, F_FITNESS: lam, F_RANKVAL: lam,
                F_YCOEFF: self.mu, F_WEIGHTS: self.mu, F_NORMS: lam, F_SHARD_RED: n * ld + ld, F_A: n * n,
                F_PCMAT: self.lm_memory * n, F_VMAT: self.lm_memory * n, F_LMB: self.lm_memory, F_LMD: self.lm_memory}[field]

    def reset(self, guess, keep_offdiag=False):
        g = f64(guess)
        check(self.L.bbgpu_cma_reset(self.h, g.ctypes.data, int(keep_offdiag)))

    def set(self, field, arr):
        a = f64(arr).ravel()
        check(self.L.bbgpu_cma_set(self.h, field, a.ctypes.data, a.size))

    def get(self, field):
        out = np.empty(self._len(field))
        check(self.L.bbgpu_cma_get(self.h, field, out.ctypes.data, out.size))
        if field in (F_B, F_INVSQRTC, F_ARX, F_Z, F_A, F_PCMAT, F_VMAT) or (field == F_C and self.kind != CMA_SEP):
            return out.reshape(-1, self.n)
        return out

    def rank_idx(self):
        idx = np.empty(self.lam, dtype=np.int32)
        check(self.L.bbgpu_cma_get_rank(self.h, idx.ctypes.data, self.lam))
        return idx

    def xnes_update(self, sigma, etamu, etab):
        """xNES: natural-gradient update of mu and B after sample / evaluate / rank; returns G_sigma."""
        gs = ctypes.c_double(0)
        check(self.L.bbgpu_cma_xnes_update(self.h, sigma, etamu, etab, ctypes.byref(gs)))
        return gs.value

    def lm_sample(self, sigma, generation, jarr, memlen, z=None, i0=None):
        """LM kind: mirrored sampling through the live memory rows jarr[:memlen]; z (pairs x n) and i0 (pairs) replay the
        reference's base draws and subset starts."""
        j = np.ascontiguousarray(jarr, dtype=np.int32)
        if z is not None:
            z, i0 = f64(z), np.ascontiguousarray(i0, dtype=np.int32)
            assert z.size == ((self.lam + 1) // 2) * self.n and i0.size == (self.lam + 1) // 2
        check(self.L.bbgpu_cma_lm_sample(self.h, sigma, generation, memlen, j.ctypes.data if j.size else None,
                                         z.ctypes.data if z is not None else None, i0.ctypes.data if z is not None else None))

    def lm_update(self, sigma, do_memory, imin, memlen, jarr):
        j = np.ascontiguousarray(jarr, dtype=np.int32)
        check(self.L.bbgpu_cma_lm_update(self.h, sigma, int(do_memory), imin, memlen, j.ctypes.data if j.size else None))

    def swarm_radius(self):
        """Cholesky kind: (mean, M2) of the candidates' Euclidean norms (converged(), cholesky_cmaes.cpp:146-160)."""
        mean, m2 = ctypes.c_double(0), ctypes.c_double(0)
        check(self.L.bbgpu_cma_swarm_radius(self.h, ctypes.byref(mean), ctypes.byref(m2)))
        return mean.value, m2.value

    def sample(self, sigma, generation=0, z=None):
        if z is not None:
            z = f64(z)
            assert z.size == self.lam * self.n
            check(self.L.bbgpu_cma_sample(self.h, sigma, generation, z.ctypes.data))
        else:
            check(self.L.bbgpu_cma_sample(self.h, sigma, generation, None))

    def evaluate(self, obj):
        check(self.L.bbgpu_cma_evaluate(self.h, obj.h))

    def rank(self, ik):
        check(self.L.bbgpu_cma_rank(self.h, ik))

    def update(self, sigma, fev):
        check(self.L.bbgpu_cma_update(self.h, sigma, fev))

    def eig(self, B=None, D=None):
        if B is None:
            check(self.L.bbgpu_cma_eig(self.h, None, None))
        else:
            B, D = f64(B), f64(D)
            check(self.L.bbgpu_cma_eig(self.h, B.ctypes.data, D.ctypes.data))

    def probe(self, sigma, iaxis):
        check(self.L.bbgpu_cma_probe(self.h, sigma, iaxis))

    def fetch(self):
        rep = CmaReport()
        check(self.L.bbgpu_cma_fetch(self.h, ctypes.byref(rep)))
        return rep

    def generation(self, obj, sigma, generation, ik, fev_after):
        check(self.L.bbgpu_cma_generation(self.h, obj.h, sigma, generation, ik, fev_after))

    # ---- candidate-sharded handles
    def attach_comm(self, comm):
        check(self.L.bbgpu_cma_attach_comm(self.h, comm.h if comm is not None else None))
        self._comm = comm

    def uses_peer_memory(self):
        """True when the rank-mu exchange runs as one kernel over NVLink peer memory (else ncclAllReduce)."""
        return bool(self.L.bbgpu_cma_shard_uses_peer_memory(self.h))

    def shard_partial(self):
        check(self.L.bbgpu_cma_shard_partial(self.h))

    def shard_finish(self, sigma, fev):
        check(self.L.bbgpu_cma_shard_finish(self.h, sigma, fev))

    def candidate(self, index):
        x = np.empty(self.n)
        check(self.L.bbgpu_cma_get_candidate(self.h, index, x.ctypes.data))
        return x

    def device_ptr(self, field):
        p, ld = ctypes.c_void_p(), ctypes.c_int64(0)
        check(self.L.bbgpu_cma_device_ptr(self.h, field, ctypes.byref(p), ctypes.byref(ld)))
        return p.value, ld.value


class De:
    """Device state + generation of one JADE / L-SHADE run (see include/bbgpu.h)."""

    def __init__(self, ctx, kind, n, np_, lower, upper, mfev=10 ** 9, archive=True, repaircr=True, pelite=0.05,
                 cdamp=0.1, sigma=0.07, h=100, npmin=4, seed=0):
        self.ctx, self.L = ctx, ctx.L
        self.kind, self.n, self.npcap = kind, n, np_
        cfg = DeConfig()
        cfg.kind, cfg.n, cfg.np, cfg.archive, cfg.repaircr = kind, n, np_, int(archive), int(repaircr)
        cfg.h, cfg.npmin, cfg.mfev = h, npmin, mfev
        cfg.pelite, cfg.cdamp, cfg.sigma, cfg.seed = pelite, cdamp, sigma, seed
        self.cfg = cfg
        lo, hi = f64(lower), f64(upper)
        self.h = ctypes.c_void_p()
        check(self.L.bbgpu_de_create(ctx.h, ctypes.byref(cfg), lo.ctypes.data, hi.ctypes.data, ctypes.byref(self.h)))
        self.np = np_
        self._keep = None

    def __del__(self):
        try:
            if self.h:
                self.L.bbgpu_de_destroy(self.h)
        except Exception:
            pass

    def init(self, obj, x=None):
        if x is not None:
            x = f64(x)
            assert x.shape == (self.npcap, self.n)
        check(self.L.bbgpu_de_init(self.h, obj.h, x.ctypes.data if x is not None else None))
        self.np = self.npcap

    def generation(self, obj, draws=None, want_report=True):
        """draws: dict with the oracle recorder's tagged draws (zcr, cauchy, ibest, r1, r2, jrand, slot, u[, ri, pi, trim])."""
        rep = DeReport()
        dd = None
        if draws is not None:
            keep = {}
            dd = DeDraws()
            for k in ("zcr", "cauchy", "u", "pi"):
                if k in draws and draws[k] is not None:
                    keep[k] = f64(draws[k])
                    setattr(dd, k, keep[k].ctypes.data)
            for k in ("ibest", "r1", "r2", "jrand", "slot", "ri", "trim"):
                if k in draws and draws[k] is not None and len(draws[k]):
                    keep[k] = i32(np.asarray(draws[k]).astype(np.int64))
                    setattr(dd, k, keep[k].ctypes.data)
            dd.ntrim = len(draws.get("trim", [])) if draws.get("trim") is not None else 0
            self._keep = keep
        check(self.L.bbgpu_de_generation(self.h, obj.h, ctypes.byref(dd) if dd is not None else None,
                                         ctypes.byref(rep) if want_report else None))
        if want_report:
            self.np = rep.np
        return rep

    def swarm(self):
        x, f = np.empty((self.np, self.n)), np.empty(self.np)
        check(self.L.bbgpu_de_get_swarm(self.h, x.ctypes.data, f.ctypes.data))
        return x, f

    def archive(self):
        rows = ctypes.c_int(0)
        buf = np.empty((self.npcap, self.n))
        check(self.L.bbgpu_de_get_archive(self.h, buf.ctypes.data, ctypes.byref(rows)))
        return buf[:rows.value].copy()

    def memory(self):
        h = max(1, self.cfg.h)
        mcr, mf = np.empty(h), np.empty(h)
        check(self.L.bbgpu_de_get_memory(self.h, mcr.ctypes.data, mf.ctypes.data))
        return mcr, mf

    @staticmethod
    def generation_batch(runs, obj, want_report=True):
        """One generation of several independent runs (same configuration) in one launch."""
        L = runs[0].L
        hs = (ctypes.c_void_p * len(runs))(*[r.h for r in runs])
        reps = (DeReport * len(runs))() if want_report else None
        check(L.bbgpu_de_generation_batch(hs, len(runs), obj.h, reps))
        if want_report:
            for r, rep in zip(runs, reps):
                r.np = rep.np
        return reps

    def phase_times(self):
        t = np.zeros(64, dtype=np.uint64)
        check(self.L.bbgpu_de_phase_times(self.h, t.ctypes.data))
        return t

    def trace(self, np_gen):
        succ = np.empty(np_gen, dtype=np.int32)
        ft, F, CR = np.empty(np_gen), np.empty(np_gen), np.empty(np_gen)
        check(self.L.bbgpu_de_get_trace(self.h, succ.ctypes.data, ft.ctypes.data, F.ctypes.data, CR.ctypes.data))
        return succ, ft, F, CR


class CmaBatch:
    """Batched small-n CMA-ES engine (include/bbgpu.h, bbgpu_cmab_*): one CTA per independent run."""

    def __init__(self, ctx, n, runs, bound=False, lower=None, upper=None):
        self.ctx, self.L, self.n, self.runs = ctx, ctx.L, n, runs
        self.h = ctypes.c_void_p()
        lo = f64(lower) if lower is not None else None
        hi = f64(upper) if upper is not None else None
        check(self.L.bbgpu_cmab_create(ctx.h, n, runs, int(bound), lo.ctypes.data if lo is not None else None,
                                       hi.ctypes.data if hi is not None else None, ctypes.byref(self.h)))
        self._status = (CmabRunStatus * runs)()

    def __del__(self):
        try:
            if self.h:
                self.L.bbgpu_cmab_destroy(self.h)
        except Exception:
            pass

    @staticmethod
    def run_params(n, lam, mfev, sigma0, tol, seed, eigenrate=0.25):
        """BaseCmaes::init + Cmaes::init constants (base_cmaes.cpp:88-129, cmaes.cpp:48) for one base run."""
        from .strategy import cma_constants
        w, c = cma_constants(0, n, lam, eigenrate=eigenrate)
        w = f64(w)
        p = CmabRunParams(lam, c["mu"], mfev, mfev // lam, c["hlen"], c["ik"], sigma0, tol, c["cc"], c["cs"], c["c1"],
                          c["cmu"], c["mueff"], c["damps"], c["chi"], c["eigenfreq"], seed, w.ctypes.data)
        p._keep = w
        return p

    def start(self, run_ids, params, guesses, keep_offdiag=False):
        ids = i32(run_ids)
        arr = (CmabRunParams * len(ids))(*params)
        g = f64(guesses).reshape(len(ids), self.n)
        check(self.L.bbgpu_cmab_start(self.h, len(ids), ids.ctypes.data, arr, g.ctypes.data, int(keep_offdiag)))

    def step(self, obj, run_ids, quantum=0, max_gens=1, z=None, eig_replay=False):
        ids = i32(run_ids)
        zz = f64(z) if z is not None else None
        check(self.L.bbgpu_cmab_step(self.h, obj.h, len(ids), ids.ctypes.data, quantum, max_gens,
                                     zz.ctypes.data if zz is not None else None, int(eig_replay)))

    def fetch(self):
        check(self.L.bbgpu_cmab_fetch(self.h, self._status))
        return self._status

    def best(self, run0=0, count=None):
        count = self.runs - run0 if count is None else count
        x = np.empty((count, self.n))
        check(self.L.bbgpu_cmab_get_best(self.h, run0, count, x.ctypes.data))
        return x

    def get(self, run, field):
        n = self.n
        out = np.empty(n * n if field in (F_C, F_B, F_INVSQRTC) else n)
        check(self.L.bbgpu_cmab_get(self.h, run, field, out.ctypes.data, out.size))
        return out.reshape(n, n) if out.size == n * n else out

    def set(self, run, field, a):
        a = f64(a)
        check(self.L.bbgpu_cmab_set(self.h, run, field, a.ctypes.data, a.size))

    def set_sigma(self, run, sigma):
        check(self.L.bbgpu_cmab_set_sigma(self.h, run, float(sigma)))

    def population(self, run, lam):
        X, f, r = np.empty((lam, self.n)), np.empty(lam), np.empty(lam, dtype=np.int32)
        check(self.L.bbgpu_cmab_get_population(self.h, run, X.ctypes.data, f.ctypes.data, r.ctypes.data))
        return X, f, r

    def adopt_basis(self, run, B, D):
        B, D = f64(B), f64(D)
        check(self.L.bbgpu_cmab_adopt_basis(self.h, run, B.ctypes.data, D.ctypes.data))
