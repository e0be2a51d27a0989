"""ORACLE (test infrastructure): ctypes wrappers over oracle/bbo_api.h.

Two libraries export the same C API:
  * ``ref()``  -> oracle/_ref/libbbo_ref.so : the unmodified reference sources
                  (/root/reference/src/...), compiled in place by ``oracle/Makefile``
  * ``port()`` -> oracle/libbbo_port.so     : our CPU restatement (oracle/port/bbo_port.cpp)

Only tests/, ``__graft_entry__.smoke()`` and bench.py's CPU-baseline legs may import this
module.  The product package (bboptpy_b200) never does.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_PATH = os.path.join(_HERE, "_ref", "libbbo_ref.so")
PORT_PATH = os.path.join(_HERE, "libbbo_port.so")

SPHERE, ELLIPSOID, ROSENBROCK, RASTRIGIN, ACKLEY = range(5)
VANILLA, ACTIVE, SEP, CHOLESKY, LM = range(5)
JADE, SHADE = range(2)

_c_dp = ctypes.c_void_p


def build(which=("port", "ref")):
    """Run the committed recipe.  'ref' is a no-op when /root/reference is absent."""
    for target in which:
        subprocess.run(["make", "-s", "-C", _HERE, target], check=True)


def _bind(lib):
    d, i, vp, cp = ctypes.c_double, ctypes.c_int, ctypes.c_void_p, ctypes.c_char_p
    lib.bbo_kind.restype = cp
    lib.bbo_seed.argtypes = [ctypes.c_uint32]
    lib.bbo_rng_draw.argtypes = [i, i, d, d, _c_dp]
    lib.bbo_eval.restype = d
    lib.bbo_eval.argtypes = [i, i, _c_dp]
    lib.bbo_cma_new.restype = vp
    lib.bbo_cma_new.argtypes = [i, i, d, i, d, i, d, d, i]
    lib.bbo_cma_free.argtypes = [vp]
    lib.bbo_cma_init.argtypes = [vp, i, i, _c_dp, _c_dp, _c_dp]
    lib.bbo_cma_set_params.argtypes = [vp, i, d, i]
    for name in ("sample", "evalsort", "update", "history", "iterate", "reset_z"):
        getattr(lib, "bbo_cma_" + name).argtypes = [vp]
    lib.bbo_cma_time_eig.argtypes = [vp]
    lib.bbo_cma_time_eig.restype = d
    lib.bbo_cma_converged.argtypes = [vp]
    lib.bbo_cma_converged.restype = i
    lib.bbo_cma_optimize.argtypes = [vp, i, i, _c_dp, _c_dp, _c_dp, _c_dp, ctypes.POINTER(i)]
    lib.bbo_cma_optimize.restype = i
    lib.bbo_cma_geti.argtypes = [vp, cp]
    lib.bbo_cma_geti.restype = i
    lib.bbo_cma_getd.argtypes = [vp, cp]
    lib.bbo_cma_getd.restype = d
    lib.bbo_cma_getv.argtypes = [vp, cp, _c_dp, i]
    lib.bbo_cma_getv.restype = i
    lib.bbo_cma_eig_log.argtypes = [vp, i]
    lib.bbo_cma_eig_log_count.argtypes = [vp]
    lib.bbo_cma_eig_log_count.restype = i
    lib.bbo_cma_eig_log_get.argtypes = [vp, i, ctypes.POINTER(i), _c_dp, _c_dp]
    lib.bbo_cma_eig_log_get.restype = i
    lib.bbo_restart_optimize.argtypes = [i, vp, i, d, i, d, i, i, d, i, i, _c_dp, _c_dp, _c_dp,
                                         _c_dp, ctypes.POINTER(i), _c_dp, i]
    lib.bbo_restart_optimize.restype = i
    lib.bbo_xnes_new.restype = vp
    lib.bbo_xnes_new.argtypes = [i, d, d, d]
    lib.bbo_xnes_free.argtypes = [vp]
    lib.bbo_xnes_init.argtypes = [vp, i, i, _c_dp, _c_dp, _c_dp]
    lib.bbo_xnes_iterate.argtypes = [vp]
    lib.bbo_xnes_converged.argtypes = [vp]
    lib.bbo_xnes_converged.restype = i
    lib.bbo_xnes_optimize.argtypes = [vp, i, i, _c_dp, _c_dp, _c_dp, _c_dp, ctypes.POINTER(i)]
    lib.bbo_xnes_optimize.restype = i
    lib.bbo_xnes_geti.argtypes = [vp, cp]
    lib.bbo_xnes_geti.restype = i
    lib.bbo_xnes_getd.argtypes = [vp, cp]
    lib.bbo_xnes_getd.restype = d
    lib.bbo_xnes_getv.argtypes = [vp, cp, _c_dp, i]
    lib.bbo_xnes_getv.restype = i
    lib.bbo_sansde_new.restype = vp
    lib.bbo_sansde_new.argtypes = [i, i, d, i, i, i, i]
    lib.bbo_sansde_free.argtypes = [vp]
    lib.bbo_sansde_init.argtypes = [vp, i, i, _c_dp, _c_dp, _c_dp]
    lib.bbo_sansde_iterate.argtypes = [vp]
    lib.bbo_sansde_converged.argtypes = [vp]
    lib.bbo_sansde_converged.restype = i
    lib.bbo_sansde_optimize.argtypes = [vp, i, i, _c_dp, _c_dp, _c_dp, _c_dp, ctypes.POINTER(i)]
    lib.bbo_sansde_optimize.restype = i
    lib.bbo_sansde_geti.argtypes = [vp, cp]
    lib.bbo_sansde_geti.restype = i
    lib.bbo_sansde_getd.argtypes = [vp, cp]
    lib.bbo_sansde_getd.restype = d
    lib.bbo_sansde_getv.argtypes = [vp, cp, _c_dp, i]
    lib.bbo_sansde_getv.restype = i
    lib.bbo_ssde_new.restype = vp
    lib.bbo_ssde_new.argtypes = [i, i, d, i, i, d, i, i, i]
    lib.bbo_ssde_free.argtypes = [vp]
    lib.bbo_ssde_init.argtypes = [vp, i, i, _c_dp, _c_dp, _c_dp]
    lib.bbo_ssde_iterate.argtypes = [vp]
    lib.bbo_ssde_converged.argtypes = [vp]
    lib.bbo_ssde_converged.restype = i
    lib.bbo_ssde_optimize.argtypes = [vp, i, i, _c_dp, _c_dp, _c_dp, _c_dp, ctypes.POINTER(i)]
    lib.bbo_ssde_optimize.restype = i
    lib.bbo_ssde_geti.argtypes = [vp, cp]
    lib.bbo_ssde_geti.restype = i
    lib.bbo_ssde_getd.argtypes = [vp, cp]
    lib.bbo_ssde_getd.restype = d
    lib.bbo_ssde_getv.argtypes = [vp, cp, _c_dp, i]
    lib.bbo_ssde_getv.restype = i
    lib.bbo_de_new.restype = vp
    lib.bbo_de_new.argtypes = [i, i, i, d, i, i, d, d, d, i, i]
    lib.bbo_de_free.argtypes = [vp]
    lib.bbo_de_init.argtypes = [vp, i, i, _c_dp, _c_dp, _c_dp]
    lib.bbo_de_iterate.argtypes = [vp]
    lib.bbo_de_converged.argtypes = [vp]
    lib.bbo_de_converged.restype = i
    lib.bbo_de_optimize.argtypes = [vp, i, i, _c_dp, _c_dp, _c_dp, _c_dp, ctypes.POINTER(i)]
    lib.bbo_de_optimize.restype = i
    lib.bbo_de_geti.argtypes = [vp, cp]
    lib.bbo_de_geti.restype = i
    lib.bbo_de_getd.argtypes = [vp, cp]
    lib.bbo_de_getd.restype = d
    lib.bbo_de_getv.argtypes = [vp, cp, _c_dp, i]
    lib.bbo_de_getv.restype = i
    return lib


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class Oracle:
    """One loaded oracle library (reference or port)."""

    def __init__(self, path):
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.lib = _bind(ctypes.CDLL(path))
        self.kind = self.lib.bbo_kind().decode()

    # ---- rng / objectives
    def seed(self, s):
        self.lib.bbo_seed(int(s))

    def draw(self, kind, count, a=0.0, b=1.0):
        out = np.empty(count, dtype=np.float64)
        self.lib.bbo_rng_draw(kind, count, float(a), float(b), out.ctypes.data)
        return out

    def normals(self, count):
        return self.draw(0, count)

    def uniform(self, a, b, count):
        return self.draw(1, count, a, b)

    def eval(self, obj, x):
        x = _f64(x)
        return self.lib.bbo_eval(obj, x.size, x.ctypes.data)

    def eval_rows(self, obj, X):
        X = _f64(X)
        return np.array([self.eval(obj, row) for row in X])

    def cma(self, kind, mfev, tol, np_, sigma0=2.0, bound=False, eigenrate=0.25, alphacov=2.0,
            adjustlr=True, stol=None, memory=0, rademacher=True, usenew=True):
        """kind CHOLESKY: `stol` (cholesky_cmaes.h:44) travels in the eigenrate slot of the C API; kind LM: memory /
        rademacher / usenew (lm_cmaes.h:62-70) in the eigenrate / alphacov / adjustlr slots."""
        if kind == CHOLESKY:
            eigenrate = 1e-8 if stol is None else stol
        if kind == LM:
            eigenrate, alphacov, adjustlr = float(memory), float(bool(rademacher)), bool(usenew)
        return OracleCma(self, kind, mfev, tol, np_, sigma0, bound, eigenrate, alphacov, adjustlr)

    def sansde(self, mfev, np_, tol, repaircr=True, crref=5, pupdate=50, crupdate=25):
        return OracleGeneric(self, "sansde", self.lib.bbo_sansde_new(mfev, np_, tol, int(repaircr), crref, pupdate, crupdate))

    def ssde(self, mfev, npinit, tol, patience=1000, npmin=4, ptop=0.11, h=100, usede=False, repaircr=True):
        return OracleGeneric(self, "ssde", self.lib.bbo_ssde_new(mfev, npinit, tol, patience, npmin, ptop, h, int(usede), int(repaircr)))

    def xnes(self, mfev, tol, a0=1.0, etamu=1.0):
        return OracleXnes(self, mfev, tol, a0, etamu)

    def de(self, kind, mfev, np_, tol, archive=True, repaircr=True, pelite=0.05, cdamp=0.1,
           sigma=0.07, h=100, npmin=4):
        return OracleDe(self, kind, mfev, np_, tol, archive, repaircr, pelite, cdamp, sigma, h, npmin)

    def restart_optimize(self, kind, base, mfev, obj, lower, upper, guess, sigma0=2.0, flag=True,
                         ksigmadec=1.6, boundlambda=True, maxlargeruns=9, kbudget=2.0, cap=256):
        lower, upper, guess = _f64(lower), _f64(upper), _f64(guess)
        n = lower.size
        x = np.empty(n)
        fev = ctypes.c_int(0)
        trace = np.zeros((cap, 5))
        rows = self.lib.bbo_restart_optimize(kind, base.h, mfev, sigma0, int(flag), ksigmadec,
                                             int(boundlambda), maxlargeruns, kbudget, obj, n,
                                             lower.ctypes.data, upper.ctypes.data, guess.ctypes.data,
                                             x.ctypes.data, ctypes.byref(fev), trace.ctypes.data, cap)
        return x, fev.value, trace[:min(rows, cap)]


class OracleCma:
    def __init__(self, o, kind, mfev, tol, np_, sigma0, bound, eigenrate, alphacov, adjustlr):
        self.o, self.lib, self.kind = o, o.lib, kind
        self.h = self.lib.bbo_cma_new(kind, mfev, tol, np_, sigma0, int(bound), eigenrate, alphacov,
                                      int(adjustlr))

    def __del__(self):
        try:
            self.lib.bbo_cma_free(self.h)
        except Exception:
            pass

    def init(self, obj, lower, upper, guess):
        lower, upper, guess = _f64(lower), _f64(upper), _f64(guess)
        self.lib.bbo_cma_init(self.h, obj, lower.size, lower.ctypes.data, upper.ctypes.data,
                              guess.ctypes.data)

    def set_params(self, np_, sigma, mfev):
        self.lib.bbo_cma_set_params(self.h, np_, sigma, mfev)

    def sample(self):
        self.lib.bbo_cma_sample(self.h)

    def evalsort(self):
        self.lib.bbo_cma_evalsort(self.h)

    def update(self):
        self.lib.bbo_cma_update(self.h)

    def history(self):
        self.lib.bbo_cma_history(self.h)

    def iterate(self):
        self.lib.bbo_cma_iterate(self.h)

    def reset_z(self):
        self.lib.bbo_cma_reset_z(self.h)

    def time_eig(self):
        return self.lib.bbo_cma_time_eig(self.h)

    def converged(self):
        return bool(self.lib.bbo_cma_converged(self.h))

    def eig_log(self, enable=True):
        """Record (fev, B, D) of every decomposition that fires from now on (eigenbasis-replay feed)."""
        self.lib.bbo_cma_eig_log(self.h, int(enable))

    def eig_events(self):
        out = []
        for k in range(self.lib.bbo_cma_eig_log_count(self.h)):
            fev = ctypes.c_int(0)
            n = self.lib.bbo_cma_eig_log_get(self.h, k, ctypes.byref(fev), None, None)
            B, D = np.empty((n, n)), np.empty(n)
            self.lib.bbo_cma_eig_log_get(self.h, k, ctypes.byref(fev), B.ctypes.data, D.ctypes.data)
            out.append((fev.value, B, D))
        return out

    def optimize(self, obj, lower, upper, guess):
        lower, upper, guess = _f64(lower), _f64(upper), _f64(guess)
        x = np.empty(lower.size)
        fev = ctypes.c_int(0)
        conv = self.lib.bbo_cma_optimize(self.h, obj, lower.size, lower.ctypes.data,
                                         upper.ctypes.data, guess.ctypes.data, x.ctypes.data,
                                         ctypes.byref(fev))
        return x, fev.value, bool(conv)

    def geti(self, name):
        return self.lib.bbo_cma_geti(self.h, name.encode())

    def getd(self, name):
        return self.lib.bbo_cma_getd(self.h, name.encode())

    def getv(self, name):
        n = self.lib.bbo_cma_getv(self.h, name.encode(), None, 0)
        if n < 0:
            raise KeyError(name)
        out = np.empty(n, dtype=np.float64)
        self.lib.bbo_cma_getv(self.h, name.encode(), out.ctypes.data, n)
        return out

    def mat(self, name):
        n = self.geti("n")
        return self.getv(name).reshape(-1, n)

    def snapshot(self):
        """Every state field the GPU engine mirrors, as a dict of numpy arrays / scalars."""
        s = {k: self.geti(k) for k in ("n", "lambda", "mu", "it", "fev", "mit", "hlen", "ik")}
        s.update({k: self.getd(k) for k in ("sigma", "mueff", "cc", "cs", "c1", "cmu", "damps",
                                            "chi", "fbest", "fworst")})
        for k in ("xmean", "xold", "pc", "ps", "weights", "fit_val", "fit_idx"):
            s[k] = self.getv(k)
        s["arx"] = self.mat("arx")
        if self.kind == CHOLESKY:
            s["A"] = self.mat("A")
            s["dmean"] = self.getv("dmean")
            return s
        if self.kind == LM:
            for k in ("memlen", "memsize", "nsteps", "t"):
                s[k] = self.geti(k)
            s["s"] = self.getd("s")
            for k in ("b", "d", "fp", "jarr", "larr"):
                s[k] = self.getv(k)
            s["pcmat"], s["vmat"] = self.mat("pcmat"), self.mat("vmat")
            return s
        s["D"] = self.getv("D")
        s["C"] = self.getv("C") if self.kind == SEP else self.mat("C")
        if self.kind != SEP:
            s["B"] = self.mat("B")
            s["invsqrtC"] = self.mat("invsqrtC")
            s["eigenlastev"] = self.geti("eigenlastev")
            s["eigenfreq"] = self.getd("eigenfreq")
        if self.kind == ACTIVE:
            s["ycoeff"] = self.getv("ycoeff")
            s["cneg"] = self.getd("cneg")
        if self.kind == SEP:
            s["ccov"] = self.getd("ccov")
        return s


class OracleGeneric:
    """An optimizer behind a bbo_<prefix>_{free,init,iterate,converged,optimize,geti,getd,getv} family."""

    def __init__(self, o, prefix, handle):
        self.o, self.lib, self.prefix, self.h = o, o.lib, prefix, handle

    def _fn(self, name):
        return getattr(self.lib, f"bbo_{self.prefix}_{name}")

    def __del__(self):
        try:
            self._fn("free")(self.h)
        except Exception:
            pass

    def init(self, obj, lower, upper, guess=None):
        lower, upper = _f64(lower), _f64(upper)
        guess = _f64(guess) if guess is not None else lower
        self._fn("init")(self.h, obj, lower.size, lower.ctypes.data, upper.ctypes.data, guess.ctypes.data)

    def iterate(self):
        self._fn("iterate")(self.h)

    def converged(self):
        return bool(self._fn("converged")(self.h))

    def optimize(self, obj, lower, upper, guess=None):
        lower, upper = _f64(lower), _f64(upper)
        guess = _f64(guess) if guess is not None else lower
        x = np.empty(lower.size)
        fev = ctypes.c_int(0)
        conv = self._fn("optimize")(self.h, obj, lower.size, lower.ctypes.data, upper.ctypes.data, guess.ctypes.data,
                                    x.ctypes.data, ctypes.byref(fev))
        return x, fev.value, bool(conv)

    def geti(self, name):
        return self._fn("geti")(self.h, name.encode())

    def getd(self, name):
        return self._fn("getd")(self.h, name.encode())

    def getv(self, name):
        n = self._fn("getv")(self.h, name.encode(), None, 0)
        if n < 0:
            raise KeyError(name)
        out = np.empty(n, dtype=np.float64)
        self._fn("getv")(self.h, name.encode(), out.ctypes.data, n)
        return out

    def mat(self, name):
        return self.getv(name).reshape(-1, self.geti("n"))


class OracleXnes:
    """nes/xnes.cpp behind bbo_xnes_*."""

    def __init__(self, o, mfev, tol, a0, etamu):
        self.o, self.lib = o, o.lib
        self.h = self.lib.bbo_xnes_new(mfev, tol, a0, etamu)

    def __del__(self):
        try:
            self.lib.bbo_xnes_free(self.h)
        except Exception:
            pass

    def init(self, obj, lower, upper, guess):
        lower, upper, guess = _f64(lower), _f64(upper), _f64(guess)
        self.lib.bbo_xnes_init(self.h, obj, lower.size, lower.ctypes.data, upper.ctypes.data, guess.ctypes.data)

    def iterate(self):
        self.lib.bbo_xnes_iterate(self.h)

    def converged(self):
        return bool(self.lib.bbo_xnes_converged(self.h))

    def optimize(self, obj, lower, upper, guess):
        lower, upper, guess = _f64(lower), _f64(upper), _f64(guess)
        x = np.empty(lower.size)
        fev = ctypes.c_int(0)
        conv = self.lib.bbo_xnes_optimize(self.h, obj, lower.size, lower.ctypes.data, upper.ctypes.data, guess.ctypes.data,
                                          x.ctypes.data, ctypes.byref(fev))
        return x, fev.value, bool(conv)

    def geti(self, name):
        return self.lib.bbo_xnes_geti(self.h, name.encode())

    def getd(self, name):
        return self.lib.bbo_xnes_getd(self.h, name.encode())

    def getv(self, name):
        n = self.lib.bbo_xnes_getv(self.h, name.encode(), None, 0)
        if n < 0:
            raise KeyError(name)
        out = np.empty(n, dtype=np.float64)
        self.lib.bbo_xnes_getv(self.h, name.encode(), out.ctypes.data, n)
        return out

    def mat(self, name):
        return self.getv(name).reshape(-1, self.geti("n"))


class OracleDe:
    def __init__(self, o, kind, mfev, np_, tol, archive, repaircr, pelite, cdamp, sigma, h, npmin):
        self.o, self.lib, self.kind = o, o.lib, kind
        self.h = self.lib.bbo_de_new(kind, mfev, np_, tol, int(archive), int(repaircr), pelite,
                                     cdamp, sigma, h, npmin)

    def __del__(self):
        try:
            self.lib.bbo_de_free(self.h)
        except Exception:
            pass

    def init(self, obj, lower, upper, guess=None):
        lower, upper = _f64(lower), _f64(upper)
        guess = _f64(guess if guess is not None else lower)
        self.lib.bbo_de_init(self.h, obj, lower.size, lower.ctypes.data, upper.ctypes.data,
                             guess.ctypes.data)

    def iterate(self):
        self.lib.bbo_de_iterate(self.h)

    def converged(self):
        return bool(self.lib.bbo_de_converged(self.h))

    def optimize(self, obj, lower, upper, guess=None):
        lower, upper = _f64(lower), _f64(upper)
        guess = _f64(guess if guess is not None else lower)
        x = np.empty(lower.size)
        fev = ctypes.c_int(0)
        conv = self.lib.bbo_de_optimize(self.h, obj, lower.size, lower.ctypes.data,
                                        upper.ctypes.data, guess.ctypes.data, x.ctypes.data,
                                        ctypes.byref(fev))
        return x, fev.value, bool(conv)

    def geti(self, name):
        return self.lib.bbo_de_geti(self.h, name.encode())

    def getd(self, name):
        return self.lib.bbo_de_getd(self.h, name.encode())

    def getv(self, name):
        n = self.lib.bbo_de_getv(self.h, name.encode(), None, 0)
        if n < 0:
            raise KeyError(name)
        out = np.empty(n, dtype=np.float64)
        self.lib.bbo_de_getv(self.h, name.encode(), out.ctypes.data, n)
        return out

    def swarm(self):
        n = self.geti("n")
        return self.getv("swarm_x").reshape(-1, n), self.getv("swarm_f")

    def archive(self):
        n = self.geti("n")
        return self.getv("arch").reshape(-1, n)

    def trace(self):
        """Tagged draws of the generation just run (port only)."""
        n = self.geti("n")
        t = {k: self.getv("t_" + k) for k in ("zcr", "cauchy", "ibest", "r1", "r2", "jrand", "slot",
                                              "succ", "ri", "pi", "trim", "F", "CR", "ftrial")}
        t["u"] = self.getv("t_u").reshape(-1, n)
        return t


_cache = {}


def ref():
    if "ref" not in _cache:
        _cache["ref"] = Oracle(REF_PATH)
    return _cache["ref"]


def port():
    if "port" not in _cache:
        _cache["port"] = Oracle(PORT_PATH)
    return _cache["port"]


def have_ref():
    return os.path.exists(REF_PATH)
