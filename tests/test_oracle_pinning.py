"""Pin the CPU restatement (oracle/port) against the UNMODIFIED reference compiled in place
(oracle/_ref/libbbo_ref.so).  Bit-exact after every generation.  Skipped only where neither
/root/reference nor a prebuilt oracle/_ref exists."""
import numpy as np
import pytest

from oracle import oracle as O

CMA_FIELDS = ["xmean", "pc", "ps", "D", "C", "arx", "fit_val", "fit_idx"]


def _pair(ref_oracle, port_oracle, make):
    out = []
    for o in (ref_oracle, port_oracle):
        out.append(make(o))
    return out


def test_rng(ref_oracle, port_oracle):
    for kind, a, b in [(3, 0, 0), (0, 0, 0), (1, -10.0, 10.0), (1, 0.2, 0.2), (2, 0, 0), (2, 0, 49),
                       (2, 0, 999), (2, 0, 1999), (2, -5, 5)]:
        ref_oracle.seed(99)
        port_oracle.seed(99)
        assert np.array_equal(ref_oracle.draw(kind, 5000, a, b), port_oracle.draw(kind, 5000, a, b))


def test_objectives(ref_oracle, port_oracle):
    rng = np.random.default_rng(0)
    for obj in range(5):
        for n in (1, 2, 3, 31, 32, 33, 64, 100, 1000):
            for scale in (0.1, 5.0, 30.0):
                x = rng.uniform(-scale, scale, n)
                assert ref_oracle.eval(obj, x) == port_oracle.eval(obj, x)


@pytest.mark.parametrize("kind,obj,n,lam,gens,bound,box", [
    (O.VANILLA, O.ROSENBROCK, 10, 20, 300, False, 10.0),
    (O.ACTIVE, O.ROSENBROCK, 10, 20, 300, False, 10.0),
    (O.SEP, O.ELLIPSOID, 10, 20, 300, False, 10.0),
    (O.VANILLA, O.RASTRIGIN, 40, 15, 150, False, 5.12),
    (O.ACTIVE, O.ACKLEY, 40, 15, 150, False, 32.768),
    (O.VANILLA, O.ROSENBROCK, 60, 128, 20, True, 10.0),
    (O.ACTIVE, O.SPHERE, 33, 64, 40, True, 10.0),
    (O.SEP, O.ELLIPSOID, 1000, 64, 20, False, 10.0),
    (O.CHOLESKY, O.ROSENBROCK, 10, 20, 300, False, 10.0),
    (O.CHOLESKY, O.RASTRIGIN, 40, 15, 150, True, 5.12),
    (O.CHOLESKY, O.ELLIPSOID, 60, 128, 20, False, 10.0),
])
def test_cma_every_generation(ref_oracle, port_oracle, kind, obj, n, lam, gens, bound, box):
    def make(o):
        o.seed(42)
        guess = o.uniform(-box, box, n)
        h = o.cma(kind, 10 ** 9, 1e-12, lam, bound=bound)
        h.init(obj, -box * np.ones(n), box * np.ones(n), guess)
        return h
    r, p = _pair(ref_oracle, port_oracle, make)
    fields = CMA_FIELDS + (["B", "invsqrtC"] if kind in (O.VANILLA, O.ACTIVE) else []) + (["ycoeff"] if kind == O.ACTIVE else [])
    if kind == O.CHOLESKY:
        fields = [k for k in CMA_FIELDS if k not in ("C", "D")] + ["A", "dmean"]
    for g in range(gens):
        r.iterate()
        p.iterate()
        for k in fields:
            assert np.array_equal(r.getv(k), p.getv(k)), (g, k)
        for k in ("sigma", "fbest", "fworst"):
            assert r.getd(k) == p.getd(k), (g, k)
        assert r.converged() == p.converged()
        assert r.geti("flag") == p.geti("flag")


@pytest.mark.parametrize("n,lam,gens,bound,kw", [
    (10, 12, 200, False, dict()),                                      # defaults: Rademacher draws, the 2015 variant
    (10, 12, 200, False, dict(rademacher=False)),                      # Gaussian draws share _Z with selectSubset
    (30, 16, 150, True, dict(memory=5)),                               # a small memory: updateSet replaces vectors (algorithm 5)
    (30, 17, 150, False, dict(memory=6, usenew=False, rademacher=False)),  # the 2014 variant, odd lambda (last pair half used)
    (200, 24, 40, False, dict()),
])
def test_lm_cma_every_generation(ref_oracle, port_oracle, n, lam, gens, bound, kw):
    """LmCmaes (cma/lm_cmaes.cpp): port bit for bit after every generation, memory bookkeeping included."""
    def make(o):
        o.seed(42)
        guess = o.uniform(-5.0, 5.0, n)
        h = o.cma(O.LM, 10 ** 9, 1e-12, lam, bound=bound, **kw)
        h.init(O.RASTRIGIN, -5.12 * np.ones(n), 5.12 * np.ones(n), guess)
        return h
    r, p = _pair(ref_oracle, port_oracle, make)
    for k in ("cc", "cs", "c1", "damps", "ccc", "sqrt1mc1", "mueff"):
        assert r.getd(k) == p.getd(k), k
    for k in ("memsize", "nsteps", "t"):
        assert r.geti(k) == p.geti(k), k
    for g in range(gens):
        r.iterate()
        p.iterate()
        assert r.geti("memlen") == p.geti("memlen")
        for k in ("xmean", "pc", "arx", "fit_val", "fit_idx", "jarr", "larr", "b", "d", "fp"):
            assert np.array_equal(r.getv(k), p.getv(k)), (g, k)
        ml = r.geti("memlen")
        rows = r.getv("jarr")[:ml].astype(int)
        assert np.array_equal(r.mat("pcmat")[rows], p.mat("pcmat")[rows]), (g, "pcmat")
        assert np.array_equal(r.mat("vmat")[rows], p.mat("vmat")[rows]), (g, "vmat")
        for k in ("sigma", "s", "fbest", "fworst"):
            assert r.getd(k) == p.getd(k), (g, k)
        assert r.converged() == p.converged()


@pytest.mark.parametrize("obj,n,gens,a0,etamu", [(O.ROSENBROCK, 10, 300, 1.0, 1.0), (O.RASTRIGIN, 40, 100, 2.0, 0.5),
                                                  (O.ELLIPSOID, 100, 15, 1.0, 1.0)])
def test_xnes_every_generation(ref_oracle, port_oracle, obj, n, gens, a0, etamu):
    """xNES (nes/xnes.cpp): port bit for bit after every generation, matrix exponential (tred2 / tql2) included."""
    hs = []
    for o in (ref_oracle, port_oracle):
        o.seed(42)
        h = o.xnes(10 ** 9, 1e-12, a0, etamu)
        h.init(obj, -5 * np.ones(n), 5 * np.ones(n), np.ones(n))
        hs.append(h)
    r, p = hs
    assert np.array_equal(r.getv("u"), p.getv("u")) and r.geti("np") == p.geti("np")
    for g in range(gens):
        r.iterate()
        p.iterate()
        for k in ("mu", "B", "x", "z", "f", "Gdelta"):
            assert np.array_equal(r.getv(k), p.getv(k)), (g, k)
        for k in ("sigma", "Gsigma"):
            assert r.getd(k) == p.getd(k), (g, k)
        assert r.converged() == p.converged()
    res = []
    for o in (ref_oracle, port_oracle):
        o.seed(7)
        res.append(o.xnes(3000, 1e-9).optimize(O.SPHERE, -5 * np.ones(6), 5 * np.ones(6), np.ones(6)))
    assert np.array_equal(res[0][0], res[1][0]) and res[0][1:] == res[1][1:]


@pytest.mark.parametrize("obj,n,np_,gens,kw", [(O.RASTRIGIN, 10, 30, 120, dict()), (O.ROSENBROCK, 20, 40, 60, dict(repaircr=False, crref=3, pupdate=10, crupdate=5)),
                                                (O.ACKLEY, 50, 64, 30, dict(pupdate=7, crupdate=4))])
def test_sansde_every_generation(ref_oracle, port_oracle, obj, n, np_, gens, kw):
    """SaNSDE (de/sansde.cpp): port bit for bit after every generation, adaptation state included."""
    box = {O.ROSENBROCK: 10.0, O.RASTRIGIN: 5.12, O.ACKLEY: 32.768}[obj]
    hs = []
    for o in (ref_oracle, port_oracle):
        o.seed(42)
        h = o.sansde(10 ** 9, np_, 1e-12, **kw)
        h.init(obj, -box * np.ones(n), box * np.ones(n))
        hs.append(h)
    r, p = hs
    for g in range(gens):
        r.iterate()
        p.iterate()
        for k in ("swarm_x", "swarm_f", "swarm_cr"):
            assert np.array_equal(r.getv(k), p.getv(k)), (g, k)
        for k in ("p", "fp", "crm", "crrec", "crdeltaf", "fpns0", "fpns1", "fpnf0", "fpnf1"):
            assert r.getd(k) == p.getd(k), (g, k)
        for k in ("fev", "it", "pns0", "pns1", "pnf0", "pnf1"):
            assert r.geti(k) == p.geti(k), (g, k)
        if g % 10 == 9:
            assert r.converged() == p.converged()
            assert np.array_equal(r.getv("swarm_x"), p.getv("swarm_x"))
    res = []
    for o in (ref_oracle, port_oracle):
        o.seed(11)
        res.append(o.sansde(20000, 30, 1e-6).optimize(O.SPHERE, -5 * np.ones(6), 5 * np.ones(6)))
    assert np.array_equal(res[0][0], res[1][0]) and res[0][1:] == res[1][1:]


@pytest.mark.parametrize("obj,n,np_,gens,kw", [(O.RASTRIGIN, 10, 30, 150, dict()), (O.RASTRIGIN, 11, 30, 100, dict(usede=True, h=5)),
                                                (O.ROSENBROCK, 20, 40, 80, dict(usede=True, repaircr=False, ptop=0.3)),
                                                (O.ACKLEY, 51, 64, 40, dict(h=7, npmin=8)), (O.ROSENBROCK, 2, 12, 60, dict(usede=True))])
def test_ssde_every_generation(ref_oracle, port_oracle, obj, n, np_, gens, kw):
    """SSDE (de/ssde.cpp): port bit for bit after every generation -- libstdc++'s std::shuffle behind the random rotation,
    the dense transform restricted to its non-zero entries, both stages, memories, population reduction."""
    box = {O.ROSENBROCK: 10.0, O.RASTRIGIN: 5.12, O.ACKLEY: 32.768}[obj]
    mfev = np_ * (gens + 4) * (2 if kw.get("usede") else 1)
    hs = []
    for o in (ref_oracle, port_oracle):
        o.seed(42)
        h = o.ssde(mfev, np_, 1e-12, **kw)
        h.init(obj, -box * np.ones(n), box * np.ones(n))
        hs.append(h)
    r, p = hs
    assert np.array_equal(r.getv("swarm_x"), p.getv("swarm_x")) and r.geti("fev") == p.geti("fev")
    g = 0
    while r.geti("fev") < mfev and g < gens:
        r.iterate()
        p.iterate()
        g += 1
        for k in ("iwork", "swarm_f", "swarm_x", "L1", "L2", "LCR"):
            assert np.array_equal(r.getv(k), p.getv(k), equal_nan=True), (g, k)
        for k in ("np", "fev", "convcount", "k", "kcr"):
            assert r.geti(k) == p.geti(k), (g, k)
        assert r.converged() == p.converged()
    assert r.geti("np") < np_, "the population reduction was exercised"
    res = []
    for o in (ref_oracle, port_oracle):
        o.seed(11)
        res.append(o.ssde(20000, 30, 1e-6, usede=True).optimize(O.SPHERE, -5 * np.ones(6), 5 * np.ones(6)))
    assert np.array_equal(res[0][0], res[1][0]) and res[0][1:] == res[1][1:]


def test_cma_phases_one_by_one(ref_oracle, port_oracle):
    def make(o):
        o.seed(1)
        guess = o.uniform(-10, 10, 12)
        h = o.cma(O.ACTIVE, 10 ** 9, 1e-12, 16)
        h.init(O.ROSENBROCK, -10 * np.ones(12), 10 * np.ones(12), guess)
        return h
    r, p = _pair(ref_oracle, port_oracle, make)
    for _ in range(5):
        for phase in ("sample", "evalsort", "update", "history"):
            getattr(r, phase)()
            getattr(p, phase)()
            for k in CMA_FIELDS + ["B", "invsqrtC"]:
                assert np.array_equal(r.getv(k), p.getv(k)), (phase, k)
    assert r.geti("it") == p.geti("it") == 5


def test_cma_constants(ref_oracle, port_oracle):
    for kind in (O.VANILLA, O.ACTIVE, O.SEP):
        for n, lam in ((10, 20), (1000, 4096), (40, 15), (10000, 4096)):
            hs = []
            for o in (ref_oracle, port_oracle):
                h = o.cma(kind, 10 ** 7, 1e-9, lam)
                if n <= 1000 or kind == O.SEP:
                    h.init(O.SPHERE, -np.ones(n), np.ones(n), np.zeros(n))
                    hs.append(h)
            if len(hs) < 2:
                continue
            for k in ("mueff", "cc", "cs", "c1", "cmu", "damps", "chi"):
                assert hs[0].getd(k) == hs[1].getd(k)
            for k in ("mu", "mit", "hlen", "ik"):
                assert hs[0].geti(k) == hs[1].geti(k)
            assert np.array_equal(hs[0].getv("weights"), hs[1].getv("weights"))
            if kind != O.SEP:
                assert hs[0].getd("eigenfreq") == hs[1].getd("eigenfreq")
            if kind == O.ACTIVE:
                assert hs[0].getd("cneg") == hs[1].getd("cneg")
            if kind == O.SEP:
                assert hs[0].getd("ccov") == hs[1].getd("ccov")


@pytest.mark.parametrize("kind,obj,n,np_,gens,box,mfev,kw", [
    (O.JADE, O.RASTRIGIN, 30, 50, 200, 5.12, 10 ** 6, {}),
    (O.SHADE, O.RASTRIGIN, 30, 50, 200, 5.12, 20000, {}),
    (O.JADE, O.ROSENBROCK, 10, 20, 300, 10.0, 10 ** 6, {"archive": False, "repaircr": False}),
    (O.SHADE, O.ACKLEY, 8, 30, 100, 32.768, 4000, {"h": 5}),
    (O.JADE, O.RASTRIGIN, 200, 100, 25, 5.12, 10 ** 6, {}),
    (O.SHADE, O.SPHERE, 3, 10, 50, 5.12, 700, {}),
])
def test_de_every_generation(ref_oracle, port_oracle, kind, obj, n, np_, gens, box, mfev, kw):
    def make(o):
        o.seed(7)
        h = o.de(kind, mfev, np_, 1e-12, **kw)
        h.init(obj, -box * np.ones(n), box * np.ones(n))
        return h
    r, p = _pair(ref_oracle, port_oracle, make)
    for g in range(gens):
        r.iterate()
        p.iterate()
        (xa, fa), (xb, fb) = r.swarm(), p.swarm()
        assert np.array_equal(xa, xb) and np.array_equal(fa, fb), g
        assert np.array_equal(r.archive(), p.archive()), g
        if kind == O.JADE:
            assert r.getd("mucr") == p.getd("mucr") and r.getd("muf") == p.getd("muf")
        else:
            assert np.array_equal(r.getv("MCR"), p.getv("MCR")) and np.array_equal(r.getv("MF"), p.getv("MF"))
            assert r.geti("k") == p.geti("k")
        assert r.geti("np") == p.geti("np") and r.geti("fev") == p.geti("fev")
        assert r.converged() == p.converged()


def test_optimize_and_restart_drivers(ref_oracle, port_oracle):
    n = 6
    for kind in (O.JADE, O.SHADE):
        res = []
        for o in (ref_oracle, port_oracle):
            o.seed(11)
            res.append(o.de(kind, 30000, 40, 1e-6).optimize(O.ROSENBROCK, -10 * np.ones(n), 10 * np.ones(n)))
        assert np.array_equal(res[0][0], res[1][0]) and res[0][1:] == res[1][1:]
    for rk in (0, 1):
        for bk in (O.VANILLA, O.ACTIVE, O.SEP, O.CHOLESKY, O.LM):
            for flag in (True, False):
                res = []
                for o in (ref_oracle, port_oracle):
                    o.seed(9)
                    guess = o.uniform(-5, 5, n)
                    base = o.cma(bk, 1000, 1e-8, 10)
                    res.append(o.restart_optimize(rk, base, 20000, O.RASTRIGIN, -5.12 * np.ones(n),
                                                  5.12 * np.ones(n), guess, flag=flag))
                assert np.array_equal(res[0][0], res[1][0]) and res[0][1] == res[1][1]


def test_ssde_nan_memory_slot(ref_oracle, port_oracle):
    """Without repaircr the reference stores CR_i itself as the successful rate: CR_i = 0 makes the weighted Lehmer mean 0 / 0,
    and every later candidate that reads that slot crosses over with std::max(0., std::min(1., NaN)) = 1.  The port (and the
    device engine, tests/test_gpu_ssde.py) must walk through that state exactly as the reference does."""
    n, np_, gens = 12, 36, 150
    kw = dict(usede=True, repaircr=False, h=7, npmin=6)
    lo, hi = -5.12 * np.ones(n), 0.8 * 5.12 * np.ones(n)
    mfev = np_ * (gens + 4) * 2
    hs = []
    for o in (ref_oracle, port_oracle):
        o.seed(7)
        h = o.ssde(mfev, np_, 1e-12, **kw)
        h.init(O.RASTRIGIN, lo, hi)
        hs.append(h)
    r, p = hs
    seen_nan, g = False, 0
    while r.geti("fev") < mfev and g < gens:
        r.iterate()
        p.iterate()
        g += 1
        seen_nan = seen_nan or bool(np.isnan(r.getv("LCR")).any())
        for k in ("swarm_f", "swarm_x", "L1", "L2", "LCR"):
            assert np.array_equal(r.getv(k), p.getv(k), equal_nan=True), (g, k)
        assert r.geti("fev") == p.geti("fev") and r.geti("np") == p.geti("np")
    assert seen_nan, "the run reaches a NaN memory slot (generation 96 with this seed)"
