"""Checkpoint / resume of the CMA-ES host classes: a run resumed from checkpoint() on a fresh object
continues bit for bit -- the free-running Philox stream is a function of (key, generation, candidate), and both travel in
the blob together with every host scalar and the device state of the kind."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pkg():
    import bboptpy_b200
    return bboptpy_b200


CASES = [("CMAES", dict(mfev=10 ** 6, tol=1e-12, np=16), ["xmean", "pc", "ps", "C", "B", "D", "invsqrtC"]),
         ("ActiveCMAES", dict(mfev=10 ** 6, tol=1e-12, np=16), ["xmean", "pc", "ps", "C", "B", "D", "invsqrtC"]),
         ("SepCMAES", dict(mfev=10 ** 6, tol=1e-12, np=16), ["xmean", "pc", "ps", "C", "D"]),
         ("CholeskyCMAES", dict(mfev=10 ** 6, tol=1e-12, stol=1e-12, np=16), ["xmean", "pc", "ps", "A"])]


@pytest.mark.parametrize("cls,kw,fields", CASES)
def test_resumed_run_continues_bit_for_bit(pkg, cls, kw, fields):
    n = 30
    lo, hi = -5 * np.ones(n), 5 * np.ones(n)
    guess = np.linspace(-3, 4, n)
    a = getattr(pkg, cls)(**kw)
    a.initialize(pkg.Rastrigin, lo, hi, guess)
    for _ in range(12):
        a.iterate()
    blob = a.checkpoint()
    assert isinstance(blob, bytes) and blob[:8] == b"BBOPTCK1"
    sol_at_checkpoint = a.solution()
    for _ in range(9):
        a.iterate()
    b = getattr(pkg, cls)(**kw)
    b.initialize(pkg.Rastrigin, lo, hi, guess + 1.0)  # another start: everything must come from the blob
    b.restore(blob)
    assert b.n_evals == 12 * 16 and np.array_equal(b.solution().x, sol_at_checkpoint.x)
    for _ in range(9):
        b.iterate()
    assert b.n_evals == a.n_evals and b.sigma == a.sigma
    for f in fields:
        assert np.array_equal(a.state(f), b.state(f)), f
    assert np.array_equal(a.state("arx"), b.state("arx"))
    assert np.array_equal(a.solution().x, b.solution().x)
    assert a.converged() == b.converged()


def test_restore_rejects_foreign_and_damaged_blobs(pkg):
    n = 8
    lo, hi = -5 * np.ones(n), 5 * np.ones(n)
    a = pkg.CMAES(mfev=10 ** 6, tol=1e-12, np=8)
    a.initialize(pkg.Sphere, lo, hi, np.ones(n))
    a.iterate()
    blob = a.checkpoint()
    s = pkg.SepCMAES(mfev=10 ** 6, tol=1e-12, np=8)
    s.initialize(pkg.Sphere, lo, hi, np.ones(n))
    with pytest.raises(Exception, match="another optimizer class"):
        s.restore(blob)
    c = pkg.CMAES(mfev=10 ** 6, tol=1e-12, np=8)
    c.initialize(pkg.Sphere, -5 * np.ones(n + 1), 5 * np.ones(n + 1), np.ones(n + 1))
    with pytest.raises(Exception, match="dimension"):
        c.restore(blob)
    d = pkg.CMAES(mfev=10 ** 6, tol=1e-12, np=8)
    d.initialize(pkg.Sphere, lo, hi, np.ones(n))
    with pytest.raises(Exception, match="truncated|corrupt"):
        d.restore(blob[:len(blob) // 2])
    with pytest.raises(Exception, match="not a bboptpy_b200 checkpoint"):
        d.restore(b"garbage")
    # a checkpoint taken with another population size (as inside an IPOP restart) re-creates the device state
    e = pkg.CMAES(mfev=10 ** 6, tol=1e-12, np=20)
    e.initialize(pkg.Sphere, lo, hi, np.ones(n))
    e.restore(blob)
    e.iterate()
    assert e.n_evals == 16
