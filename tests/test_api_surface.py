"""CPU-side checks of the Python surface: same class names, constructor keywords and defaults
as the reference's pybind layer (py/multivariate_py.cpp:99-171), and a loud failure without a GPU."""
import inspect

import numpy as np
import pytest

import __graft_entry__ as entry


@pytest.fixture(scope="module")
def pkg():
    entry.build()
    import bboptpy_b200
    return bboptpy_b200


def test_classes_and_hierarchy(pkg):
    for name in ("MultivariateSearch", "MultivariateSolution", "BaseCMAES", "CMAES", "ActiveCMAES", "SepCMAES",
                 "IPopCMAES", "BiPopCMAES", "JADE", "SHADE"):
        assert hasattr(pkg, name), name
    assert issubclass(pkg.CMAES, pkg.BaseCMAES) and issubclass(pkg.ActiveCMAES, pkg.CMAES)
    assert issubclass(pkg.SepCMAES, pkg.BaseCMAES) and issubclass(pkg.BaseCMAES, pkg.MultivariateSearch)
    for cls in (pkg.IPopCMAES, pkg.BiPopCMAES, pkg.JADE, pkg.SHADE):
        assert issubclass(cls, pkg.MultivariateSearch)
    for meth in ("optimize", "initialize", "iterate", "solution"):
        assert hasattr(pkg.MultivariateSearch, meth)


def test_constructor_keywords_match_the_reference(pkg):
    # keyword names and defaults from py/multivariate_py.cpp:103-171
    pkg.CMAES(mfev=100, tol=1e-4, np=8)
    pkg.CMAES(mfev=100, tol=1e-4, np=8, sigma0=2., bound=False, eigenrate=0.25)
    pkg.ActiveCMAES(mfev=100, tol=1e-4, np=8, sigma0=2., bound=False, alphacov=2., eigenrate=0.25)
    pkg.SepCMAES(mfev=100, tol=1e-4, np=8, sigma0=2., bound=False, adjustlr=True)
    base = pkg.CMAES(mfev=100, tol=1e-4, np=8)
    pkg.IPopCMAES(base=base, mfev=1000, print=False, sigma0=2., nipop=True, ksigmadec=1.6, boundlambda=True)
    pkg.BiPopCMAES(base=base, mfev=1000, print=False, sigma0=2., maxlargeruns=9, nbipop=True, ksigmadec=1.6, kbudget=2.)
    pkg.JADE(mfev=1000, np=20, tol=1e-6, archive=True, repaircr=True, pelite=0.05, cdamp=0.1, sigma=0.07)
    pkg.SHADE(mfev=1000, npinit=20, tol=1e-6, archive=True, repaircr=True, h=100, npmin=4)
    doc = pkg.SepCMAES.__init__.__doc__
    assert "adjustlr: bool = True" in doc  # the pybind default differs from the C++ default, as in the reference
    assert "= 0.25" in pkg.CMAES.__init__.__doc__ and "eigenrate" in pkg.CMAES.__init__.__doc__


def test_registry_objectives_exist(pkg):
    ids = [pkg.Sphere.id, pkg.Ellipsoid.id, pkg.Rosenbrock.id, pkg.Rastrigin.id, pkg.Ackley.id]
    assert ids == [0, 1, 2, 3, 4]


def test_no_gpu_means_a_loud_error_not_a_cpu_fallback(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    alg = pkg.ActiveCMAES(mfev=1000, tol=1e-4, np=20)
    n = 5
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        alg.optimize(pkg.Rosenbrock, -10 * np.ones(n), 10 * np.ones(n), np.zeros(n))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pkg.JADE(mfev=1000, np=20, tol=1e-6).optimize(lambda x: float((x * x).sum()), -np.ones(n), np.ones(n), np.zeros(n))


def test_python_strategy_constants_match_the_oracle(port_oracle):
    from bboptpy_b200.strategy import cma_constants
    for kind in (0, 1, 2):
        for n, lam in ((10, 20), (40, 15), (1000, 4096), (1000, 16384)):
            h = port_oracle.cma(kind, 10 ** 7, 1e-9, lam, adjustlr=True)
            h.init(0, -np.ones(n), np.ones(n), np.zeros(n))
            w, c = cma_constants(kind, n, lam)
            assert np.array_equal(w, h.getv("weights"))
            for k in ("mueff", "cc", "cs", "c1", "cmu", "damps", "chi"):
                assert c[k] == h.getd(k), (kind, n, lam, k)
            assert c["hlen"] == h.geti("hlen") and c["ik"] == h.geti("ik") and c["mu"] == h.geti("mu")
            if kind != 2:
                assert c["eigenfreq"] == h.getd("eigenfreq")
            if kind == 1:
                assert c["cneg"] == h.getd("cneg")
            if kind == 2:
                assert c["ccov"] == h.getd("ccov")
