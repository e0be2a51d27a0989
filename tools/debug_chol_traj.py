"""Cholesky-CMA-ES trajectory under draw replay, oracle vs engine, to the reference's own stop: prints the drift and the
inputs of converged() per generation."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bboptpy_b200 as pkg
from oracle import oracle as O
P = O.port()
n, lam = 6, int(sys.argv[1]) if len(sys.argv) > 1 else 18
tol = 1e-8
lo, hi = -5.12 * np.ones(n), 5.12 * np.ones(n)
guess = np.random.default_rng(3).uniform(-5, 5, n)
P.seed(42)
h = P.cma(O.CHOLESKY, 10 ** 6, tol, lam, sigma0=1.25, stol=tol)
h.init(O.RASTRIGIN, lo, hi, guess)
alg = pkg.CholeskyCMAES(mfev=10 ** 6, tol=tol, stol=tol, np=lam, sigma0=1.25)
alg.set_draw_replay(42)
alg.initialize(pkg.Rastrigin, lo, hi, guess)
low = np.tril_indices(n)
for g in range(2000):
    h.iterate(); alg.iterate()
    Xo, Xg = h.mat("arx"), alg.state("arx").reshape(lam, n)
    ro, rg = np.sqrt((Xo * Xo).sum(1)), np.sqrt((Xg * Xg).sum(1))
    m2o, m2g = ((ro - ro.mean()) ** 2).sum(), ((rg - rg.mean()) ** 2).sum()
    fo = h.getv("fit_val")
    dA = np.abs(alg.state("A").reshape(n, n)[low] - h.mat("A")[low]).max() / np.abs(h.mat("A")).max()
    dm = np.abs(alg.state("xmean") - h.getv("xmean")).max() / np.abs(h.getv("xmean")).max()
    dX = np.abs(Xg - Xo).max()
    same_rank = np.array_equal(alg.state("rank"), h.getv("fit_idx"))
    co, cg = h.converged(), alg.converged()
    if g % 20 == 0 or co or cg or not same_rank:
        print(f"gen {g}: sigma {h.getd('sigma'):.3e} dsig {abs(alg.sigma / h.getd('sigma') - 1):.1e} dA {dA:.1e} dm {dm:.1e} dX {dX:.1e} "
              f"spread {fo[-1] - fo[0]:.3e} m2 {m2o:.3e}/{m2g:.3e} thr {(lam - 1) * tol * tol:.1e} rank_same {same_rank} conv {co}/{cg}")
    if co or cg:
        break
