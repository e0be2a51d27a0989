"""Generate tests/golden/*.npz from the COMPILED REFERENCE (oracle/_ref/libbbo_ref.so).

Run in the build container (where /root/reference exists):
    python tests/golden/make_golden.py
The reference has no golden vectors, known-answer tests or fixtures of its own
(SURVEY.md section 4), so these are outputs of the reference itself, seeded through
effolkronium::random_static::seed (src/random.hpp:238-241).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))

CMA_CASES = [
    # name, kind, obj, n, lambda, box, bound, generations recorded
    ("vanilla_rosen10", O.VANILLA, O.ROSENBROCK, 10, 20, 10.0, False, (1, 2, 10, 60)),
    ("active_rosen10", O.ACTIVE, O.ROSENBROCK, 10, 20, 10.0, False, (1, 2, 10, 60)),
    ("sep_ellip10", O.SEP, O.ELLIPSOID, 10, 20, 10.0, False, (1, 2, 10, 60)),
    ("vanilla_rast40", O.VANILLA, O.RASTRIGIN, 40, 15, 5.12, False, (1, 5, 40)),
    ("active_ackley33_bound", O.ACTIVE, O.ACKLEY, 33, 64, 32.768, True, (1, 5, 20)),
    ("sep_sphere100_bound", O.SEP, O.SPHERE, 100, 32, 10.0, True, (1, 5, 20)),
]
DE_CASES = [
    ("jade_rast30", O.JADE, O.RASTRIGIN, 30, 50, 5.12, 10**6, (1, 5, 50)),
    ("shade_rast30", O.SHADE, O.RASTRIGIN, 30, 50, 5.12, 20000, (1, 5, 50)),
    ("jade_rosen10", O.JADE, O.ROSENBROCK, 10, 20, 10.0, 10**6, (1, 5, 50)),
    ("shade_ackley8", O.SHADE, O.ACKLEY, 8, 30, 32.768, 4000, (1, 5, 50)),
]


def main():
    O.build(("ref",))
    R = O.ref()
    assert R.kind == "reference"
    g = {}
    # --- random stack (libstdc++ through random.hpp)
    R.seed(123)
    g["rng_u32"] = R.draw(3, 64)
    R.seed(123)
    g["rng_normal"] = R.draw(0, 65)
    R.seed(123)
    g["rng_real"] = R.draw(1, 32, -10.0, 10.0)
    R.seed(123)
    g["rng_int1000"] = R.draw(2, 64, 0, 999)
    R.seed(123)
    g["rng_int1"] = R.draw(2, 8, 0, 0)
    np.savez(os.path.join(OUT, "rng.npz"), **g)

    # --- objectives
    g = {}
    rng = np.random.default_rng(2024)
    for obj in range(5):
        for n in (1, 2, 10, 33, 1000):
            x = rng.uniform(-5, 5, n)
            g[f"x_{obj}_{n}"] = x
            g[f"f_{obj}_{n}"] = np.array([R.eval(obj, x)])
    np.savez(os.path.join(OUT, "objectives.npz"), **g)

    # --- CMA trajectories
    for name, kind, obj, n, lam, box, bound, gens in CMA_CASES:
        g = {"meta": np.array([kind, obj, n, lam, box, int(bound)], dtype=np.float64)}
        R.seed(42)
        guess = R.uniform(-box, box, n)
        h = R.cma(kind, 10**9, 1e-12, lam, bound=bound)
        h.init(obj, -box * np.ones(n), box * np.ones(n), guess)
        g["guess"] = guess
        for it in range(1, max(gens) + 1):
            h.iterate()
            if it in gens:
                s = h.snapshot()
                for k in ("sigma", "xmean", "pc", "ps", "D", "C", "fit_val", "fit_idx"):
                    g[f"g{it}_{k}"] = np.asarray(s[k], dtype=np.float64)
                if kind != O.SEP:
                    g[f"g{it}_invsqrtC"] = s["invsqrtC"]
                    g[f"g{it}_B"] = s["B"]
                if it == 1:
                    g["g1_arx"] = s["arx"]
        np.savez(os.path.join(OUT, f"cma_{name}.npz"), **g)

    # --- README anchor: ActiveCMAES(10000, 1e-4, 20) on 10-D Rosenbrock (examples/multivariate.py)
    R.seed(42)
    guess = R.uniform(-10.0, 10.0, 10)
    h = R.cma(O.ACTIVE, 10000, 1e-4, 20)
    x, fev, conv = h.optimize(O.ROSENBROCK, -10 * np.ones(10), 10 * np.ones(10), guess)
    np.savez(os.path.join(OUT, "readme_active_rosen10.npz"), guess=guess, x=x,
             fev=np.array([fev]), conv=np.array([int(conv)]), f=np.array([R.eval(O.ROSENBROCK, x)]))

    # --- DE trajectories
    for name, kind, obj, n, np_, box, mfev, gens in DE_CASES:
        g = {"meta": np.array([kind, obj, n, np_, box, mfev], dtype=np.float64)}
        R.seed(7)
        h = R.de(kind, mfev, np_, 1e-12, h=10)
        h.init(obj, -box * np.ones(n), box * np.ones(n))
        x0, f0 = h.swarm()
        g["init_x"], g["init_f"] = x0, f0
        for it in range(1, max(gens) + 1):
            h.iterate()
            if it in gens:
                x, f = h.swarm()
                g[f"g{it}_f"] = f
                g[f"g{it}_x"] = x
                g[f"g{it}_arch"] = h.archive()
                if kind == O.JADE:
                    g[f"g{it}_mu"] = np.array([h.getd("mucr"), h.getd("muf")])
                else:
                    g[f"g{it}_MCR"], g[f"g{it}_MF"] = h.getv("MCR"), h.getv("MF")
                g[f"g{it}_counts"] = np.array([h.geti("np"), h.geti("fev"), h.geti("larch")], dtype=np.float64)
        np.savez(os.path.join(OUT, f"de_{name}.npz"), **g)

    # --- restart drivers
    g = {}
    for rk, rname in ((0, "ipop"), (1, "bipop")):
        for bk in (O.VANILLA, O.ACTIVE, O.SEP):
            R.seed(5)
            n = 6
            guess = R.uniform(-5, 5, n)
            base = R.cma(bk, 1000, 1e-8, 10)
            x, fev, _ = R.restart_optimize(rk, base, 30000, O.RASTRIGIN, -5.12 * np.ones(n),
                                           5.12 * np.ones(n), guess)
            g[f"{rname}_{bk}_x"] = x
            g[f"{rname}_{bk}_fev"] = np.array([fev])
            g[f"{rname}_{bk}_guess"] = guess
    np.savez(os.path.join(OUT, "restarts.npz"), **g)
    print("golden vectors written to", OUT)


if __name__ == "__main__":
    main()
