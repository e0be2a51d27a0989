"""Strategy constants of the CMA-ES family, as the reference derives them
(src/multivariate/cma/base_cmaes.cpp:94-129, active_cmaes.cpp:48-64, sep_cmaes.cpp:44-58).

The host C++ classes compute the same numbers (bboptpy_b200/host/cma.cpp); this Python form exists
for callers that drive the C ABI directly (bench.py, multi-GPU drivers)."""
import math

import numpy as np


def cma_constants(kind, n, lam, alphacov=2.0, adjustlr=True, eigenrate=0.25):
    """kind: 0 Cmaes, 1 ActiveCmaes, 2 SepCmaes, 3 CholeskyCmaes (the base constants only, cma/cholesky_cmaes.cpp:40-52).
    Returns (weights, constants dict)."""
    mu = lam // 2
    w = np.array([math.log(0.5 * (lam + 1.0)) - math.log(i + 1.0) for i in range(mu)])
    s = 0.0
    for v in w:
        s += v
    w = w * (1.0 / s)
    lenw = 0.0
    for v in w:
        lenw = lenw + v * v
    mueff = 1.0 / lenw
    c = {"mueff": mueff, "mu": mu}
    c["chi"] = math.sqrt(n) * (1.0 - 1.0 / (4.0 * n) + 1.0 / (21.0 * n * n))
    c["cc"] = (4.0 + mueff / n) / (n + 4.0 + 2.0 * mueff / n)
    c["cs"] = (mueff + 2.0) / (5.0 + n + mueff)
    c["c1"] = 2.0 / ((1.3 + n) * (1.3 + n) + mueff)
    c["cmu"] = min(1.0 - c["c1"], 2.0 * (mueff - 2.0 + 1.0 / mueff) / ((2.0 + n) * (2.0 + n) + mueff))
    c["damps"] = 1.0 + c["cs"] + 2.0 * max(0.0, math.sqrt((mueff - 1.0) / (n + 1.0)) - 1.0)
    c["hlen"] = 10 + int(math.ceil((30.0 * n) / lam))
    c["ik"] = int(math.ceil(0.1 + lam / 4.0))
    if kind == 0:
        c["eigenfreq"] = eigenrate * lam / (c["c1"] + c["cmu"]) / n
    elif kind == 1:
        c["cm"], c["alphaold"] = 1.0, 0.5
        c["cc"] = (4.0 + 0.0 * mueff / n) / (n + 4.0 + 0.0 * 2.0 * mueff / n)
        c["cs"] = (mueff + 2.0) / (3.0 + n + mueff)
        c["c1"] = alphacov * min(1.0, lam / 6.0) / ((n + 1.3) * (n + 1.3) + mueff)
        c["cmu"] = min(1.0 - c["c1"], alphacov * (mueff - 2.0 + 1.0 / mueff) / ((2.0 + n) * (2.0 + n) + alphacov * mueff / 2.0))
        c["cneg"] = (1.0 - c["cmu"]) * (alphacov / 8.0) * mueff / (math.pow(n + 2.0, 1.5) + 2.0 * mueff)
        c["damps"] = 1.0 + c["cs"] + 2.0 * max(0.0, math.sqrt((mueff - 1.0) / (n + 1.0)) - 1.0)
        c["eigenfreq"] = eigenrate * (1.0 / (c["c1"] + c["cmu"] + c["cneg"])) / n
    elif kind == 3:
        pass
    elif kind == 4:
        raise ValueError("LM-CMA-ES constants depend on its constructor arguments: see host/cma.cpp LmCmaes::init")
    else:
        c["cc"] = 4.0 / (n + 4.0)
        c["cs"] = (mueff + 2.0) / (3.0 + n + mueff)
        c["damps"] = 1.0 + c["cs"] + 2.0 * max(0.0, math.sqrt((mueff - 1.0) / (n + 1.0)) - 1.0)
        ccov = 2.0 / ((n + math.sqrt(2.0)) * (n + math.sqrt(2.0)) * mueff)
        ccov += min(1.0, (2.0 * mueff - 1.0) / ((n + 2.0) * (n + 2.0) + mueff)) * (1.0 - 1.0 / mueff)
        if adjustlr:
            ccov *= ((n + 2.0) / 3.0)
        c["ccov"] = ccov
    return w, c


def update_sigma(sigma, pslen, c, f0, fk, it=0, fworst=1.0, fbest=0.0):
    """BaseCmaes::updateSigma (base_cmaes.cpp:176-189) from the device-computed ||p_sigma||."""
    sigma *= math.exp(min(1.0, (c["cs"] / c["damps"]) * (pslen / c["chi"] - 1.0)))
    if f0 == fk:
        sigma *= math.exp(0.2 + c["cs"] / c["damps"])
    if it >= c["hlen"] and fworst - fbest == 0.0:
        sigma *= math.exp(0.2 + c["cs"] / c["damps"])
    return sigma


def flops_per_generation(kind, n, lam):
    """ALGORITHMIC flops of one generation: sampling GEMM 2 n^2 lambda, lower-triangle
    rank-mu update mu n (n+1), recombination 2 n mu, C^-1/2 gemv 2 n^2; Active adds the mu Mahalanobis
    norms 2 n^2 mu and the negative half of the SYRK mu n (n+1).  The eigendecomposition is excluded."""
    mu = lam // 2
    f = 2.0 * n * n * lam + mu * n * (n + 1.0) + 2.0 * n * mu + 2.0 * n * n
    if kind == 1:
        f += mu * n * (n + 1.0) + 2.0 * n * n * mu
    return f
