"""Host-side logic of the multi-GPU paths, free of any device code so that it
runs under a CPU `gloo` group as well as under NCCL:

  * which candidates a rank owns in a candidate-sharded run (`shard_bounds`),
  * the per-candidate recombination weight / covariance coefficient derived from the replicated ranking
    (`shard_coefficients`: the numpy statement of csrc/cma.cu shard_coef_kernel),
  * the algebra of the two exchange steps (`sharded_update`: local partials -> all-reduce -> replicated finish),
  * the best-of reduce over independent runs (`best_of`).

The CUDA engine does the same steps with NCCL on device buffers (csrc/comm.cu); this module is what the
world_size-2 CPU tests exercise and what bench.py uses for its rank bookkeeping."""
import numpy as np


def shard_bounds(lam, world, rank):
    """Candidates [lo, hi) owned by `rank`; lambda must split evenly (the engine enforces the same)."""
    if lam % world != 0 or lam // world < 2:
        raise ValueError(f"lambda = {lam} does not split into {world} equal shards of at least 2")
    per = lam // world
    return rank * per, (rank + 1) * per


def shard_coefficients(rank_idx, lam, mu, weights, lo, hi, cpos, cneg=0.0, norms=None):
    """(coef, wmean, ycoeff) for the candidates [lo, hi): coef multiplies y y^T in the rank-mu (and active) sum,
    wmean multiplies x in the new mean.  cma/cmaes.cpp:139; cma/active_cmaes.cpp:115-132,146-158."""
    coef = np.zeros(hi - lo)
    wmean = np.zeros(hi - lo)
    active = norms is not None
    ycoeff = np.zeros(mu) if active else None
    for r in range(lam):
        c = int(rank_idx[r])
        yc = 0.0
        worst = active and r >= lam - mu
        i = lam - 1 - r
        if worst:
            top = norms[rank_idx[lam - mu + i]]
            bot = max(norms[rank_idx[lam - 1 - i]], 1e-8)
            yc = top / bot
            ycoeff[i] = yc
        if not (lo <= c < hi):
            continue
        if r < mu:
            wmean[c - lo] = weights[r]
            coef[c - lo] = cpos * weights[r]
        elif worst:
            coef[c - lo] = -(cneg * weights[i]) * yc
    return coef, wmean, ycoeff


def sharded_update(dist, X_loc, lo, hi, rank_idx, state, consts, kind, sigma, fev, lam, norms_loc=None):
    """One candidate-sharded updateDistribution(): `state` holds the replicated xmean, pc, ps, C (lower triangle
    authoritative), invsqrtC; X_loc the local candidates.  dist: a torch.distributed module with an initialised group
    (gloo or nccl) or None for a single rank.  Returns the new (xmean, ps, pc, C_lower, ycoeff)."""
    import torch
    n = X_loc.shape[1]
    mu = lam // 2
    w = consts["weights"]
    active = kind == 1
    norms = None
    if active:  # exchange step 1b: every rank needs every candidate's Mahalanobis norm
        norms = np.zeros(lam)
        norms[lo:hi] = norms_loc
        if dist is not None:
            t = torch.from_numpy(norms)
            dist.all_reduce(t)
            norms = t.numpy()
    cpos = consts["cmu"] + consts["cneg"] * (1.0 - consts["alphaold"]) if active else consts["cmu"]
    coef, wmean, ycoeff = shard_coefficients(rank_idx, lam, mu, w, lo, hi, cpos, consts.get("cneg", 0.0), norms)
    Y = (X_loc - state["xmean"]) / sigma
    red = np.concatenate([((Y * coef[:, None]).T @ Y).ravel(), wmean @ X_loc])
    if dist is not None:  # exchange step 2
        t = torch.from_numpy(red)
        dist.all_reduce(t)
        red = t.numpy()
    S, sm = red[:n * n].reshape(n, n), red[n * n:]
    xold = state["xmean"]
    cm = consts.get("cm", 1.0) if active else 1.0
    xmean = xold * (1.0 - cm) + sm * cm if active else sm
    dm = (xmean - xold) / (cm * sigma if active else sigma)
    cs, cc, c1, cmu, mueff, chi = (consts[k] for k in ("cs", "cc", "c1", "cmu", "mueff", "chi"))
    ps = (1.0 - cs) * state["ps"] + np.sqrt(cs * (2.0 - cs) * mueff) * (state["invsqrtC"] @ dm)
    denom = 1.0 - (1.0 - cs) ** (2.0 * fev / lam)
    hsig = 1.0 if np.linalg.norm(ps) / np.sqrt(denom) / chi < 1.4 + 2.0 / (n + 1.0) else 0.0
    pc = (1.0 - cc) * state["pc"] + hsig * np.sqrt(cc * (2.0 - cc) * mueff) * dm
    decay = 1.0 - c1 - cmu + (consts["cneg"] * consts["alphaold"] if active else 0.0)
    C = decay * state["C"] + c1 * (np.outer(pc, pc) + (1.0 - hsig) * cc * (2.0 - cc) * state["C"]) + S
    return xmean, ps, pc, np.tril(C), ycoeff


def best_of(dist, f_local, x_local):
    """(f_best, owner_rank, x_best) over the ranks of the group; ties go to the lowest rank."""
    import torch
    if dist is None:
        return float(f_local), 0, np.asarray(x_local, dtype=np.float64)
    world, rank = dist.get_world_size(), dist.get_rank()
    f = torch.zeros(world, dtype=torch.float64)
    f[rank] = float(f_local)
    dist.all_reduce(f)
    owner = int(torch.argmin(f).item())
    x = torch.from_numpy(np.array(x_local, dtype=np.float64))
    dist.broadcast(x, src=owner)
    return float(f[owner].item()), owner, x.numpy()
