"""Shared helpers for the parity tests."""
import numpy as np

REL = 1e-10  # BASELINE.json north_star: mean, sigma, C and population within 1e-10 relative (fp64)


ELEM_FLOOR = 1e-6   # element-wise companion: entries above ELEM_FLOOR * max|want| ...
ELEM_REL = 1e-8     # ... agree to ELEM_REL relative, entry by entry


def assert_close(got, want, rel=REL, what="", elem=True):
    """Two checks.  (1) Norm-wise: max|got-want| <= rel * max|want| (arrays) or rel*|want| (scalars) -- the north
    star's "1e-10 relative".  (2) Element-wise companion (arrays): every entry with |want| > ELEM_FLOOR * max|want| agrees
    to max(ELEM_REL, 100 rel) of ITS OWN magnitude, so small entries (off-diagonal C entries ~1e-3 of the diagonal, mean
    coordinates near zero) are held to a relative tolerance too.  Two fp64 implementations that sum K terms in
    different orders differ by ~sqrt(K) eps * max|.| in every entry, hence the floor: an entry 1e-6 of the scale can
    only be pinned to ~1e-8 of itself."""
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    assert got.shape == want.shape, (what, got.shape, want.shape)
    scale = max(float(np.max(np.abs(want))) if want.size else 0.0, 1e-300)
    err = float(np.max(np.abs(got - want))) if want.size else 0.0
    assert np.isfinite(got).all(), what
    assert err <= rel * scale, f"{what}: max abs err {err:.3e} vs scale {scale:.3e} (rel {err / scale:.3e} > {rel})"
    if elem and want.ndim > 0 and want.size > 1:
        big = np.abs(want) > ELEM_FLOOR * scale
        if big.any():
            erel = float(np.max(np.abs(got[big] - want[big]) / np.abs(want[big])))
            tol = max(ELEM_REL, 100.0 * rel)
            assert erel <= tol, f"{what}: element-wise rel err {erel:.3e} > {tol:.1e} on entries above {ELEM_FLOOR:g} * max"


def oracle_constants(h, kind):
    """Strategy constants of an initialised oracle optimizer, as the C ABI's bbgpu_cma_config wants them."""
    c = {k: h.getd(k) for k in ("cc", "cs", "c1", "cmu", "mueff", "chi")}
    if kind == 1:
        c.update(cneg=h.getd("cneg"), alphaold=h.getd("alphaold"), cm=h.getd("cm"))
    if kind == 2:
        c["ccov"] = h.getd("ccov")
    return c


def update_sigma_host(sigma, pslen, cs, damps, chi, f0, fk, it, hlen, fworst, fbest):
    """BaseCmaes::updateSigma (cma/base_cmaes.cpp:176-189) given the device-computed ||p_sigma||."""
    import math
    sigma *= math.exp(min(1.0, (cs / damps) * (pslen / chi - 1.0)))
    if f0 == fk:
        sigma *= math.exp(0.2 + cs / damps)
    if it >= hlen and fworst - fbest == 0.0:
        sigma *= math.exp(0.2 + cs / damps)
    return sigma
