"""Shared helpers for the parity tests."""
import numpy as np

REL = 1e-10  # BASELINE.json north_star: mean, sigma, C and population within 1e-10 relative (fp64)


def assert_close(got, want, rel=REL, what=""):
    """Norm-wise relative agreement: max|got-want| <= rel * max|want| (arrays) or rel*|want| (scalars)."""
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    assert got.shape == want.shape, (what, got.shape, want.shape)
    scale = max(float(np.max(np.abs(want))) if want.size else 0.0, 1e-300)
    err = float(np.max(np.abs(got - want))) if want.size else 0.0
    assert np.isfinite(got).all(), what
    assert err <= rel * scale, f"{what}: max abs err {err:.3e} vs scale {scale:.3e} (rel {err / scale:.3e} > {rel})"


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
