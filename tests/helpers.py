"""Shared test helpers (tests only)."""
import numpy as np

from oracle import reference_path as rp


def continua(profile, s0, s1, slab):
    """Phase-speed intervals in which some resonance (Alfven, cusp; for the slab also the
    sound point of F) sits inside the layer: there the ODE is singular, the reference's
    odeint output is solver noise, and no parity is claimed (the 'noise floor')."""
    x = np.linspace(s0, s1, 4001)
    rho, _, c2, _, vA2, _ = profile.speeds(x)
    cT2 = c2 * vA2 / (c2 + vA2)
    iv = [(np.sqrt(vA2.min()), np.sqrt(vA2.max())), (np.sqrt(cT2.min()), np.sqrt(cT2.max()))]
    if slab:
        iv.append((np.sqrt(c2.min()), np.sqrt(c2.max())))
    return iv


def regular_mask(W, intervals, margin=0.01):
    W = np.asarray(W)
    ok = np.ones(W.shape, bool)
    for lo, hi in intervals:
        ok &= (W < lo - margin) | (W > hi + margin)
    return ok


def cyl_profile(width=0.95, medium=rp.CYL_CORONAL):
    return rp.GaussianDensity(medium, width=width, const_B=True)


def slab_profile(width=0.9, medium=rp.SLAB_CORONAL):
    return rp.GaussianDensity(medium, width=width)
