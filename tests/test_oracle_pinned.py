"""Pin the oracle (CPU, no GPU):

  reference itself (executed in the build container, fixtures in tests/golden/)
      == oracle/reference_path.py at the reference's own solver tolerances
  reference's shipped root tables (Example data/*.pickle)
      satisfy the reference's acceptance test under the oracle
  oracle/reference_path.py at tight tolerances (the converged value of the reference's
  formulation)  == oracle/dispersion_rk.c (the fast C restatement used on whole grids)
"""
import os
import warnings

import numpy as np
import pytest

from helpers import continua, cyl_profile, regular_mask, slab_profile
from oracle import reference_path as rp
from oracle import rk_oracle as ork

warnings.filterwarnings("ignore")
TIGHT = dict(rtol=1e-12, atol=1e-30, shoot="linear")


def _models(kind, width=None, n_int_out=500):
    if kind == "cylinder_density":
        prof = cyl_profile(width or 0.95)
        return prof, {0: rp.CylinderDensity(prof, 0), 1: rp.CylinderDensity(prof, 1),
                      2: rp.CylinderDensity(prof, 2)}, (-1.0, -0.001), False
    prof = slab_profile(width or 0.9)
    return prof, {0: rp.SlabDensity(prof, "sausage", n_int_out), 1: rp.SlabDensity(prof, "kink", n_int_out)}, \
        (-1.0, 1.0), True


@pytest.mark.parametrize("name,kind,tol,stride", [("cylinder_density_coronal", "cylinder_density", 1e-7, 3),
                                                  ("slab_density_coronal", "slab_density", 1e-8, 4)])
def test_oracle_equals_executed_reference(golden_dir, name, kind, tol, stride):
    """D from the reference's own sausage()/kink() vs the restatement at the SAME solver
    settings (scipy defaults, fsolve, the reference's output grids): agreement is at the
    1e-10 level, i.e. the restatement evaluates the same functions in the same order."""
    g = np.load(os.path.join(golden_dir, "ref_D_%s.npz" % name))
    prof, models, (s0, s1), slab = _models(kind, n_int_out=None)
    iv = continua(prof, s0, s1, slab)
    n_checked = n_skipped = 0
    # a stride keeps the CPU suite short; the fixture holds all the points
    for mode, k, w, Dref in list(zip(g["mode"], g["k"], g["w"], g["D"]))[::stride]:
        Dor = rp.D(models[int(mode)], k, w)
        if np.isnan(Dref):
            assert np.isnan(Dor)            # the reference skipped it (m_e < 0): so must we
            n_skipped += 1
            continue
        if not regular_mask(w / k, iv):
            continue
        assert np.sign(Dor) == np.sign(Dref)
        assert abs(Dor - Dref) <= tol * abs(Dref), (mode, k, w, Dref, Dor)
        n_checked += 1
    assert n_checked >= 12 and n_skipped >= 3


@pytest.mark.parametrize("name,kind", [("cylinder_density_coronal", "cylinder_density"),
                                       ("slab_density_coronal", "slab_density")])
def test_reference_scan_and_bisection(golden_dir, name, kind):
    """The reference's own scan+bisection output (sol_ks/sol_omegas) over a few intervals:
    every mode it reports is a root of the oracle's D to within its 1 % acceptance band, and
    where it reports none the oracle has no accepted mode either."""
    g = np.load(os.path.join(golden_dir, "ref_scan_%s.npz" % name))
    _, models, _, _ = _models(kind)
    n = 0
    while "scan%d_k" % n in g.files:
        mode = int(g["scan%d_mode" % n][0]); k = float(g["scan%d_k" % n][0])
        freq = g["scan%d_freq" % n]; ws = g["scan%d_sol_ws" % n]
        mine = rp.find_roots(models[mode], k, freq, **TIGHT)
        # the reference additionally needs >2 (cylinder) points before a sign change counts
        if len(ws):
            for w in ws:
                e, i = rp.dispersion(models[mode], k, float(w), **TIGHT)
                assert rp.mismatch_percent(e, i) < 1.5
                assert np.min(np.abs(mine - w)) < 5e-3 * abs(w)
        n += 1
    assert n >= 2


def test_shipped_root_tables_pass_acceptance_under_oracle(golden_dir):
    """Example data/*.pickle = roots the reference accepted (<1 % mismatch).  With the same
    profile width the oracle must agree: nearly all of them are inside the 1 % band."""
    g = np.load(os.path.join(golden_dir, "ref_roots.npz"))
    total = inside = 0
    medians = []
    for fam, kind in (("cyl_coronal", "cylinder_density"), ("slab_coronal", "slab_density")):
        tags = sorted(set(f.split("_")[2] for f in g.files if f.startswith(fam)))
        for tag in tags:
            width = float(g["%s_%s_width" % (fam, tag)][0])
            prof, _, (s0, s1), slab = _models(kind, width)
            iv = continua(prof, s0, s1, slab)
            model = ork.make_model(kind, width=width)
            for mi, mode in ((0, "sausage"), (1, "kink")):
                k = g["%s_%s_%s_k" % (fam, tag, mode)][::3]
                w = g["%s_%s_%s_w" % (fam, tag, mode)][::3]
                if not len(k):
                    continue
                reg = regular_mask(w / k, iv, 0.0) & (np.abs(w / k) < 6)
                pct = []
                for kk, ww in zip(k[reg], w[reg]):
                    e, i = ork.point(model, mi, kk, ww)
                    pct.append(rp.mismatch_percent(e, i))
                pct = np.array(pct)
                total += len(pct)
                inside += int((pct < 1.5).sum())
                if len(pct) > 5:
                    medians.append(np.median(pct))
    assert total > 700
    assert inside / total > 0.90, (inside, total)
    assert max(medians) < 0.8          # accepted anywhere below 1 % -> median near 0.5 %


@pytest.mark.parametrize("kind", ["cylinder_density", "slab_density"])
def test_c_oracle_equals_converged_scipy_path(kind):
    prof, models, (s0, s1), slab = _models(kind)
    iv = continua(prof, s0, s1, slab)
    model = ork.make_model(kind)
    rng = np.random.default_rng(7)
    worst = 0.0
    n = 0
    lo, hi = (0.5, 4.99) if kind == "cylinder_density" else (0.41, 2.99)
    while n < 24:
        k = rng.uniform(0.05, 4.5)
        W = rng.uniform(lo, hi)
        if not regular_mask(W, iv, 0.03):
            continue
        mode = int(rng.integers(0, len(models)))
        e, i = rp.dispersion(models[mode], k, W * k, **TIGHT)
        e2, i2 = ork.point(model, mode, k, W * k)
        worst = max(worst, abs(e2 - e) / abs(e), abs(i2 - i) / abs(i))
        n += 1
    assert worst < 2e-9, worst     # the scipy path at rtol 1e-12 is itself good to ~1e-10


def test_fsolve_and_linear_shooting_agree():
    prof = cyl_profile()
    m = rp.CylinderDensity(prof, 1)
    for k, W in ((1.0, 3.3), (2.5, 1.6), (0.5, 0.7)):
        a = rp.dispersion(m, k, W * k, rtol=1e-11, atol=1e-30, xtol=1e-14)
        b = rp.dispersion(m, k, W * k, rtol=1e-11, atol=1e-30, shoot="linear")
        assert abs(a[1] - b[1]) < 1e-8 * abs(b[1])
    s = rp.SlabDensity(slab_profile(), "sausage")
    a = rp.dispersion(s, 0.75, 0.7 * 0.75, rtol=1e-11, atol=1e-30, xtol=1e-14)
    b = rp.dispersion(s, 0.75, 0.7 * 0.75, rtol=1e-11, atol=1e-30, shoot="linear")
    assert abs(a[1] - b[1]) < 1e-8 * abs(b[1])


def test_skip_rule_and_brackets():
    prof = cyl_profile()
    m = rp.CylinderDensity(prof, 1)
    e, i = rp.dispersion(m, 1.0, 5.2)      # W > vA_e: m_e < 0 -> skipped (Density_cylinder.py:760)
    assert np.isnan(e) and np.isnan(i)
    d = np.array([1.0, -1.0, np.nan, -2.0, 3.0, 3.0, 0.0, -1.0])
    assert list(rp.brackets(d)) == [0, 3]
    ki, wi = ork.brackets(d[None, :])
    assert list(wi) == [0, 3]
