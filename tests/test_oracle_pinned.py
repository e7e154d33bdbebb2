"""Pin the oracle (CPU, no GPU):

  reference itself (executed in the build container, fixtures in tests/golden/)
      == oracle/reference_path.py at the reference's own solver settings
  reference's shipped root tables (Example data/*.pickle)
      satisfy the reference's acceptance test under the oracle
  oracle/reference_path.py at tight tolerances (the converged value of the reference's
  formulation)  == oracle/dispersion_rk.c (the fast C restatement used on whole grids)

for every solver variant in helpers.CASES (cylinder/slab density, coronal and
photospheric parameter sets, the sheared-flow slab, the axial-flow and the rotational-flow cylinder).
"""
import os
import warnings

import numpy as np
import pytest

from helpers import CASES, ROOT_CASES, regular_mask
from oracle import reference_path as rp
from oracle import rk_oracle as ork

warnings.filterwarnings("ignore")
TIGHT = dict(rtol=1e-12, atol="scaled", shoot="linear")


@pytest.mark.parametrize("name,tol,stride", [("cylinder_density", 1e-7, 3), ("cylinder_epstein", 1e-7, 3),
                                             ("cylinder_density_w15", 1e-7, 3), ("slab_density_w3", 1e-8, 2),
                                             ("slab_density", 1e-8, 4),
                                             ("cylinder_photospheric", 1e-7, 1),
                                             ("slab_photospheric", 1e-7, 3), ("slab_flow", 1e-7, 2),
                                             ("cylinder_flow", 1e-7, 3), ("slab_flow_photospheric", 1e-7, 2)])
def test_oracle_equals_executed_reference(golden_dir, name, tol, stride):
    """D from the reference's own sausage()/kink() vs the restatement at the SAME solver
    settings (scipy defaults, fsolve, the reference's output grids): agreement is at the
    1e-10 level, i.e. the restatement evaluates the same functions in the same order."""
    case = CASES[name]
    g = np.load(os.path.join(golden_dir, "ref_D_%s.npz" % case.fixture))
    iv = case.intervals()
    models = {m: case.scipy_model(m, fast=False) for m in (0, 1)}
    if name == "slab_photospheric":
        for m in models.values():
            m.n_int_out = 10**5           # that script: ix = linspace(-1, 1, 1e5)
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
        # where the reference's fsolve stops short of convergence ("not making good progress")
        # its D is termination noise; the linear-shooting value exposes those points
        el, il = rp.dispersion(models[int(mode)], k, w, shoot="linear")
        if abs((el - il) - Dor) > 1e-4 * max(abs(el), abs(il)):
            continue
        if name == "slab_flow_photospheric" and abs((el - il) - Dref) > 1e-4 * max(abs(el), abs(il)):
            # that script starts fsolve at 0.5 while the 7-wavelength exterior makes the slope ~1e6-1e9:
            # at some points the reference's own fsolve gives up where the restatement's converges
            continue
        assert np.sign(Dor) == np.sign(Dref)
        assert abs(Dor - Dref) <= tol * max(abs(el), abs(il)), (mode, k, w, Dref, Dor)
        n_checked += 1
    assert n_checked >= 6 and n_skipped >= 2


def test_per_point_sympy_variant_is_the_same_function(golden_dir):
    """CylinderDensityPerPoint rebuilds D, C1, C2, C3, F, dF, g with sympy and lambdifies six functions at every
    (k, omega), as Density_cylinder.py:705-757 does (the cost structure bench.py's `unhoisted` CPU rate
    times): same values as the executed reference and as the hoisted closed forms."""
    case = CASES["cylinder_density"]
    g = np.load(os.path.join(golden_dir, "ref_D_%s.npz" % case.fixture))
    prof = rp.GaussianDensity(rp.CYL_CORONAL, width=0.95, const_B=True)
    n = 0
    for mode in (0, 1):
        hoisted, per_point = rp.CylinderDensity(prof, mode), rp.CylinderDensityPerPoint(prof, mode)
        for j in np.nonzero(g["mode"] == mode)[0][4::13]:
            k, w, Dref = g["k"][j], g["w"][j], g["D"][j]
            Dp = rp.D(per_point, k, w)
            if np.isnan(Dref):
                assert np.isnan(Dp)
                continue
            e, i = rp.dispersion(hoisted, k, w)
            assert abs(Dp - (e - i)) <= 1e-7 * max(abs(e), abs(i)), (mode, k, w)
            assert abs(Dp - Dref) <= 1e-7 * max(abs(e), abs(i)), (mode, k, w)
            n += 1
    assert n >= 8


@pytest.mark.parametrize("name", ["cylinder_density", "slab_density", "cylinder_flow", "slab_flow"])
def test_reference_scan_and_bisection(golden_dir, name):
    """The reference's own scan+bisection output (sol_ks/sol_omegas) over a few intervals:
    every mode it reports is a root of the oracle's D to within its acceptance band (1 %; 6 % for
    the axial-flow cylinder script)."""
    case = CASES[name]
    tol = 6.0 if name == "cylinder_flow" else 1.0          # the script's xi_tol / p_tol
    g = np.load(os.path.join(golden_dir, "ref_scan_%s.npz" % case.fixture))
    n = found = 0
    while "scan%d_k" % n in g.files:
        mode = int(g["scan%d_mode" % n][0]); k = float(g["scan%d_k" % n][0])
        freq = g["scan%d_freq" % n]; ws = g["scan%d_sol_ws" % n]
        model = case.scipy_model(mode)
        mine = rp.find_roots(model, k, freq, **TIGHT)
        for w in ws:
            # at the reference's own solver settings the restatement reproduces its decision to accept ...
            e, i = rp.dispersion(case.scipy_model(mode, fast=False), k, float(w))
            assert rp.mismatch_percent(e, i) < tol
            # ... and the accepted point sits next to a converged root (the reference stops anywhere inside
            # its band, and its D carries the solver noise described in DESIGN.md: up to 0.5 % in omega for
            # the steep sausage branch of the flow slab)
            assert np.min(np.abs(mine - w)) < 1e-2 * tol * abs(w)
            found += 1
        n += 1
    assert n >= 2 and found >= 1


def test_shipped_root_tables_pass_acceptance_under_oracle(golden_dir):
    """Example data/*.pickle = roots the reference accepted (<1 % mismatch).  With the same
    profile width the oracle must agree: nearly all of them are inside the 1 % band."""
    g = np.load(os.path.join(golden_dir, "ref_roots.npz"))
    stats = {}
    for name, case in ROOT_CASES.items():
        fam = case.family
        tags = sorted(set(f[len(fam) + 1:].split("_")[0] for f in g.files if f.startswith(fam + "_")))
        total = inside = 0
        medians = []
        for tag in tags:
            width = float(g["%s_%s_width" % (fam, tag)][0])
            iv = case.intervals(width)
            model = case.c_model(width)
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
                pct = pct[np.isfinite(pct)]
                total += len(pct)
                inside += int((pct < 1.5 * case.tol_percent).sum())
                if len(pct) > 5:
                    medians.append(np.median(pct) / case.tol_percent)
        stats[name] = (inside, total, max(medians) if medians else None)
    for name, (inside, total, med) in stats.items():
        assert total > 50, (name, total)
        assert inside / total > 0.85, (name, inside, total)
        assert med < 0.9, (name, med)      # accepted anywhere below tol (1 %; 6 % cylinder flow) -> median near tol/2


@pytest.mark.parametrize("name", list(CASES))
def test_c_oracle_equals_converged_scipy_path(name):
    case = CASES[name]
    model = case.c_model()
    rng = np.random.default_rng(7)
    worst = 0.0
    n = tries = 0
    lo, hi = case.W
    while n < 16 and tries < 4000:
        tries += 1
        k = rng.uniform(0.05, 4.5)
        W = rng.uniform(lo, hi)
        mode = int(rng.integers(0, len(case.modes)))
        if not case.regular(k, np.array([W]), mode, margin=0.03)[0, 0]:
            continue
        e, i = rp.dispersion(case.scipy_model(mode), k, W * k, **TIGHT)
        if np.isnan(e):
            continue
        e2, i2 = ork.point(model, mode, k, W * k)
        worst = max(worst, abs(e2 - e) / abs(e), abs(i2 - i) / abs(i))
        n += 1
    assert n == 16
    assert worst < 5e-9, worst     # the scipy path at rtol 1e-12 is itself good to ~1e-9 in amplitude


@pytest.mark.parametrize("fixture,mode,v_twist,power", [
    ("cylinder_rotation_sausage", 0, 0.15, 1.25), ("cylinder_rotation_kink", 1, 0.15, 1.25),
    ("cylinder_rotation_sausage_p1", 0, 0.1, 1.0), ("cylinder_rotation_kink_p1", 1, 0.1, 1.0),
    ("cylinder_rotation_kink_p08", 1, 0.25, 0.8), ("cylinder_rotation_kink_slow_p08", 1, 0.1, 0.8)])
def test_rotation_oracle_equals_executed_reference(golden_dir, fixture, mode, v_twist, power):
    """Rotational-flow cylinder: D from the reference's own sausage()/kink() (scipy defaults, sympy
    coefficients rebuilt per point) vs the restatement (sympy coefficients built once), for the
    rotation laws v_phi = 0.15 r^1.25, v_phi = 0.1 r and - the two kink scripts exactly as shipped -
    v_phi = 0.25 r^0.8 (..._kink_fast.py:176-177) and 0.1 r^0.8 (..._kink_slow.py:176-177)."""
    import helpers
    md = rp.CYL_PHOTOSPHERIC
    s_end = 0.01 if mode == 0 else 0.001
    g = np.load(os.path.join(golden_dir, "ref_D_%s.npz" % fixture))
    model = rp.CylinderRotation(md, mode, v_twist, power, s_end=s_end)
    c_model = ork.make_model("cylinder_rotation", medium=md, v_twist=v_twist, power=power, s_end=s_end)

    def regular(k, W, margin=0.03):
        ok = helpers.regular_mask(np.array([W]), helpers.rotation_continua(md, v_twist, power, s_end, mode, k),
                                  margin)[0]
        for dW in (-margin, 0.0, margin):
            ok = ok and helpers.rotation_regular(md, v_twist, power, s_end, mode, k, np.array([W + dW]))[0]
        return ok

    # power < 1, layer down to r = 0.001: LSODA at its default 1.5e-8 tolerance is amplified by the
    # 1/r^2 growth of the coefficients towards the axis, so two evaluations of the SAME formulation at the
    # reference's settings (fsolve vs two linear shots) already differ by 1e-3..5e-2 of the scale, and the
    # reference's value sits 2-14 % off the converged one (same sign) - its usual amplitude error
    # (DESIGN.md).  The port reproduces the executed reference to within that same noise (1e-10 at most
    # points, up to 6e-2 where fsolve's termination decides), and the C oracle the
    # converged value (< 1e-10).
    steep = power < 1.0
    skip_tol, agree_tol, stride = (0.1, 0.1, 1) if steep else (1e-4, 2e-3, 2)
    n_checked = n_skipped = 0
    for k, w, Dref in list(zip(g["k"], g["w"], g["D"]))[::stride]:
        Dor = rp.D(model, k, w)
        if np.isnan(Dref):
            assert np.isnan(Dor)
            n_skipped += 1
            continue
        if not regular(k, w / k):
            continue
        el, il = rp.dispersion(model, k, w, shoot="linear")
        if abs((el - il) - Dor) > skip_tol * max(abs(el), abs(il)):
            continue                      # reference fsolve stopped short of convergence
        # coefficients evaluated in a different floating-point order (sympy cse once vs per point)
        # steer LSODA through different step sequences: agreement at its 1e-8 tolerance amplified
        # by the 1/r^2 growth towards the axis, not at rounding level
        assert abs(Dor - Dref) <= agree_tol * max(abs(el), abs(il)), (k, w, Dref, Dor)
        # and the C restatement (first-order system, no coefficient derivatives) agrees with the converged path
        e2, i2 = ork.point(c_model, mode, k, w)
        et, it_ = rp.dispersion(model, k, w, **TIGHT)
        assert abs((e2 - i2) - (et - it_)) <= 1e-7 * max(abs(et), abs(it_)), (k, w)
        # the reference's own value: the converged one with its amplitude error, never another sign
        # (where D is not a small difference of ext and int, i.e. away from a root)
        if abs(et - it_) > 1e-2 * max(abs(et), abs(it_)):
            assert 0.5 < Dref / (et - it_) < 1.6, (k, w, Dref, et - it_)
        n_checked += 1
    assert n_checked >= (12 if steep else 6) and n_skipped >= 2


ROT_AMPLITUDES = {"005": 0.05, "01": 0.1, "015": 0.15, "025": 0.25}
ROT_POWERS = {"08": 0.8, "09": 0.9, "1": 1.0, "125": 1.25}


def rotation_tables(g):
    """(key, v_twist, power, mode, s_end) of every shipped rotational root table in ref_roots.npz"""
    out = []
    for f in sorted(g.files):
        if not (f.startswith("rot_v") and f.endswith("_k")):
            continue
        key = f[:-2]
        vt, pw, kind = key[len("rot_v"):].split("_", 2)
        mode = 0 if "sausage" in kind else 1
        out.append((key, ROT_AMPLITUDES[vt], ROT_POWERS[pw[1:]], mode, 0.01 if mode == 0 else 0.001))
    return out


def test_rotation_shipped_root_tables(golden_dir):
    """Example data of the rotational-flow solvers: ALL 54 shipped tables - four rotation amplitudes,
    powers 0.8, 0.9, 1, 1.25, the sausage and the kink scripts (22 + 32 tables).  The scripts accept at
    1.5-4.5 % (their xi_tol / P_tol) and the kink files were produced with end points and tolerances that
    changed between runs, so the band checked here is 5 %.  Every table, power 0.8 and kink included, has
    >= 85 % of its regular points inside the band under the C oracle, and >= 65 % of the sampled points of every
    table are regular (no resonance inside the layer): the shipped rotational results lie where parity
    with the reference is defined."""
    import helpers
    g = np.load(os.path.join(golden_dir, "ref_roots.npz"))
    md = rp.CYL_PHOTOSPHERIC
    tables = rotation_tables(g)
    assert len(tables) == 54 and sum(t[3] for t in tables) == 32
    total = inside = 0
    for key, vt, pw, mode, s_end in tables:
        k, w = g[key + "_k"][::3], g[key + "_w"][::3]
        model = ork.make_model("cylinder_rotation", medium=md, v_twist=vt, power=pw, s_end=s_end)
        pct = np.array([rp.mismatch_percent(*ork.point(model, mode, a, b)) for a, b in zip(k, w)])
        reg = np.array([helpers.rotation_regular(md, vt, pw, s_end, mode, a, np.array([b / a]))[0]
                        for a, b in zip(k, w)])
        ok = np.isfinite(pct) & reg
        assert ok.sum() >= 0.65 * len(k), (key, int(ok.sum()), len(k))
        assert (pct[ok] < 5.0).mean() > 0.85, (key, float((pct[ok] < 5.0).mean()))
        total += int(ok.sum())
        inside += int((pct[ok] < 5.0).sum())
    assert total > 2000 and inside / total > 0.95, (inside, total)


def test_fsolve_and_linear_shooting_agree():
    m = CASES["cylinder_density"].scipy_model(1)
    for k, W in ((1.0, 3.3), (2.5, 1.6), (0.5, 0.7)):
        a = rp.dispersion(m, k, W * k, rtol=1e-11, atol="scaled", xtol=1e-14)
        b = rp.dispersion(m, k, W * k, rtol=1e-11, atol="scaled", shoot="linear")
        assert abs(a[1] - b[1]) < 1e-8 * abs(b[1])
    for name, k, W in (("slab_density", 0.75, 0.7), ("slab_flow", 1.5, 1.6)):
        s = CASES[name].scipy_model(0)
        a = rp.dispersion(s, k, W * k, rtol=1e-11, atol="scaled", xtol=1e-14)
        b = rp.dispersion(s, k, W * k, rtol=1e-11, atol="scaled", shoot="linear")
        assert abs(a[1] - b[1]) < 1e-8 * abs(b[1])


def test_skip_rule_and_brackets():
    m = CASES["cylinder_density"].scipy_model(1)
    e, i = rp.dispersion(m, 1.0, 5.2)      # W > vA_e: m_e < 0 -> skipped (Density_cylinder.py:760)
    assert np.isnan(e) and np.isnan(i)
    d = np.array([1.0, -1.0, np.nan, -2.0, 3.0, 3.0, 0.0, -1.0])
    assert list(rp.brackets(d)) == [0, 3]
    ki, wi = ork.brackets(d[None, :])
    assert list(wi) == [0, 3]
