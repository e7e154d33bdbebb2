"""GPU parity tests (run with -m gpu on the B200 box).  Everything goes through the C ABI
(ctypes -> libeigensolver_b200.so); the oracles are only the checkers.

Tolerances (north_star): bracket indices identical wherever D is above the noise floor
(= outside the resonant continua, see helpers.continua), refined roots within 1e-9
relative, D itself within 1e-9 of max(|ext|,|int|).
"""
import os
import warnings

import numpy as np
import pytest

import eigensolver_b200 as esb
from helpers import CASES, ROOT_CASES, regular_mask
from oracle import reference_path as rp
from oracle import rk_oracle as ork

pytestmark = pytest.mark.gpu
warnings.filterwarnings("ignore")

D_TOL = 1e-9       # |D_gpu - D_oracle| / max(|ext|, |int|)
ROOT_TOL = 1e-9    # relative, refined roots (north_star)
TIGHT = dict(rtol=1e-12, atol="scaled", shoot="linear")

@pytest.fixture(scope="module")
def solvers():
    s = {name: c.gpu_solver(esb) for name, c in CASES.items()}
    yield s
    for v in s.values():
        v.close()


def _grid_case(name, nk=20, nw=240):
    c = CASES[name]
    k = np.linspace(0.05, 4.5, nk)
    W = np.linspace(c.W[0], c.W[1], nw)
    return k, W


@pytest.mark.parametrize("name", list(CASES))
def test_grid_brackets_and_roots_match_c_oracle(solvers, name):
    s = solvers[name]
    case = CASES[name]
    model = case.c_model()
    k, W = _grid_case(name)
    for mode in case.modes:
        reg = case.regular(k, W, mode)
        ext, inq = s.dispersion_grid(mode, k, W)
        e0, i0 = ork.grid(model, mode, k, W)
        # the skip rule (m_e < 0 -> not evaluated) is identical everywhere; overflow of the
        # exterior growth (only the 7-wavelength photospheric slab near cT_e) is "no value" too
        assert np.array_equal(np.isnan(ext), ~(np.isfinite(e0) & np.isfinite(i0)))
        assert np.array_equal(np.isnan(inq), np.isnan(ext))
        ok = reg & np.isfinite(e0) & np.isfinite(i0)
        e0 = np.where(np.isfinite(e0) & np.isfinite(i0), e0, np.nan)
        i0 = np.where(np.isfinite(e0), i0, np.nan)
        assert ok.sum() > case.min_regular * ok.size
        dev = np.abs((ext - inq) - (e0 - i0)) / np.maximum(np.abs(e0), np.abs(i0))
        assert np.nanmax(dev[ok]) < D_TOL, (mode, np.nanmax(dev[ok]))
        assert np.nanmax(np.abs(ext - e0)[ok] / np.abs(e0)[ok]) < 1e-10     # closed-form exterior
        # brackets: identical index sets where both end points are above the noise floor
        tab = s.find_roots(mode, k, W)
        gk, gw = ork.brackets(ext - inq)
        assert np.array_equal(gk, tab.k_index) and np.array_equal(gw, tab.w_index)   # device ballot == numpy
        ok_iv = reg[:, :-1] & reg[:, 1:]
        ok_, ow_ = ork.brackets(e0 - i0)
        sel_o = ok_iv[ok_, ow_]
        sel_g = ok_iv[tab.k_index, tab.w_index]
        assert np.array_equal(ok_[sel_o], tab.k_index[sel_g])
        assert np.array_equal(ow_[sel_o], tab.w_index[sel_g])
        assert sel_g.sum() >= 5
        # refined roots (accepted modes) vs the oracle's own refinement of the same bracket
        idx = np.nonzero(sel_g & (tab.accepted == 1))[0]
        assert len(idx) >= 3
        for j in idx:                       # EVERY accepted root above the noise floor
            kk = k[tab.k_index[j]]
            r, er, ir = ork.refine(model, mode, kk, kk * W[tab.w_index[j]], kk * W[tab.w_index[j] + 1])
            assert abs(tab.omega[j] - r) <= ROOT_TOL * abs(r), (mode, kk, r, tab.omega[j])
            assert rp.mismatch_percent(er, ir) < 1.0
        # poles (sign change through infinity) are found but never accepted; a bracket that
        # straddles a sliver of skipped points (m_e < 0 between two evaluated ones) ends in NaN
        for j in np.nonzero(sel_g & (tab.accepted == 0))[0][:20]:
            assert not rp.mismatch_percent(tab.ext[j], tab.intq[j]) < 1.0


@pytest.mark.parametrize("name", list(CASES))
def test_roots_match_converged_scipy_reference_path(solvers, name):
    """The north_star statement itself: roots within 1e-9 relative of the reference's
    numpy/scipy path (odeint + shooting, run to convergence) on identical inputs."""
    s = solvers[name]
    case = CASES[name]
    k = np.array([0.6, 1.0, 2.2, 3.7])
    W = np.linspace(case.roots_window[0], case.roots_window[1], 60)
    n = 0
    for mode in (0, 1):
        tab = s.find_roots(mode, k, W)
        m = case.scipy_model(mode)
        reg = case.regular(k, W, mode, margin=0.03)
        good = (tab.accepted == 1) & reg[tab.k_index, tab.w_index] & reg[tab.k_index, tab.w_index + 1]
        for j in np.nonzero(good)[0][:4]:
            kk = k[tab.k_index[j]]
            w, _, _ = rp.refine(m, kk, kk * W[tab.w_index[j]], kk * W[tab.w_index[j] + 1], **TIGHT)
            assert abs(tab.omega[j] - w) <= ROOT_TOL * abs(w), (mode, kk, w, tab.omega[j])
            n += 1
    assert n >= 3


@pytest.mark.parametrize("name", list(CASES))
def test_against_executed_reference_fixture(solvers, golden_dir, name):
    """Golden D values from the reference's own functions (scipy default tolerances).  The
    reference's exterior integration starts below its absolute tolerance (|y0| = 1e-8 <
    atol = 1.5e-8), so its D carries a common amplitude error of 10-25 %; sign, skip pattern
    and the ext/int ratio are what it determines, and those must agree."""
    case = CASES[name]
    iv = None if case.kind == "cylinder_rotation" else case.intervals()
    s = solvers[name]
    n = 0
    for mode in (0, 1):
        if case.kind == "cylinder_rotation":
            # one fixture per script: the sausage script stops at r = 0.01, the kink script at 0.001
            g = np.load(os.path.join(golden_dir, "ref_D_cylinder_rotation_%s%s.npz" % (("sausage", "kink")[mode],
                                                                                  case.fixture or "")))
            k, w, Dref = g["k"], g["w"], g["D"]
            rot = esb.DispersionSolver("cylinder_rotation", profile=esb.PowerLawRotation(case.v_twist, case.power),
                                       s_end=0.01 if mode == 0 else 0.001)
            ext, inq = rot.dispersion_grid(mode, k, w[:, None], layout="per_k")
            rot.close()
            D = (ext - inq)[:, 0]
            assert np.array_equal(np.isnan(D), np.isnan(Dref))
            scale = np.maximum(np.abs(ext), np.abs(inq))[:, 0]
            reg = np.array([case.regular(kk, np.array([ww / kk]), mode, margin=0.03)[0, 0] for kk, ww in zip(k, w)])
            ok = ~np.isnan(Dref) & reg & (np.abs(D) > 1e-2 * scale)
            assert ok.sum() >= 10, ok.sum()
            assert np.array_equal(np.sign(D[ok]), np.sign(Dref[ok]))
            ratio = D[ok] / Dref[ok]
            assert ratio.min() > 0.5 and ratio.max() < 1.6, (ratio.min(), ratio.max())
            n += 2 * ok.sum()
            continue
        g = np.load(os.path.join(golden_dir, "ref_D_%s.npz" % case.fixture))
        sel = g["mode"] == mode
        k, w, Dref = g["k"][sel], g["w"][sel], g["D"][sel]
        ext, inq = s.dispersion_grid(mode, k, w[:, None], layout="per_k")
        D = (ext - inq)[:, 0]
        assert np.array_equal(np.isnan(D), np.isnan(Dref))
        scale = np.maximum(np.abs(ext), np.abs(inq))[:, 0]
        # exclude: continua; points where D is a small difference of ext and int (the reference's
        # 1e-7 solver noise decides the sign there); the slab's W -> vA_e corner where the exterior
        # solution never leaves the reference's absolute-tolerance noise (see test_oracle_pinned)
        ok = ~np.isnan(Dref) & regular_mask(w / k, iv) & (np.abs(D) > 1e-3 * scale)
        if name.startswith("slab_density"):
            ok &= w / k < 2.9
        if name == "slab_flow_photospheric":
            # that script starts fsolve at 0.5 while its 7-wavelength exterior makes the slope 1e6-1e9:
            # where the reference's own fsolve gave up (its D differs from the value linear shooting
            # gives at the same solver settings) the fixture value is termination noise
            m = case.scipy_model(mode)
            for j in np.nonzero(ok)[0]:
                el, il = rp.dispersion(m, k[j], w[j], shoot="linear")
                ok[j] = abs((el - il) - Dref[j]) <= 1e-4 * max(abs(el), abs(il))
        assert ok.sum() >= 12, ok.sum()
        assert np.array_equal(np.sign(D[ok]), np.sign(Dref[ok]))
        ratio = D[ok] / Dref[ok]
        if name not in ("slab_photospheric", "slab_flow_photospheric"):
            # (that script's 7-wavelength exterior amplifies the reference's start-up error to
            #  factors of 4-100 for small W: only the sign is meaningful there)
            assert ratio.min() > 0.5 and ratio.max() < 1.6, (ratio.min(), ratio.max())
        n += ok.sum()
    assert n >= 30


def test_shipped_root_tables(golden_dir):
    """The reference's Example data root tables pass its own 1 % acceptance test on the GPU."""
    g = np.load(os.path.join(golden_dir, "ref_roots.npz"))
    for name, case in ROOT_CASES.items():
        fam = case.family
        tags = sorted(set(f[len(fam) + 1:].split("_")[0] for f in g.files if f.startswith(fam + "_")))
        total = inside = 0
        for tag in tags:
            width = float(g["%s_%s_width" % (fam, tag)][0])
            iv = case.intervals(width)
            with case.gpu_solver(esb, width) as s:
                for mi, mode in ((0, "sausage"), (1, "kink")):
                    k = g["%s_%s_%s_k" % (fam, tag, mode)]
                    w = g["%s_%s_%s_w" % (fam, tag, mode)]
                    if not len(k):
                        continue
                    reg = regular_mask(w / k, iv, 0.0) & (np.abs(w / k) < 6)
                    k, w = k[reg], w[reg]
                    if not len(k):
                        continue
                    e, i = s.dispersion_grid(mi, k, w[:, None], layout="per_k")
                    pct = np.abs(e - i)[:, 0] * 100 / np.maximum(np.abs(e), np.abs(i))[:, 0]
                    pct = pct[np.isfinite(pct)]
                    total += len(pct)
                    inside += int((pct < 1.5 * case.tol_percent).sum())
        assert total > 300 and inside / total > 0.85, (name, inside, total)


#: the assignment-line edits of tests/golden/make_scan_golden.py, as ReferenceScript keywords
DROPIN_OVERRIDES = {
    "cylinder_density": {}, "cylinder_density_photospheric": {}, "slab_density": {},
    "slab_density_photospheric": dict(width=0.9), "slab_flow": dict(width=1.0), "slab_flow_photospheric": {},
    "cylinder_flow": dict(medium=esb.AxialFlowMedium(1.0, 2.0, 5.0, 0.5, U_i0=0.35), width=1.0),
    "rotation_sausage": {}, "rotation_kink": dict(profile=esb.PowerLawRotation(0.15, 1.25)),
    # the two kink scripts exactly as shipped (0.25 r^0.8 / 0.1 r^0.8): fixture names carry a variant tag
    "rotation_kink@p08": {}, "rotation_kink_slow@p08": {},
}


class _Q:
    def __init__(self):
        self.items = []

    def put(self, x):
        self.items.append(x)


def test_reference_rule_reproduces_the_scripts_scan(golden_dir):
    """sausage()/kink() keep the reference's signature and, under the "reference" accept rule, reproduce
    the reference's OWN scan + bisection output: tests/golden/ref_scans.npz holds sol_omegas of 48 + 6 calls
    of the unmodified scripts (9 scripts, both modes, windows with and without modes; most windows that
    contain a mode return nothing - the scripts' recursion follows the upper half only).  A solution of
    the scripts is a point of the dyadic refinement of the frequency grid, so agreement is to ROUNDING.

    Measured: 37 of 48 calls identical (same count, same points to 1e-9).  The other 11, by cause:
      * the scripts' bisection keeps module-global state (`loop_ws`, `xi_diff_loop_check`) from one bracket
        of a call to the next, which sometimes lets a later bracket recurse into its lower half
        (cylinder k = 3.5, photospheric cylinder k = 1.5, flow slab): not emulated;
      * flow_multiprocessor.py asks for a mismatch below 1e-6 %, which its own odeint noise cannot reach:
        it returns nothing where the converged D does reach it;
      * the rotational kink script evaluates the second-order form, noisy where C3 ~ 0: a few
        grid points are inside the 2.5 % band for the converged D and not for the script.
    For those calls both directions are checked in the weaker sense: every solution of the script lies
    within one grid interval of a root of the converged rule, and every solution reported here passes
    the script's own acceptance test by construction."""
    g = np.load(os.path.join(golden_dir, "ref_scans.npz"))
    n_cases = n_sols = exact = n_steep = 0
    scripts = {}
    try:
        n = 0
        while "c%d_script" % n in g.files:
            name = str(g["c%d_script" % n])
            if name not in scripts:
                scripts[name] = esb.ReferenceScript(name.split("@")[0], **DROPIN_OVERRIDES[name])
            script = scripts[name]
            mode = int(g["c%d_mode" % n][0]); k = float(g["c%d_k" % n][0])
            freq = g["c%d_freq" % n]; ws_ref = np.sort(g["c%d_sol_ws" % n])
            ws, ks = _Q(), _Q()
            (script.kink if mode == 1 else script.sausage)(k, ws, ks, freq)
            assert len(ws.items) == 1 and len(ks.items) == 1 and len(ws.items[0]) == len(ks.items[0])
            assert all(kk == k for kk in ks.items[0])
            got = np.sort(np.array(ws.items[0], dtype=np.float64))
            steep = "@" in name
            # "@p08": the two kink scripts exactly as shipped (0.25 r^0.8 / 0.1 r^0.8).  Their own D is 2-14 %
            # off in amplitude there (tests/test_oracle_pinned.py), which decides the band test at marginal
            # grid points: these six calls are held to the two-direction check below only - and the fast
            # script's solution must be among the GPU's points exactly
            n_steep += steep
            n_cases += not steep
            n_sols += len(ws_ref)
            if name == "rotation_kink@p08":
                assert all(np.min(np.abs(got - w)) <= 1e-9 * abs(w) for w in ws_ref), (name, k, got, ws_ref)
            if not steep and len(got) == len(ws_ref) and np.allclose(got, ws_ref, rtol=1e-9, atol=0):
                exact += 1
            elif not (len(got) == len(ws_ref) and np.allclose(got, ws_ref, rtol=1e-9, atol=0)):
                script.solver.set_accept_rule("converged")
                conv = script.solver.find_roots(mode, [k], freq, layout="shared", tol_percent=script.tol)
                script.solver.set_accept_rule(script.rule)
                dw = abs(freq[1] - freq[0])
                for w in ws_ref:
                    assert len(conv.omega) and np.min(np.abs(conv.omega - w)) <= dw, (name, mode, k, w)
                if len(got):
                    e, q = script.solver.dispersion_grid(mode, [k], got, layout="shared")
                    assert np.all(np.abs(e - q) * 100 / np.maximum(np.abs(e), np.abs(q)) < script.tol)
            n += 1
    finally:
        for sc in scripts.values():
            sc.close()
    assert n_cases >= 40 and n_sols >= 12 and n_steep == 6
    assert exact >= 0.75 * n_cases, (exact, n_cases)


def test_reference_api_driver_loop():
    with esb.ReferenceScript("cylinder_density", rule="converged") as script:
        out = script.run(np.linspace(0.5, 4.0, 8), speeds=[2.95, 4.0, 4.95], n_freq=40)
    assert len(out) == 4 and len(out[0]) == len(out[1]) and len(out[2]) == len(out[3])
    assert len(out[2]) >= 4 and np.all(out[2] / out[3] > 2.9) and np.all(out[2] / out[3] < 5.0)
    # the reference rule returns a subset of the branches' points (it gives up on lower-half roots) plus
    # the grid points inside the band; all of them pass the acceptance test
    with esb.ReferenceScript("cylinder_density") as script:
        ref = script.run(np.linspace(0.5, 4.0, 8), speeds=[2.95, 4.0, 4.95], n_freq=40)
        assert len(ref) == 4 and len(ref[2]) == len(ref[3])
        if len(ref[2]):
            e, q = script.solver.dispersion_grid(1, ref[3], np.asarray(ref[2])[:, None], layout="per_k")
            assert np.all(np.abs(e - q) * 100 / np.maximum(np.abs(e), np.abs(q)) < 1.0)


def test_scan_models_equals_single_sweeps():
    """esb_scan_models (configs[4]: every equilibrium enqueued back to back, no host synchronisation,
    one compact table) returns what one sweep per equilibrium returns, and its values match the C
    oracle of each equilibrium."""
    from eigensolver_b200.scan import density_flow_grid
    dens, flow = density_flow_grid([0.15, 0.2055, 0.3], [0.2, 0.5, 0.9])
    k = np.linspace(0.4, 4.0, 12)
    for kind, pts, modes, W in (("cylinder_density", dens, [0, 1, 2], np.linspace(0.55, 4.5, 300)),
                                ("slab_flow", flow, [0, 1], np.linspace(1.25, 2.45, 200))):
        with esb.DispersionSolver(kind) as s:
            s.upload_axes(k, W)
            s.set_schedule("lane")
            tab, nb = s.scan_models(pts, modes)
            tab = {name: np.array(v) for name, v in tab.items()}
            assert nb.shape == (len(pts), len(modes)) and nb.sum() == len(tab["omega"]) > 0
            # ordered by (model, slot, k index, omega index)
            key = ((tab["model"].astype(np.int64) * 8 + tab["slot"]) * len(k) + tab["k_index"]) * len(W) + tab["w_index"]
            assert np.all(np.diff(key) > 0)
            for i, p in enumerate(pts):
                s.reconfigure(medium=p["medium"], profile=p["profile"])
                singles = s.find_roots_multi(modes, k, W)
                for slot, t in enumerate(singles):
                    sel = (tab["model"] == i) & (tab["slot"] == slot)
                    assert sel.sum() == nb[i, slot] == len(t.omega)
                    assert np.array_equal(tab["k_index"][sel], t.k_index) and np.array_equal(tab["w_index"][sel], t.w_index)
                    assert np.array_equal(tab["omega"][sel], t.omega, equal_nan=True)
                    assert np.array_equal(tab["accepted"][sel], t.accepted)
            # a table that outgrows its room is reported, not silently truncated
            with pytest.raises(esb.EsbError, match="capacity"):
                s.scan_models(pts, modes, capacity_per_table=2)
    # against the oracle: one accepted mode of every density equilibrium
    with esb.DispersionSolver("cylinder_density") as s:
        Wr = np.linspace(3.3, 4.95, 120)          # above the Alfven continuum of every contrast (edge <= 3.05)
        s.upload_axes(k, Wr)
        tab, nb = s.scan_models(dens, [1])
        for i, p in enumerate(dens):
            model = ork.make_model("cylinder_density", medium=p["medium"], width=p["profile"].width)
            sel = np.nonzero((tab["model"] == i) & (tab["accepted"] == 1))[0]
            assert len(sel) >= 3
            for j in sel[:: max(1, len(sel) // 6)]:
                kk = k[tab["k_index"][j]]
                r, _, _ = ork.refine(model, 1, kk, kk * Wr[tab["w_index"][j]], kk * Wr[tab["w_index"][j] + 1])
                assert abs(tab["omega"][j] - r) <= ROOT_TOL * abs(r), (i, kk, r, tab["omega"][j])
    # ... and accepted modes of every flow equilibrium, both parities, forward branch (the speeds between the
    # Doppler-shifted continua and the exterior cut-off)
    from helpers import flow_continua, regular_mask
    from oracle import reference_path as rp
    with esb.DispersionSolver("slab_flow") as s:
        Wf = np.linspace(1.25, 2.45, 200)
        s.upload_axes(k, Wf)
        tab, nb = s.scan_models(flow, [0, 1])
        for i, p in enumerate(flow):
            rmd = rp.FlowMedium(width=p["profile"].width, U_i0=p["medium"].U_i0)
            model = ork.make_model("slab_flow", medium=rmd, width=p["profile"].width)
            checked = 0
            for slot in (0, 1):
                sel = np.nonzero((tab["model"] == i) & (tab["slot"] == slot) & (tab["accepted"] == 1))[0]
                sel = sel[regular_mask(Wf[tab["w_index"][sel]], flow_continua(rmd), 0.03) &
                          regular_mask(Wf[tab["w_index"][sel] + 1], flow_continua(rmd), 0.03)]
                for j in sel[:: max(1, len(sel) // 6)]:
                    kk = k[tab["k_index"][j]]
                    r, _, _ = ork.refine(model, slot, kk, kk * Wf[tab["w_index"][j]], kk * Wf[tab["w_index"][j] + 1])
                    assert abs(tab["omega"][j] - r) <= ROOT_TOL * abs(r), (i, slot, kk, r, tab["omega"][j])
                    checked += 1
            assert checked >= 3, (i, checked)


def test_edge_cases(solvers):
    s = solvers["cylinder_density"]
    # every point skipped (leaky side, W > vA_e): all NaN, no brackets, no roots
    e, i = s.dispersion_grid(1, [1.0, 2.0], np.linspace(5.1, 6.0, 33))
    assert np.isnan(e).all() and np.isnan(i).all()
    tab = s.find_roots(1, [1.0, 2.0], np.linspace(5.1, 6.0, 33))
    assert len(tab.omega) == 0 and tab.n_brackets == 0
    # smallest grids, ragged sizes (not multiples of the block / warp size)
    for nk, nw in ((1, 2), (1, 33), (3, 129), (5, 31)):
        k = np.linspace(0.7, 2.0, nk); W = np.linspace(3.0, 4.9, nw)
        e, i = s.dispersion_grid(0, k, W)
        e0, i0 = ork.grid(ork.make_model("cylinder_density"), 0, k, W)
        assert np.nanmax(np.abs((e - i) - (e0 - i0)) / np.maximum(abs(e0), abs(i0))) < D_TOL
        tab = s.find_roots(0, k, W)
        bk, bw = ork.brackets(e0 - i0)
        assert np.array_equal(bk, tab.k_index) and np.array_equal(bw, tab.w_index)
    # a single omega: a grid is fine, a root search needs two points
    e, i = s.dispersion_grid(1, [1.0], [3.3])
    assert e.shape == (1, 1) and np.isfinite(e).all()
    with pytest.raises(esb.EsbError):
        s.find_roots(1, [1.0], [3.3])
    # invalid modes / capacity
    with pytest.raises(esb.EsbError):
        s.dispersion_grid(4, [1.0], [3.3])
    with pytest.raises(esb.EsbError):
        solvers["slab_density"].dispersion_grid(2, [1.0], [0.5])
    with pytest.raises(esb.EsbError, match="max_roots"):
        s.find_roots(1, np.linspace(0.5, 4, 16), np.linspace(2.95, 4.95, 64), max_roots=2)
    # very long exterior domain (tiny k) and the resonance W -> cT_e: finite or NaN, never a crash
    e, i = s.dispersion_grid(0, [1e-3, 1e-2], np.linspace(0.4976, 0.52, 16))
    assert e.shape == (2, 16)
    # omega given directly and per-k agree bit for bit with the phase-speed layout
    k = np.linspace(0.5, 3.0, 7); W = np.linspace(3.0, 4.9, 50)
    a = s.dispersion_grid(1, k, W, layout="phase_speed")
    b = s.dispersion_grid(1, k, k[:, None] * W[None, :], layout="per_k")
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    c = s.dispersion_grid(1, k[:1], k[0] * W, layout="shared")
    assert np.array_equal(a[0][:1], c[0])


@pytest.mark.parametrize("kind", list(CASES))
def test_fused_modes_equal_single_mode(solvers, kind):
    """The multi-mode scan (shared coefficients / shared integration) returns what the
    single-mode calls return."""
    s = solvers[kind]
    k, W = _grid_case(kind, nk=9, nw=150)
    modes = list(CASES[kind].modes)
    E, I = s.dispersion_grid_multi(modes, k, W)
    tabs = s.find_roots_multi(modes, k, W)
    for slot, m in enumerate(modes):
        e, i = s.dispersion_grid(m, k, W)
        assert np.array_equal(np.isnan(E[slot]), np.isnan(e))
        ok = ~np.isnan(e) & CASES[kind].regular(k, W, m)   # inside a continuum rounding differences are amplified
        assert np.max(np.abs(E[slot][ok] - e[ok]) / np.abs(e[ok])) < 1e-13
        assert np.max(np.abs(I[slot][ok] - i[ok]) / np.abs(i[ok])) < 1e-11
        t = s.find_roots(m, k, W)
        same = np.array_equal(t.k_index, tabs[slot].k_index) and np.array_equal(t.w_index, tabs[slot].w_index)
        if same:
            acc = (t.accepted == 1) & (tabs[slot].accepted == 1)
            assert np.max(np.abs(t.omega[acc] - tabs[slot].omega[acc]) / np.abs(t.omega[acc])) < 1e-12
        else:   # a sign flip of a value at rounding level inside a continuum may move a noise bracket
            assert abs(len(t.omega) - len(tabs[slot].omega)) <= 0.02 * len(t.omega) + 2
    if kind == "cylinder_density":
        # n = 3 (second fluting order) through the single-mode path
        e3, i3 = s.dispersion_grid(3, k, W)
        e0, i0 = ork.grid(CASES[kind].c_model(), 3, k, W)
        reg = CASES[kind].regular(k, W, 3) & ~np.isnan(e0)
        dev = np.abs((e3 - i3) - (e0 - i0)) / np.maximum(np.abs(e0), np.abs(i0))
        assert np.nanmax(dev[reg]) < D_TOL


def test_pinned_download_equals_pageable(solvers):
    """esb_roots_pinned (one packed copy into page-locked buffers owned by the context) returns the
    table esb_download_roots_slot returns; the views survive further grid calls, not the next download."""
    s = solvers["cylinder_density"]
    k = np.linspace(0.5, 4.0, 17); W = np.linspace(0.55, 4.95, 700)
    s.upload_axes(k, W)
    ns = s.sweep_resident_multi([0, 1, 2])
    for slot, n in enumerate(ns):
        a = s.download_roots(n, slot)
        b = s.download_roots_pinned(slot)
        assert len(b.omega) == n > 0
        for name in ("k_index", "w_index", "k", "omega", "ext", "intq", "accepted", "iterations"):
            assert np.array_equal(getattr(a, name), getattr(b, name), equal_nan=True), name
    # poles recognised from the scan: no evaluation, NaN quantities, position inside the bracket
    t = s.download_roots(ns[1], 1)
    poles = t.iterations == 0
    assert poles.sum() > 0 and np.isnan(t.ext[poles]).all() and (t.accepted[poles] == 0).all()
    lo = k[t.k_index] * W[t.w_index]; hi = k[t.k_index] * W[t.w_index + 1]
    assert np.all((t.omega >= lo) & (t.omega <= hi))
    # an empty slot
    s.upload_axes([1.0, 2.0], np.linspace(5.1, 6.0, 33))
    assert s.sweep_resident_multi([1]) == [0]
    assert len(s.download_roots_pinned(0).omega) == 0


@pytest.mark.parametrize("name", list(CASES))
def test_refine_lane_and_warp_kernels_agree(solvers, name):
    """One lane per point / bracket (throughput) and one warp per point / bracket (latency: sub-interval
    transfer matrices multiplied by a shuffle tree) are two schedules of the same computation."""
    s = solvers[name]
    case = CASES[name]
    k, W = _grid_case(name, nk=11, nw=400)
    try:
        tabs = {}
        for mode_name in ("lane", "warp"):
            s.set_schedule(mode_name)
            tabs[mode_name] = [s.find_roots(m, k, W) for m in case.modes]
    finally:
        s.set_schedule("auto")
    n_acc = 0
    for m, a, b in zip(case.modes, tabs["lane"], tabs["warp"]):
        reg = case.regular(k, W, m)
        oka = reg[a.k_index, a.w_index] & reg[a.k_index, a.w_index + 1]
        okb = reg[b.k_index, b.w_index] & reg[b.k_index, b.w_index + 1]
        # outside the continua: the same brackets, the same classification and the same roots to rounding
        assert np.array_equal(a.k_index[oka], b.k_index[okb]) and np.array_equal(a.w_index[oka], b.w_index[okb])
        assert np.array_equal(a.accepted[oka], b.accepted[okb])
        acc = a.accepted[oka] == 1
        n_acc += int(acc.sum())
        if acc.any():
            wa, wb = a.omega[oka][acc], b.omega[okb][acc]
            assert np.max(np.abs(wa - wb) / np.abs(wa)) < 1e-11
        # inside them D is solver noise: the two summation orders may move or re-file a noise bracket
        assert abs(len(a.omega) - len(b.omega)) <= 0.05 * len(a.omega) + 2
    assert n_acc >= 5
    # each schedule is bit-reproducible, and the scan grids of the two agree to rounding
    s.set_schedule("warp")
    e1, i1 = s.dispersion_grid(case.modes[-1], k, W)
    e2, i2 = s.dispersion_grid(case.modes[-1], k, W)
    assert np.array_equal(e1, e2, equal_nan=True) and np.array_equal(i1, i2, equal_nan=True)
    s.set_schedule("lane")
    e3, i3 = s.dispersion_grid(case.modes[-1], k, W)
    s.set_schedule("auto")
    ok = case.regular(k, W, case.modes[-1]) & np.isfinite(e1)
    assert np.array_equal(np.isnan(e1), np.isnan(e3))
    assert np.max(np.abs(e1[ok] - e3[ok]) / np.abs(e3[ok])) < 1e-13
    assert np.max(np.abs(i1[ok] - i3[ok]) / np.maximum(np.abs(i3[ok]), np.abs(e3[ok]))) < 1e-9


def test_slab_symmetric_and_general_layers():
    """Mirror-symmetric slabs (x0 = 0, every shipped script) integrate the even and the odd solution over
    half the layer; a profile centred off the mid-plane takes the general two-point path.  Both against
    the C oracle, which always integrates the full layer with the reference's two-point condition."""
    k = np.linspace(0.3, 4.0, 10)
    for kind, x0, W, prof, mk in (
            ("slab_density", 0.3, np.linspace(1.75, 2.95, 160), lambda x0: esb.GaussianDensity(0.9, x0),
             lambda x0: ork.make_model("slab_density", width=0.9, x0=x0)),
            ("slab_flow", 0.25, np.linspace(1.25, 2.45, 160), lambda x0: esb.GaussianFlow(1.0, x0),
             lambda x0: ork.make_model("slab_flow", medium=rp.FlowMedium(width=1.0, x0=x0), width=1.0, x0=x0))):
        for shift in (0.0, x0):
            model = mk(shift)
            with esb.DispersionSolver(kind, profile=prof(shift)) as s:
                for mode in (0, 1):
                    e, i = s.dispersion_grid(mode, k, W)
                    e0, i0 = ork.grid(model, mode, k, W)
                    ok = np.isfinite(e0) & np.isfinite(i0)
                    assert np.array_equal(np.isnan(e), ~ok)
                    dev = np.abs((e - i) - (e0 - i0)) / np.maximum(np.abs(e0), np.abs(i0))
                    assert np.nanmax(dev[ok]) < D_TOL, (kind, shift, mode, np.nanmax(dev[ok]))
                    tab = s.find_roots(mode, k, W)
                    bk, bw = ork.brackets(e0 - i0)
                    assert np.array_equal(bk, tab.k_index) and np.array_equal(bw, tab.w_index)
                    for j in np.nonzero(tab.accepted == 1)[0][:6]:
                        kk = k[tab.k_index[j]]
                        r, _, _ = ork.refine(model, mode, kk, kk * W[tab.w_index[j]], kk * W[tab.w_index[j] + 1])
                        assert abs(tab.omega[j] - r) <= ROOT_TOL * abs(r)


def test_convergence_check_flags_sharp_profiles():
    """The fixed-step integrator reports its own discretisation error: negligible for the shipped
    profile width, visible for a much sharper profile, and cured by more steps."""
    k = np.linspace(0.5, 4.5, 9); W = np.linspace(3.0, 4.9, 120)          # above the Alfven continuum
    with esb.DispersionSolver("cylinder_density") as s:
        err, where = s.convergence_check([0, 1, 2], k, W)
        assert err < 1e-9 and where["n_steps_fine"] == 2 * where["n_steps"]
    # a dense shell of width 0.05 at r = 0.5, in the uniform part of the mesh (6 steps per width)
    sharp = esb.GaussianDensity(0.05, x0=-0.5)
    W2 = np.linspace(4.6, 4.95, 60)           # above the Alfven continuum (vA = 4.41 outside the shell)
    with esb.DispersionSolver("cylinder_density", profile=sharp) as s:
        err_default, _ = s.convergence_check([1], k, W2)
    with esb.DispersionSolver("cylinder_density", profile=sharp, n_steps=576) as s:
        err_fine, _ = s.convergence_check([1], k, W2)
    assert err_default > 1e-10                # the default step count is visibly too coarse for it ...
    assert err_fine < 1e-3 * err_default      # ... and 4x the steps (8th order) cure it


def test_discretisation_guard_flags_sharp_profiles():
    """The guard built into every sweep (esb_set_guard_fields: about 32 k (point, mode) samples re-evaluated
    at 2 x n_steps on a side stream) reports the discretisation error without anyone calling
    convergence_check: below 1e-9 for every shipped equilibrium, visible - and warned about by the
    host-returning root search - for the sharp shell, silent again with enough steps."""
    import warnings as _w
    k = np.linspace(0.5, 4.5, 64); W = np.linspace(0.5, 5.0, 512)       # the bench window, continua included
    for name, case in CASES.items():
        kk = np.linspace(0.25, 4.0, 64) if case.kind == "cylinder_rotation" else k
        WW = np.linspace(case.W[0], case.W[1], 512)
        with case.gpu_solver(esb) as s:
            with _w.catch_warnings():
                _w.simplefilter("error", esb.DiscretisationWarning)
                s.find_roots_multi(list(case.modes)[:2], kk, WW)
            rep = s.guard_report()
            assert rep["stride"] == 64 and rep["n_checked"] >= 100, (name, rep)
            assert rep["worst"] < 1e-9 and rep["n_above"] == 0, (name, rep)
    sharp = esb.GaussianDensity(0.05, x0=-0.5)
    W2 = np.linspace(4.6, 4.95, 512)          # above the Alfven continuum (vA = 4.41 outside the shell)
    with esb.DispersionSolver("cylinder_density", profile=sharp) as s:
        with pytest.warns(esb.DiscretisationWarning):
            s.find_roots_multi([0, 1], k, W2)
        rep = s.guard_report()
        err, _ = s.convergence_check([0, 1], k, W2)
        assert rep["worst"] > 1e-9 and rep["n_above"] > 0
        assert err > 1e-9                        # the explicit check (all points, D itself) sees it too
    with esb.DispersionSolver("cylinder_density", profile=sharp, n_steps=576) as s:
        with _w.catch_warnings():
            _w.simplefilter("error", esb.DiscretisationWarning)
            s.find_roots_multi([0, 1], k, W2)
        assert s.guard_report()["worst"] < 1e-9
    # worker-sized sweeps are not sampled; a solver built without the guard reports stride 0
    with esb.DispersionSolver("cylinder_density") as s:
        s.find_roots("kink", k[:1], W2[:90])
        assert s.guard_report()["n_checked"] == 0
    with esb.DispersionSolver("cylinder_density", guard=0) as s:
        s.find_roots_multi([0, 1], k, W2)
        assert s.guard_report()["stride"] == 0


def test_bessel_jy_and_leaky_exterior_on_the_device(solvers):
    """The J_n / Y_n evaluators and the closed-form leaky exterior as the DEVICE build computes them
    (esb_bessel_jy_dev, esb_exterior_leaky_dev) against scipy and against the host build."""
    import ctypes as C
    from scipy import special as sp
    s = solvers["cylinder_density"]
    x = np.concatenate([np.geomspace(1e-3, 5.0, 400), np.linspace(4.9, 8.1, 400), np.geomspace(8.0, 5e3, 800)])
    for n in range(4):
        got = s.bessel_jy_device(n, x)
        J, dJ, Y, dY = sp.jv(n, x), sp.jvp(n, x), sp.yv(n, x), sp.yvp(n, x)
        amp, damp = np.hypot(J, Y), np.hypot(dJ, dY)
        assert np.max(np.abs(got[:, 0] - J) / amp) < 5e-15 and np.max(np.abs(got[:, 2] - Y) / amp) < 5e-15
        assert np.max(np.abs(got[:, 1] - dJ) / damp) < 5e-15 and np.max(np.abs(got[:, 3] - dY) / damp) < 5e-15
    rng = np.random.default_rng(3)
    k = rng.uniform(0.05, 4.5, 2000)
    W = np.where(rng.random(2000) < 0.7, rng.uniform(5.001, 9.0, 2000), rng.uniform(0.4976, 0.4999, 2000))
    out = (C.c_double * 2)()
    for n in range(4):
        dev = s.exterior_leaky_device(n, k, k * W)
        assert np.isfinite(dev).all()
        for j in range(0, 2000, 40):
            assert s.lib.esb_exterior_leaky(C.byref(s.model), n, k[j], k[j] * W[j], out) == 0
            amp = np.hypot(out[0], out[1])
            # (the exterior starts thousands of radii out at small k: rounding of the phase, eps x z0)
            assert abs(dev[j, 0] - out[0]) < 1e-10 * amp and abs(dev[j, 1] - out[1]) < 1e-10 * amp
    # the regular side is not theirs
    assert np.isnan(s.exterior_leaky_device(1, np.array([1.0]), np.array([3.0]))).all()


def test_leaky_grid_on_the_device(solvers):
    """esb_dispersion_grid_leaky: the whole grid, m_e < 0 included, against the C oracle with its skip rule
    lifted; on the regular side (m_e >= 0) the values of the default evaluation."""
    for name, windows in (("cylinder_density", [(5.02, 8.0), (0.4976, 0.4999)]), ("slab_flow", [(2.52, 4.0)]),
                          ("cylinder_rotation", [(1.52, 3.2)])):
        case = CASES[name]
        s = solvers[name]
        model = case.c_model()
        model.leaky = 1
        k = np.linspace(0.3, 4.0, 9)
        W = np.concatenate([np.linspace(a, b, 40) for a, b in windows] + [np.linspace(*case.roots_window, 40)])
        modes = list(case.modes)[:2]
        E, I = s.dispersion_grid_leaky(modes, k, W)
        E0, I0 = s.dispersion_grid_multi(modes, k, W)
        skipped = np.isnan(E0)
        assert skipped.any() and (~skipped).any() and np.isfinite(E).all()
        assert np.allclose(E[~skipped], E0[~skipped], rtol=1e-12) and np.allclose(I[~skipped], I0[~skipped], rtol=1e-9)
        for slot, mode in enumerate(modes):
            e0, i0 = ork.grid(model, mode, k, W)
            reg = case.regular(k, W, mode, margin=0.03) & skipped[slot]
            assert reg.sum() > 100
            dev = np.abs((E[slot] - I[slot]) - (e0 - i0)) / np.maximum(np.abs(e0), np.abs(i0))
            assert np.quantile(dev[reg], 0.9) < 1e-10 and dev[reg].max() < 1e-7, (name, mode, dev[reg].max())


def test_device_pointer_entry_points(solvers):
    """esb_dispersion_grid_dev / esb_brackets_dev on caller-owned device buffers (torch tensors) and a
    caller-owned stream: the same grids and the same sorted bracket list as the host entry points."""
    import ctypes as C
    import torch
    s = solvers["cylinder_density"]
    lib = s.lib
    dev = torch.device("cuda", 0)
    for nk, nw in ((7, 300), (40, 2500)):          # warp-per-point and thread-per-point scan kernels
        k = np.linspace(0.5, 4.0, nk); W = np.linspace(0.55, 4.95, nw)
        e_host, i_host = s.dispersion_grid(1, k, W)
        tab = s.find_roots(1, k, W)
        stream = torch.cuda.Stream(dev)
        d_k = torch.from_numpy(k).to(dev); d_w = torch.from_numpy(W).to(dev)
        d_e = torch.empty((nk, nw), dtype=torch.float64, device=dev); d_i = torch.empty_like(d_e)
        torch.cuda.synchronize()
        p = lambda t: C.c_void_p(t.data_ptr())
        rc = lib.esb_dispersion_grid_dev(s.ctx, 1, p(d_k), nk, p(d_w), nw, 1, p(d_e), p(d_i),
                                         C.c_void_p(stream.cuda_stream))
        assert rc == 0
        d_off = torch.empty(nk + 1, dtype=torch.int32, device=dev)
        cap = len(tab.omega) + 8
        d_bk = torch.empty(cap, dtype=torch.int32, device=dev); d_bw = torch.empty_like(d_bk)
        n = C.c_int32(-1)
        rc = lib.esb_brackets_dev(s.ctx, p(d_e), p(d_i), nk, nw, p(d_off), p(d_bk), p(d_bw), cap, C.byref(n),
                                  C.c_void_p(stream.cuda_stream))
        assert rc == 0
        stream.synchronize()
        assert np.array_equal(d_e.cpu().numpy(), e_host, equal_nan=True)
        assert np.array_equal(d_i.cpu().numpy(), i_host, equal_nan=True)
        assert n.value == len(tab.omega)
        assert np.array_equal(d_bk[:n.value].cpu().numpy(), tab.k_index)
        assert np.array_equal(d_bw[:n.value].cpu().numpy(), tab.w_index)
        off = d_off.cpu().numpy()
        assert off[0] == 0 and off[-1] == n.value
        assert np.array_equal(np.diff(off), np.bincount(tab.k_index, minlength=nk))
        # too small a capacity is reported, the count is still returned
        rc = lib.esb_brackets_dev(s.ctx, p(d_e), p(d_i), nk, nw, p(d_off), p(d_bk), p(d_bw), 3, C.byref(n),
                                  C.c_void_p(stream.cuda_stream))
        assert rc == -3 and n.value == len(tab.omega)
    # a context can be moved onto the caller's stream
    st = torch.cuda.Stream(dev)
    with esb.DispersionSolver("slab_density") as t:
        a = t.find_roots(0, np.linspace(0.3, 1.5, 9), np.linspace(1.75, 2.95, 200))
        t.set_stream(st.cuda_stream)
        b = t.find_roots(0, np.linspace(0.3, 1.5, 9), np.linspace(1.75, 2.95, 200))
        assert np.array_equal(a.omega, b.omega, equal_nan=True) and np.array_equal(a.accepted, b.accepted)


def test_rk4_and_rk8_agree():
    k = np.linspace(0.3, 4.0, 6); W = np.linspace(3.0, 4.9, 40)
    with esb.DispersionSolver("cylinder_density", scheme="rk8") as a, \
            esb.DispersionSolver("cylinder_density", scheme="rk4", n_steps=2048) as b:
        ea, ia = a.dispersion_grid(1, k, W)
        eb, ib = b.dispersion_grid(1, k, W)
    assert np.array_equal(ea, eb)
    assert np.max(np.abs(ia - ib) / np.abs(ia)) < 5e-9
    # the classical scheme on the other second-order kinds (generic, unscaled coefficient path)
    for kind, kw, W in (("slab_density", {}, np.linspace(1.75, 2.95, 40)),
                        ("cylinder_flow", {}, np.linspace(3.0, 4.9, 40)),
                        ("slab_flow", dict(profile=esb.GaussianFlow(1.0)), np.linspace(1.25, 2.45, 40))):
        with esb.DispersionSolver(kind, scheme="rk8", **kw) as a, \
                esb.DispersionSolver(kind, scheme="rk4", n_steps=2048, **kw) as b:    # 200 KB table limit
            ea, ia = a.dispersion_grid(1, k, W)
            eb, ib = b.dispersion_grid(1, k, W)
        assert np.array_equal(ea, eb, equal_nan=True)
        assert np.nanmax(np.abs(ia - ib) / np.abs(ia)) < 1e-6, kind           # 4th order at 2048 steps


#: BASELINE.json configs[0..3]: (case, modes, k axis, W axis) at full size
BASELINE_GRIDS = {
    "configs[0]": ("slab_density", [0, 1], np.linspace(0.001, 0.75, 200), np.linspace(0.41, 2.95, 2000)),
    "configs[1]": ("cylinder_density", [0, 1, 2, 3], np.linspace(0.01, 4.5, 1000), np.linspace(0.5, 5.0, 10000)),
    "configs[2]": ("slab_flow", [0, 1], np.linspace(0.01, 4.5, 2000), np.linspace(-2.7, 2.7, 20000)),
    "configs[3]": ("cylinder_rotation", [0, 1, 2, 3], np.linspace(0.25, 4.0, 2000), np.linspace(0.40, 1.6, 20000)),
}


@pytest.mark.parametrize("config", list(BASELINE_GRIDS))
def test_oracle_samples_on_baseline_grids(solvers, config):
    """The oracle ON the BASELINE coordinates: 40 rows of the full-size k axis (the smallest wavenumbers
    0.01 <= k < 0.05 included, where the exterior starts at |r| = 1885) x 400 columns of the full-size
    omega axis, evaluated by the throughput kernel (one thread per point, every mode fused, n = 3
    included), every point above the noise floor compared with the C oracle: > 2000 points per mode."""
    name, modes, k_full, W_full = BASELINE_GRIDS[config]
    case, s = CASES[name], solvers[name]
    rng = np.random.default_rng(2024)
    small = np.nonzero(k_full < 0.05)[0]
    rows = np.unique(np.concatenate([small[:: max(1, len(small) // 8)][:8] if len(small) else [0],
                                     rng.choice(len(k_full), 34, replace=False)]))
    cols = np.sort(rng.choice(len(W_full), 400, replace=False))
    k, W = k_full[rows], W_full[cols]
    model = case.c_model()
    s.set_schedule("lane")
    try:
        fused = modes[:3]
        E, I = s.dispersion_grid_multi(fused, k, W)
        grids = {m: (E[j], I[j]) for j, m in enumerate(fused)}
        for m in modes[3:]:
            grids[m] = s.dispersion_grid(m, k, W)
    finally:
        s.set_schedule("auto")
    for m in modes:
        e, i = grids[m]
        e0, i0 = ork.grid(model, m, k, W)
        fin = np.isfinite(e0) & np.isfinite(i0)
        assert np.array_equal(np.isfinite(e) & np.isfinite(i), fin)
        ok = case.regular(k, W, m) & fin
        assert ok.sum() > 2000, (config, m, ok.sum())
        dev = (np.abs((e - i) - (e0 - i0)) / np.maximum(np.abs(e0), np.abs(i0)))[ok]
        # 99.9 % within 1e-10, every point within D_TOL except the grid points next to a pole of D (Y -> 0:
        # int = N/Y amplifies the 1e-10 error of Y), which stay within 2e-8
        assert np.quantile(dev, 0.999) < 2e-10, (config, m, np.quantile(dev, 0.999))
        assert (dev > D_TOL).sum() <= 2 and dev.max() < 2e-8, (config, m, dev.max(), (dev > D_TOL).sum())
        if len(small):
            lo = ok & (k[:, None] < 0.05)
            assert lo.sum() > 300 and (np.abs((e - i) - (e0 - i0)) / np.maximum(np.abs(e0), np.abs(i0)))[lo].max() < D_TOL
        # the brackets of these rows: identical index sets above the noise floor, every accepted root checked
        tab = s.find_roots(m, k, W)
        ok_iv = ok[:, :-1] & ok[:, 1:]
        ok_, ow_ = ork.brackets(e0 - i0)
        sel_o, sel_g = ok_iv[ok_, ow_], ok_iv[tab.k_index, tab.w_index]
        assert np.array_equal(ok_[sel_o], tab.k_index[sel_g]) and np.array_equal(ow_[sel_o], tab.w_index[sel_g])
        for j in np.nonzero(sel_g & (tab.accepted == 1))[0]:
            kk = k[tab.k_index[j]]
            r, _, _ = ork.refine(model, m, kk, kk * W[tab.w_index[j]], kk * W[tab.w_index[j] + 1])
            assert abs(tab.omega[j] - r) <= ROOT_TOL * abs(r), (config, m, kk, r, tab.omega[j])


def test_full_size_properties(solvers):
    """BASELINE configs[1] size (1000 k x 10000 omega): properties that need no oracle."""
    s = solvers["cylinder_density"]
    k = np.linspace(0.01, 4.5, 1000)
    W = np.linspace(0.40, 5.05, 10000)
    e1, i1 = s.dispersion_grid(1, k, W)
    e2, i2 = s.dispersion_grid(1, k, W)
    assert np.array_equal(e1, e2, equal_nan=True) and np.array_equal(i1, i2, equal_nan=True)   # deterministic
    # skipped region = exactly where m_e < 0 :  cT_e <= W <= c_e  or  W >= vA_e
    md = s.medium
    leaky = ((W > md.cT_e) & (W < md.c_e)) | (W > md.vA_e)
    assert np.array_equal(np.isnan(e1).all(axis=0), leaky)
    # D is even in omega for the cylinder (only omega^2 enters)
    em, im = s.dispersion_grid(1, k[::50], -W[::10])
    assert np.array_equal(em, e1[::50, ::10], equal_nan=True) and np.array_equal(im, i1[::50, ::10], equal_nan=True)
    # homogeneity: doubling the exterior initial values doubles D exactly (power of two)
    with esb.DispersionSolver("cylinder_density", ext_ic=(2e-8, 2e-15)) as s2:
        e3, i3 = s2.dispersion_grid(1, k[::50], W[::10])
    assert np.array_equal(e3, 2 * e1[::50, ::10], equal_nan=True)
    assert np.array_equal(i3, 2 * i1[::50, ::10], equal_nan=True)
    # root table: count == numpy's on the grid, sorted, inside its bracket, acceptance consistent
    tab = s.find_roots(1, k, W)
    bk, bw = ork.brackets(e1 - i1)
    assert np.array_equal(bk, tab.k_index) and np.array_equal(bw, tab.w_index)
    key = tab.k_index.astype(np.int64) * len(W) + tab.w_index
    assert np.all(np.diff(key) > 0)
    lo = k[tab.k_index] * W[tab.w_index]; hi = k[tab.k_index] * W[tab.w_index + 1]
    assert np.all((tab.omega >= lo) & (tab.omega <= hi))
    pct = np.abs(tab.ext - tab.intq) * 100 / np.maximum(np.abs(tab.ext), np.abs(tab.intq))
    fin = np.isfinite(pct)
    assert np.array_equal(tab.accepted[fin] == 1, pct[fin] < 1.0)
    assert tab.accepted.sum() > 2000
    # sharding the k axis (what the multi-GPU path does) changes nothing (one schedule for all three
    # calls: by size the halves may take the warp-per-bracket refinement, which agrees to rounding only)
    try:
        s.set_schedule("lane")
        tab = s.find_roots(1, k, W)
        lo_half = s.find_roots(1, k[:500], W)
        hi_half = s.find_roots(1, k[500:], W)
    finally:
        s.set_schedule("auto")
    assert np.array_equal(np.concatenate([lo_half.omega, hi_half.omega]), tab.omega)
    assert np.array_equal(np.concatenate([lo_half.k_index, hi_half.k_index + 500]), tab.k_index)


def _check_table_against_grid(tab, k, W, D, tol_percent=1.0):
    """Size-independent root-table invariants: brackets == numpy's on the grid, sorted, every refined
    root inside its bracket, acceptance flag == the reference's test on the stored (ext, int)."""
    bk, bw = ork.brackets(D)
    assert np.array_equal(bk, tab.k_index) and np.array_equal(bw, tab.w_index)
    key = tab.k_index.astype(np.int64) * len(W) + tab.w_index
    assert np.all(np.diff(key) > 0)
    a = k[tab.k_index] * W[tab.w_index]; b = k[tab.k_index] * W[tab.w_index + 1]
    lo, hi = np.minimum(a, b), np.maximum(a, b)
    assert np.all((tab.omega >= lo) & (tab.omega <= hi))
    pct = np.abs(tab.ext - tab.intq) * 100 / np.maximum(np.abs(tab.ext), np.abs(tab.intq))
    fin = np.isfinite(pct)
    assert np.array_equal(tab.accepted[fin] == 1, pct[fin] < tol_percent)


def _leaky(md, W):
    """Columns the reference skips: m_e^2 (W) < 0 (Density_cylinder.py:699,760)."""
    W2 = W * W
    cT2 = md.cT_e**2
    return (md.vA_e**2 - W2) * (md.c_e**2 - W2) / (cT2 - W2) < 0


def test_full_size_slab_flow_properties():
    """BASELINE configs[2] size: slab with a sheared flow, backward and forward branches,
    2000 k x 20000 omega."""
    k = np.linspace(0.01, 4.5, 2000)
    W = np.linspace(-2.7, 2.7, 20000)
    md = esb.FlowMedium(U_i0=0.35)
    with esb.DispersionSolver("slab_flow", medium=md, profile=esb.GaussianFlow(1.0)) as s:
        e, i = s.dispersion_grid(0, k, W)
        # skip rule m_e < 0 (flow script :205), here with U_e = 0
        assert np.array_equal(np.isnan(e).all(axis=0), _leaky(md, W))
        tab = s.find_roots(0, k, W)
        _check_table_against_grid(tab, k, W, e - i)
        assert tab.accepted.sum() > 2000
        # both branches are populated
        assert (tab.omega[tab.accepted == 1] > 0).sum() > 500 and (tab.omega[tab.accepted == 1] < 0).sum() > 500
        # sausage and kink share one integration: the fused scan returns the single-mode grids
        E, I = s.dispersion_grid_multi([0, 1], k[::40], W[::20])
        assert np.array_equal(E[0], e[::40, ::20], equal_nan=True) and np.array_equal(I[0], i[::40, ::20], equal_nan=True)
    # flow reversal: the backward branch of U is the forward branch of -U; every factor is odd or
    # even in (omega - k U), so (ext, int)(-omega; -U) = -(ext, int)(omega; U) bit for bit
    with esb.DispersionSolver("slab_flow", medium=esb.FlowMedium(U_i0=-0.35), profile=esb.GaussianFlow(1.0)) as r:
        er, ir = r.dispersion_grid(0, k[::40], -W[::20])
    assert np.array_equal(er, -e[::40, ::20], equal_nan=True) and np.array_equal(ir, -i[::40, ::20], equal_nan=True)


def test_full_size_rotation_properties():
    """BASELINE configs[3] size: cylinder with rotational flow, n = 0..3, 2000 k x 20000 omega
    (n = 0, 1, 2 in one fused scan, n = 3 on its own)."""
    k = np.linspace(0.25, 4.0, 2000)
    W = np.linspace(0.40, 1.6, 20000)
    rot = esb.PowerLawRotation(0.15, 1.25)
    sub = (slice(None, None, 50), slice(None, None, 25))
    with esb.DispersionSolver("cylinder_rotation", profile=rot, s_end=0.01) as s:
        tabs = s.find_roots_multi([0, 1, 2], k, W)
        e3, i3 = s.dispersion_grid(3, k, W)
        t3 = s.find_roots(3, k, W)
        _check_table_against_grid(t3, k, W, e3 - i3)
        assert np.array_equal(np.isnan(e3).all(axis=0), _leaky(s.medium, W))
        for m, t in enumerate(tabs):
            e, i = s.dispersion_grid(m, k[sub[0]], W)
            one = s.find_roots(m, k[sub[0]], W)
            rows = np.isin(t.k_index, np.arange(len(k))[sub[0]])
            # the fused scan's table restricted to those rows == the single-mode table of the rows
            assert rows.sum() > 0 and abs(int(rows.sum()) - len(one.omega)) <= 0.02 * len(one.omega) + 2
            _check_table_against_grid(one, k[sub[0]], W, e - i)
            assert t.accepted.sum() > 200
        e0, i0 = s.dispersion_grid(0, k[sub[0]], W[sub[1]])
        em, im = s.dispersion_grid(0, k[sub[0]], -W[sub[1]])
        # n = 0 carries no Doppler shift: D is even in omega
        assert np.array_equal(e0, em, equal_nan=True) and np.array_equal(i0, im, equal_nan=True)
        e1, i1 = s.dispersion_grid(1, k[sub[0]], W[sub[1]])
    # reversing the rotation maps the forward branch onto the backward one: (omega, v_phi) -> (-omega, -v_phi)
    # leaves Om^2, T, Q, C1, C2, C3 unchanged, bit for bit
    with esb.DispersionSolver("cylinder_rotation", profile=esb.PowerLawRotation(-0.15, 1.25), s_end=0.01) as r:
        er, ir = r.dispersion_grid(1, k[sub[0]], -W[sub[1]])
    assert np.array_equal(er, e1, equal_nan=True) and np.array_equal(ir, i1, equal_nan=True)


def test_device_side_gather_equals_host_table(solvers):
    """The NCCL gather used by the multi-GPU bench reads the root table straight from the
    library's device buffers; with one rank it must reproduce the host-side table."""
    import socket
    import torch
    import torch.distributed as dist
    from eigensolver_b200.distributed import gather_modes_device, gather_root_tables_device
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dev = torch.device("cuda", 0)
    dist.init_process_group("nccl", rank=0, world_size=1, device_id=dev)
    try:
        s = solvers["cylinder_density"]
        k = np.linspace(0.5, 4.0, 33); W = np.linspace(2.95, 4.95, 257)
        s.upload_axes(k, W)
        ns = s.sweep_resident_multi([0, 1])
        for slot, n in enumerate(ns):
            host = s.download_roots(n, slot)
            g = gather_root_tables_device(s, slot, 1000, dev).cpu().numpy()
            assert g.shape == (n, 3)
            assert np.array_equal(g[:, 0], host.k_index + 1000.0)
            assert np.array_equal(g[:, 1], host.omega)
            assert np.array_equal(g[:, 2], host.accepted.astype(np.float64))
            # modes only, strided global rows, sorted: what the multi-GPU bench gathers
            a = gather_root_tables_device(s, slot, 3, dev, k_stride=8, accepted_only=True, sort=True).cpu().numpy()
            acc = host.accepted == 1
            assert a.shape == (int(acc.sum()), 2)
            assert np.array_equal(a[:, 0], host.k_index[acc] * 8.0 + 3.0)
            assert np.array_equal(a[:, 1], host.omega[acc])
        # all slots in one exchange: (global row, omega, slot), sorted by (slot, row, omega)
        allm = gather_modes_device(s, len(ns), 3, dev, k_stride=8, sort=True).cpu().numpy()
        hosts = [s.download_roots(n, slot) for slot, n in enumerate(ns)]
        want = np.concatenate([np.stack([h.k_index[h.accepted == 1] * 8.0 + 3.0, h.omega[h.accepted == 1],
                                         np.full(int((h.accepted == 1).sum()), float(slot))], axis=1)
                               for slot, h in enumerate(hosts)])
        assert np.array_equal(allm, want)
        # the payload itself (esb_pack_modes_dev): header row = (rows, entries scanned, overflow flag); a
        # capacity that is too small is reported, not overrun
        import ctypes as C
        n_acc = len(want)
        for cap in (n_acc + 5, n_acc, max(n_acc - 7, 1)):
            buf = torch.full((cap + 3, 3), -1.0, dtype=torch.float64, device=dev)
            rc = s.lib.esb_pack_modes_dev(s.ctx, len(ns), 3.0, 8.0, C.c_void_p(buf.data_ptr()), cap, None)
            assert rc == 0
            s.lib.esb_tables_wait(s.ctx, None)
            torch.cuda.synchronize()
            b = buf.cpu().numpy()
            assert b[0, 0] == min(n_acc, cap) and b[0, 1] == sum(ns) and b[0, 2] == (1.0 if n_acc > cap else 0.0)
            assert np.array_equal(b[1: 1 + min(n_acc, cap)], want[: min(n_acc, cap)])
            assert (b[1 + cap:] == -1.0).all()                       # nothing written past the capacity
        # a parameter scan left on the device and gathered from there == the downloaded compact table
        from eigensolver_b200.scan import density_flow_grid, gather_scan_modes_device, parameter_scan
        dens, _ = density_flow_grid([0.15, 0.3], [0.5])
        with esb.DispersionSolver("cylinder_density") as sc:
            host = parameter_scan(sc, dens, k, W, [0, 1])
            acc = np.asarray(host.table["accepted"]) == 1
            want = np.stack([np.asarray(host.table[c])[acc].astype(np.float64) for c in ("model", "slot", "k_index", "omega")], axis=1)
            res = parameter_scan(sc, dens, k, W, [0, 1], download=False)
            assert res.table is None
            got = gather_scan_modes_device(sc, res, dev).cpu().numpy()
        assert len(want) > 0 and np.array_equal(got, want)
    finally:
        dist.destroy_process_group()


def test_parameter_scan_guards_the_ends_of_the_range():
    """The batched scan job is not sampled by the discretisation guard; guard_ends=True sweeps the two ends of
    the parameter range on their own: clean for a family of shipped-like profile widths, flagged for a family
    of sharp shells."""
    import warnings as _w
    from eigensolver_b200.scan import parameter_scan
    k = np.linspace(0.5, 4.5, 48); W = np.linspace(4.6, 4.95, 256)      # above the Alfven continuum
    smooth = [dict(profile=esb.GaussianDensity(w), label={"width": w}) for w in (0.9, 0.95, 1.0)]
    sharp = [dict(profile=esb.GaussianDensity(w, x0=-0.5), label={"shell width": w}) for w in (0.05, 0.04, 0.03)]
    with esb.DispersionSolver("cylinder_density") as s:
        with _w.catch_warnings():
            _w.simplefilter("error", esb.DiscretisationWarning)
            res = parameter_scan(s, smooth, k, W, [0, 1], guard_ends=True)
        assert len(res.guard) == 2 and all(r["n_checked"] > 50 and r["worst"] < 1e-9 for r in res.guard)
        assert s.spec.profile == esb.GaussianDensity(0.95)          # the solver's own equilibrium is back
        assert len(res.points) == 3 and res.table is not None
        with pytest.warns(esb.DiscretisationWarning):
            res = parameter_scan(s, sharp, k, W, [0, 1], guard_ends=True)
        assert min(r["worst"] for r in res.guard) > 1e-9


def test_parameter_scan_matches_fresh_solvers():
    """configs[4]-style scan through eigensolver_b200.scan: the batched job gives what a freshly built
    solver gives for every equilibrium, and sharding the WAVENUMBERS (strided) over ranks covers the
    scan exactly once."""
    from eigensolver_b200.scan import density_flow_grid, parameter_scan
    dens, flow = density_flow_grid([0.15, 0.2055, 0.3], [0.2, 0.5, 0.9])
    k = np.linspace(0.4, 4.0, 12)
    for kind, pts, modes, W in (("cylinder_density", dens, [0, 1, 2], np.linspace(0.55, 4.5, 300)),
                                ("slab_flow", flow, [0, 1], np.linspace(1.25, 2.45, 200))):
        with esb.DispersionSolver(kind) as s:
            s.set_schedule("lane")
            res = parameter_scan(s, pts, k, W, modes, keep_tables=True)
            halves = [parameter_scan(s, pts, k, W, modes, rank=r, world=2, keep_tables=True) for r in range(2)]
        assert [p.label for p in res.points] == [p["label"] for p in pts]
        for i in range(len(pts)):
            for m in range(len(modes)):
                assert halves[0].points[i].n_brackets[m] + halves[1].points[i].n_brackets[m] == res.points[i].n_brackets[m]
                # rank r owns rows r, r+2, ...: the union of the two shards is the single-rank table
                rows = np.concatenate([h.points[i].tables[m]["k_index"] * h.k_stride + h.k_offset for h in halves])
                om = np.concatenate([h.points[i].tables[m]["omega"] for h in halves])
                full = res.points[i].tables[m]
                order = np.lexsort((om, rows))
                order_full = np.lexsort((full["omega"], full["k_index"]))
                assert np.array_equal(rows[order], full["k_index"][order_full])
                assert np.array_equal(om[order], full["omega"][order_full], equal_nan=True)
        for p, r in zip(pts, res.points):
            with esb.DispersionSolver(kind, medium=p["medium"], profile=p["profile"]) as fresh:
                fresh.set_schedule("lane")
                tabs = fresh.find_roots_multi(modes, k, W)
            for a, b in zip(tabs, r.tables):
                assert np.array_equal(a.k_index, b["k_index"]) and np.array_equal(a.omega, b["omega"], equal_nan=True)
            assert sum(r.n_modes) > 0
    # the contrast helper reproduces the reference's own rho_e for its own vA_e
    from eigensolver_b200.scan import medium_for_density_contrast
    m = medium_for_density_contrast(esb.CYLINDER_CORONAL, esb.CYLINDER_CORONAL.rho_e)
    assert abs(m.vA_e - 5.0) < 1e-12


def test_resolve_steps_restores_the_tolerance_on_a_sharp_profile():
    """Error control in place of odeint's adaptivity (Density_cylinder.py:783): resolve_steps raises n_steps,
    by the factor the order predicts from the guard's own measurement, until the discretisation error of the
    window is below 1e-9 - confirmed independently by convergence_check (D itself, a second context) and by
    the next root search staying silent; the shipped profile needs nothing; a shell no staged table can
    resolve is reported, not hidden."""
    import warnings as _w
    k = np.linspace(0.5, 4.5, 9); W2 = np.linspace(4.6, 4.95, 60)         # a worker-sized window
    sharp = esb.GaussianDensity(0.05, x0=-0.5)
    with esb.DispersionSolver("cylinder_density", profile=sharp) as s:
        n0 = int(s.model.n_steps)
        err0, _ = s.convergence_check([0, 1], k, W2)
        with _w.catch_warnings():
            _w.simplefilter("error", esb.DiscretisationWarning)
            out = s.resolve_steps([0, 1], k, W2)
        assert out["resolved"] and out["n_steps"] == int(s.model.n_steps) > n0, out
        assert out["history"][0] == (n0, out["history"][0][1]) and out["history"][0][1] > 1e-9, out
        assert out["worst"] <= 1e-9 and len(out["history"]) <= 3, out
        err1, where = s.convergence_check([0, 1], k, W2)
        # (convergence_check takes the maximum of D's own relative deviation over ALL points, the ones next to a
        # pole of D included, where int = N / Y amplifies any error without bound: it falls by the same factor)
        assert where["n_steps"] == out["n_steps"] and err1 < 1e-3 * err0, (err0, err1, out)
        with _w.catch_warnings():
            _w.simplefilter("error", esb.DiscretisationWarning)
            tabs = s.find_roots_multi([0, 1], np.linspace(0.5, 4.5, 64), np.linspace(4.6, 4.95, 512))
        assert s.guard_report()["n_checked"] > 0 and sum(len(t.omega) for t in tabs) > 0
    with esb.DispersionSolver("cylinder_density") as s:
        out = s.resolve_steps([0, 1, 2], k, np.linspace(3.0, 4.9, 120))
        assert out["resolved"] and out["n_steps"] == n0 and len(out["history"]) == 1, out
    with esb.DispersionSolver("cylinder_density", profile=esb.GaussianDensity(0.004, x0=-0.5)) as s:
        with pytest.warns(esb.DiscretisationWarning):
            out = s.resolve_steps([1], k, W2)
        cap = s.spec.max_steps("rk8") // 2          # the guard's table of twice the steps is staged as well
        assert not out["resolved"] and out["n_steps"] == cap - cap % 2 and out["worst"] > 1e-9, out
    with esb.DispersionSolver("cylinder_density", guard=0) as s:
        with pytest.raises(ValueError):
            s.resolve_steps([1], k, W2)
