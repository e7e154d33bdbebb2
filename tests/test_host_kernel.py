"""CPU parity tests of the kernels' arithmetic: csrc/core.cuh compiled for the host
(tests/host_kernel.py) against the C oracle, on the grids the GPU parity tests use and on the
BASELINE coordinates (k from 0.01, every mode incl. n = 3).  The GPU suite shows that the device
build agrees with this host build; together they tie the GPU results to the oracle."""
import numpy as np
import pytest

import eigensolver_b200 as esb
import host_kernel as hk
from helpers import CASES
from oracle import rk_oracle as ork

D_TOL = 1e-9


def spec_of(case, **kw):
    """the ModelSpec a GPU solver of this case would upload"""
    class _Capture:
        """stands in for the package: records the arguments of DispersionSolver(...)"""
        def __getattr__(self, name):
            return getattr(esb, name)

        @staticmethod
        def DispersionSolver(kind, **k):
            k.pop("device", None)
            return esb.ModelSpec(kind, **k)
    return case.gpu_solver(_Capture(), **kw)


@pytest.mark.parametrize("name", list(CASES))
def test_host_build_matches_c_oracle(name):
    case = CASES[name]
    model = case.c_model()
    k = np.linspace(0.05, 4.5, 20)
    W = np.linspace(case.W[0], case.W[1], 240)
    modes = list(case.modes)
    e, i, _ = hk.grid(spec_of(case), modes, k, W)
    for slot, mode in enumerate(modes):
        e0, i0 = ork.grid(model, mode, k, W)
        fin = np.isfinite(e0) & np.isfinite(i0)
        assert np.array_equal(np.isfinite(e[slot]) & np.isfinite(i[slot]), fin)
        ok = case.regular(k, W, mode) & fin
        dev = np.abs((e[slot] - i[slot]) - (e0 - i0)) / np.maximum(np.abs(e0), np.abs(i0))
        assert ok.sum() > case.min_regular * ok.size
        assert np.nanmax(dev[ok]) < D_TOL, (mode, np.nanmax(dev[ok]))


@pytest.mark.parametrize("name,kmin,kmax,modes", [("cylinder_density", 0.01, 4.5, [0, 1, 2, 3]),
                                                  ("slab_density", 0.001, 0.75, [0, 1]),
                                                  ("slab_flow", 0.01, 4.5, [0, 1]),
                                                  ("cylinder_rotation", 0.25, 4.0, [0, 1, 2, 3]),
                                                  ("cylinder_flow", 0.01, 4.0, [0, 1, 2, 3])])
def test_host_build_on_baseline_coordinates(name, kmin, kmax, modes):
    """Random (k, omega) from the BASELINE k ranges - 40 % of them in the lowest decade of k, where the
    exterior solution starts thousands of radii out - and every azimuthal order up to n = 3."""
    case = CASES[name]
    model = case.c_model()
    rng = np.random.default_rng(11)
    k = np.concatenate([rng.uniform(kmin, min(5 * kmin, kmax), 160), rng.uniform(kmin, kmax, 240)])
    W = rng.uniform(case.W[0], case.W[1], k.size)
    sp = spec_of(case)
    for mode in modes:
        e, i, _ = hk.evaluate(sp, [mode], k, k * W)
        ref = np.array([ork.point(model, mode, kk, kk * ww) for kk, ww in zip(k, W)])
        e0, i0 = ref[:, 0], ref[:, 1]
        fin = np.isfinite(e0) & np.isfinite(i0)
        assert np.array_equal(np.isfinite(e[0]) & np.isfinite(i[0]), fin)
        ok = fin & np.array([case.regular(kk, np.array([ww]), mode)[0, 0] for kk, ww in zip(k, W)])
        assert ok.sum() > 100
        dev = np.abs((e[0] - i[0]) - (e0 - i0)) / np.maximum(np.abs(e0), np.abs(i0))
        assert dev[ok].max() < D_TOL, (mode, dev[ok].max())


@pytest.mark.parametrize("kind", ["cylinder_density", "cylinder_flow", "slab_density"])
def test_normal_form_and_first_derivative_form_agree(kind):
    """ESB_RK8N (u = sqrt|F| y, Nystrom form) and ESB_RK8 ((y, h y') variables) are two discretisations of
    the same problem: they agree to the discretisation error wherever the problem is regular, and the
    scheme's own switch (near_resonance) keeps the points next to a continuum on the accurate form."""
    case = CASES[kind]
    k = np.linspace(0.01, 4.5, 16)
    W = np.linspace(case.W[0], case.W[1], 600)
    modes = list(case.modes)
    a = spec_of(case)
    assert a.scheme == "rk8n"
    b = spec_of(case, scheme="rk8")
    ea, ia, da = hk.grid(a, modes, k, W)
    eb, ib, db = hk.grid(b, modes, k, W)
    assert np.array_equal(np.isnan(ea), np.isnan(eb))
    for slot, mode in enumerate(modes):
        ok = case.regular(k, W, mode) & np.isfinite(ea[slot]) & np.isfinite(ia[slot])
        sc = np.maximum(np.abs(eb[slot]), np.abs(ib[slot]))
        assert np.max(np.abs(ea[slot] - eb[slot])[ok] / np.abs(eb[slot])[ok]) < 1e-13     # same exterior
        assert np.max((np.abs(ia[slot] - ib[slot]) / sc)[ok]) < 2 * D_TOL
        # the denominator Y (pole-free refinement) is the same quantity in both
        assert np.max((np.abs(da[slot] - db[slot]) / np.abs(db[slot]))[ok]) < 1e-7


def test_fused_equals_single_mode_on_host():
    case = CASES["cylinder_density"]
    k = np.linspace(0.3, 4.0, 7)
    W = np.linspace(2.95, 4.95, 50)
    sp = spec_of(case)
    e3, i3, d3 = hk.grid(sp, [0, 1, 2], k, W)
    for slot, m in enumerate((0, 1, 2)):
        e1, i1, d1 = hk.grid(sp, [m], k, W)
        assert np.allclose(e3[slot], e1[0], rtol=1e-13, equal_nan=True)
        assert np.allclose(i3[slot], i1[0], rtol=1e-11, equal_nan=True)


LEAKY_WINDOWS = {
    # phase speeds with m_e < 0: above max(c_e, vA_e) and inside (cT_e, min(c_e, vA_e))
    "cylinder_density": [(5.02, 8.0), (0.4976, 0.4999)], "slab_density": [(3.02, 5.0)],
    "cylinder_flow": [(5.02, 8.0)], "slab_flow": [(2.52, 4.0), (-4.0, -2.52)],
    "cylinder_rotation": [(1.52, 3.2), (0.475, 0.499)], "cylinder_photospheric": [(1.52, 3.2), (0.475, 0.499)],
    "slab_photospheric": [(1.32, 2.5)],
}


@pytest.mark.parametrize("name", list(LEAKY_WINDOWS))
def test_leaky_evaluation_matches_oracle_without_the_skip(name):
    """The opt-in leaky evaluation (eval_point<..., LEAKY>: closed-form J_n / Y_n or cos / sin exterior where
    m_e < 0, interior and matching unchanged) against the C oracle with its skip rule lifted - the oracle
    integrates the exterior numerically, so it needs no new code for that side.  This is what the reference's
    scan loop would evaluate without `if m_e < 0: pass` (Density_cylinder.py:760)."""
    case = CASES[name]
    model = case.c_model()
    model.leaky = 1
    sp = spec_of(case)
    rng = np.random.default_rng(5)
    W = np.concatenate([rng.uniform(a, b, 120 // len(LEAKY_WINDOWS[name])) for a, b in LEAKY_WINDOWS[name]])
    k = rng.uniform(0.3, 4.0, W.size)
    for mode in list(case.modes)[:3]:
        e, i, _ = hk.evaluate(sp, [mode], k, k * W, leaky=True)
        ref = np.array([ork.point(model, mode, a, a * b) for a, b in zip(k, W)])
        e0, i0 = ref[:, 0], ref[:, 1]
        assert np.isfinite(e0).all() and np.isfinite(e[0]).all()           # nothing is skipped
        reg = np.array([case.regular(a, np.array([b]), mode, margin=0.03)[0, 0] for a, b in zip(k, W)])
        assert reg.sum() >= 60
        assert np.max(np.abs(e[0] - e0)[reg] / np.abs(e0)[reg]) < 1e-9         # the closed-form exterior
        dev = np.abs((e[0] - i[0]) - (e0 - i0)) / np.maximum(np.abs(e0), np.abs(i0))
        assert np.quantile(dev[reg], 0.9) < 1e-10 and dev[reg].max() < 1e-7, (mode, dev[reg].max())
    # and the default evaluation still skips those points
    e, i, _ = hk.evaluate(sp, [case.modes[0]], k, k * W)
    assert np.isnan(e).all() and np.isnan(i).all()


def _fine(sp, factor=2):
    """what DispersionSolver._fine_spec uploads as the guard's model"""
    kw = sp.solver_kwargs()
    kw["n_steps"] *= factor
    if kw["scheme"] == "rk8n" and kw["n_steps"] > 680:
        kw["scheme"] = "rk8"
    return esb.ModelSpec(**kw)


def test_guard_measure_on_host():
    """The discretisation guard's own judgement (core.cuh guard_deviation behind resonance_free - the code the
    device's guard_kernel runs) on the host build: (i) every shipped equilibrium, every azimuthal order up to
    n = 3, is far below the 1e-9 threshold at its default step count; (ii) the measure is projective: for the
    fluting n = 3 order the numerator N = int Y and the denominator Y both carry a 2.5e-9 amplitude error of the
    solution that dominates towards the axis, which cancels in int = N / Y - D is converged to 1e-13 - and which
    G = D Y alone (the measure of ABI 1.3) reported as an error of the sweep; (iii) the sharp shell is seen,
    and cured by the step count the order predicts."""
    for name, case in CASES.items():
        k = np.linspace(0.25, 4.0, 12) if case.kind == "cylinder_rotation" else np.linspace(0.5, 4.5, 12)
        W = np.linspace(case.W[0], case.W[1], 120)
        sp = spec_of(case)
        modes = list(case.modes) + ([3] if case.kind.startswith("cylinder") else [])
        dev = hk.guard_grid(sp, _fine(sp), modes, k, W)
        for m, d in zip(modes, dev):
            if name == "cylinder_rotation_p08" and m >= 2:
                continue                                  # (almost) every point next to a resonance: nothing judged
            assert np.isfinite(d).sum() > 100 and np.nanmax(d) < 5e-11, (name, m, np.nanmax(d), np.isfinite(d).sum())
    case = CASES["cylinder_density"]
    sp = spec_of(case)
    k, W = np.linspace(0.5, 4.5, 12), np.linspace(3.0, 4.9, 60)
    e0, i0, d0 = hk.grid(sp, [3], k, W)
    e1, i1, d1 = hk.grid(_fine(sp), [3], k, W)
    ok = np.isfinite(e0) & np.isfinite(i0)
    common = (d0 / d1 - 1.0)[ok]
    assert 1e-9 < np.median(np.abs(common)) < 1e-8                          # Y: a visible factor, point by point ...
    assert np.median(np.abs((i0 * d0) / (i1 * d1) - 1.0 - (d0 / d1 - 1.0))[ok]) < 1e-12     # ... shared by N ...
    assert np.median(np.abs((e0 - i0) - (e1 - i1))[ok] / np.maximum(np.abs(e1), np.abs(i1))[ok]) < 1e-13   # ... not in D
    old = np.abs((e0 - i0) * d0 - (e1 - i1) * d1) / (np.abs(e1 * d1) + np.abs(i1 * d1))
    new = hk.guard_grid(sp, _fine(sp), [3], k, W)
    assert np.nanmedian(old[np.isfinite(new)]) > 1e-10 and np.nanmax(new) < 5e-11
    k, W2 = np.linspace(0.5, 4.5, 12), np.linspace(4.6, 4.95, 80)
    worst = {}
    for n in (None, 448):
        sp = esb.ModelSpec("cylinder_density", profile=esb.GaussianDensity(0.05, x0=-0.5), n_steps=n)
        worst[n] = np.nanmax(hk.guard_grid(sp, _fine(sp), [0, 1], k, W2))
    assert worst[None] > 1e-7 and worst[448] < 1e-9, worst


def test_max_steps_is_the_capacity_of_the_staged_table():
    """esb_model_max_steps (host code of the library): the largest step count whose staged table fits the
    200 KB of shared memory - the table of that many steps builds, one more step is refused."""
    expect = {("cylinder_density", "rk8n"): 710, ("cylinder_density", "rk8"): 1279, ("cylinder_rotation", "rk8"): 710,
              ("slab_density", "rk8n"): 710, ("slab_flow", "rk8"): 1278, ("cylinder_flow", "rk8n"): 710}
    for (kind, scheme), n_max in expect.items():
        sp = esb.ModelSpec(kind, scheme=scheme)
        assert sp.max_steps() == n_max, (kind, scheme, sp.max_steps())
        step = 2 if kind.startswith("slab") else 1
        kw = sp.solver_kwargs()
        mode = [1]
        kw["n_steps"] = n_max
        e, i, _ = hk.evaluate(esb.ModelSpec(**kw), mode, [1.0], [1.3 if kind == "cylinder_rotation" else 4.7])
        kw["n_steps"] = n_max + step
        with pytest.raises(RuntimeError):
            hk.evaluate(esb.ModelSpec(**kw), mode, [1.0], [4.7])
    assert esb.ModelSpec("cylinder_density").max_steps("rk8") == 1279


def test_host_build_on_the_scanned_equilibria():
    """The two families of BASELINE configs[4] (eigensolver_b200.scan.density_flow_grid: cylinder density models
    over the density contrast, slab flow models over the flow amplitude): the kernels' arithmetic against the C
    oracle at random (k, omega) outside each equilibrium's own continua, every mode."""
    from eigensolver_b200.scan import density_flow_grid
    from helpers import continua, cyl_profile, flow_continua, regular_mask
    from oracle import reference_path as rp
    dens, flow = density_flow_grid(np.linspace(0.05, 0.4, 5), np.linspace(0.1, 0.9, 5))
    rng = np.random.default_rng(5)
    for p in dens:
        md = p["medium"]
        sp = esb.ModelSpec("cylinder_density", medium=md, profile=p["profile"])
        model = ork.make_model("cylinder_density", medium=md, width=p["profile"].width)
        iv = continua(cyl_profile(p["profile"].width, rp.Medium(c_i0=md.c_i0, vA_i0=md.vA_i0, vA_e=md.vA_e, c_e=md.c_e)),
                      -1.0, -0.001, False)
        k = rng.uniform(0.05, 4.5, 150)
        W = rng.uniform(0.45, min(md.vA_e, 5.0), k.size)          # the phase speeds of the BASELINE window
        e, i, _ = hk.evaluate(sp, [0, 1, 2], k, k * W)
        for slot, mode in enumerate((0, 1, 2)):
            ref = np.array([ork.point(model, mode, kk, kk * ww) for kk, ww in zip(k, W)])
            fin = np.isfinite(ref[:, 0]) & np.isfinite(ref[:, 1])
            assert np.array_equal(np.isfinite(e[slot]) & np.isfinite(i[slot]), fin)
            ok = fin & regular_mask(W, iv, 0.02)
            assert ok.sum() > 30, (p["label"], mode, ok.sum())
            dev = np.abs((e[slot] - i[slot]) - (ref[:, 0] - ref[:, 1])) / np.maximum(np.abs(ref[:, 0]), np.abs(ref[:, 1]))
            # next to a pole of D (int = N / Y, Y -> 0) the error of Y is amplified: at most one such point
            assert np.quantile(dev[ok], 0.98) < 2e-10 and (dev[ok] > D_TOL).sum() <= 1 and dev[ok].max() < 2e-8, \
                (p["label"], mode, dev[ok].max())
    for p in flow:
        md = p["medium"]
        sp = esb.ModelSpec("slab_flow", medium=md, profile=p["profile"])
        rmd = rp.FlowMedium(width=p["profile"].width, U_i0=md.U_i0)
        model = ork.make_model("slab_flow", medium=rmd, width=p["profile"].width)
        iv = flow_continua(rmd)
        k = rng.uniform(0.05, 4.5, 150)
        W = rng.uniform(-2.7, 2.7, k.size)
        e, i, _ = hk.evaluate(sp, [0, 1], k, k * W)
        for slot, mode in enumerate((0, 1)):
            ref = np.array([ork.point(model, mode, kk, kk * ww) for kk, ww in zip(k, W)])
            fin = np.isfinite(ref[:, 0]) & np.isfinite(ref[:, 1])
            assert np.array_equal(np.isfinite(e[slot]) & np.isfinite(i[slot]), fin)
            ok = fin & regular_mask(W, iv, 0.02)
            assert ok.sum() > 30, (p["label"], mode, ok.sum())
            dev = np.abs((e[slot] - i[slot]) - (ref[:, 0] - ref[:, 1])) / np.maximum(np.abs(ref[:, 0]), np.abs(ref[:, 1]))
            # next to a pole of D (int = N / Y, Y -> 0) the error of Y is amplified: at most one such point
            assert np.quantile(dev[ok], 0.98) < 2e-10 and (dev[ok] > D_TOL).sum() <= 1 and dev[ok].max() < 2e-8, \
                (p["label"], mode, dev[ok].max())
