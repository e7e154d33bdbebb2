"""Generate the golden fixtures in tests/golden/ from the reference itself.

Runs ONLY in the build container (needs /root/reference):

    python tests/golden/make_golden.py

Produces
  ref_D_<solver>.npz      D(omega,k) values obtained by executing the reference's own
                          sausage()/kink() functions (see run_reference.py) at a set of
                          (mode, k, omega) points: arrays mode, k, w, D (nan = skipped).
  ref_scan_<solver>.npz   the reference's own scan + bisection (its `sol_ks`, `sol_omegas`)
                          over a few (k, frequency-interval) pairs.
  ref_roots.npz           the root tables the reference ships as "Example data/*.pickle"
                          ([sausage w, sausage k, kink w, kink k]) for the density solvers,
                          keyed <family>_<width>_<mode>_{k,w}.
"""
from __future__ import annotations

import glob
import os
import pickle
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from run_reference import REF_ROOT, ReferenceSolver  # noqa: E402

# phase speeds inside the windows where no Alfven/cusp/sound resonance sits in the layer
POINTS = {
    "cylinder_density_coronal": dict(
        ks=[0.1, 0.5, 1.0, 2.0, 3.2, 4.5],
        Ws=[0.52, 0.6, 0.75, 0.88, 1.35, 1.6, 1.95, 2.95, 3.3, 4.0, 4.9, 0.45, 5.2],  # last two: skipped / leaky
    ),
    "slab_density_coronal": dict(
        ks=[0.05, 0.3, 0.75, 1.5, 3.0],
        Ws=[0.42, 0.5, 0.6, 0.74, 1.75, 2.0, 2.5, 2.95, 0.398, 3.1],
    ),
}

POINTS.update({
    # (overrides applied to the script's own assignment lines)
    "cylinder_density_photospheric": dict(
        ks=[0.1, 0.6, 1.5, 3.0, 4.5], Ws=[0.55, 0.65, 0.95, 1.2, 1.45, 0.49, 1.6], overrides={}),
    "slab_density_photospheric": dict(
        ks=[0.1, 0.6, 1.5, 3.0], Ws=[0.3, 0.5, 0.65, 1.05, 1.2, 1.28, 0.75, 1.4], overrides={"dx": 0.9}),
    "slab_flow_coronal": dict(
        ks=[0.1, 0.6, 1.5, 3.0, 4.5], Ws=[1.3, 1.6, 2.0, 2.4, -0.3, -1.0, -2.0, -0.1, 2.6, 0.1995],
        overrides={"dx": 1.0}),
    "slab_flow_photospheric": dict(
        ks=[0.1, 0.6, 1.5, 3.0], Ws=[0.2, 0.3, 0.45, 0.5, 0.58, -0.3, -0.5, -0.7, -0.85, 0.65, -0.95], overrides={}),
    # cylinder with an axial flow: the script ships with U_i0 = 0, dr = 1e5 (no flow at all)
    "cylinder_flow_coronal": dict(
        ks=[0.1, 0.6, 1.5, 3.0, 4.0],
        Ws=[0.52, 0.75, 1.5, 2.7, 3.3, 4.0, 4.9, -0.6, -1.6, -3.0, -4.5, 0.45, 5.2],
        overrides={"U_i0": 0.35, "dr": 1.0}),
    # rotational flow in its regular regime (power >= 1); the kink script is run with the sausage
    # script's v_twist/power instead of its own (0.25, 0.8), see tests/helpers.py
    "cylinder_rotation_sausage": dict(
        ks=[0.5, 1.0, 2.0, 3.0, 4.0], Ws=[0.55, 0.8, 0.97, 1.1, 1.2, 1.3, 1.45, 0.49, 1.6], overrides={}),
    "cylinder_rotation_kink": dict(
        ks=[0.5, 1.0, 2.0, 3.0, 4.0], Ws=[0.55, 0.8, 0.97, 1.1, 1.2, 1.3, 1.45, 0.49, 1.6],
        overrides={"v_twist": 0.15, "power": 1.25}),
})

# the Epstein profile the cylinder script carries as a comment (:139-142), switched on the way its author would
EPSTEIN_PATCH = [(
    "def profile(r):       # Define the internal profile as a function of variable x   (Inverted Gaussian)\n"
    "    return (rho_e + ((rho_i0 - rho_e)*sym.exp(-(r-r0)**2/dr**2)))",
    "a = 1.   #inhomogeneity width for epstein profile\n"
    "def profile(x):       # Define the internal profile as a function of variable x   (Epstein Profile)\n"
    "    return ((rho_i0 - rho_e)/(sym.cosh(x/a)**4)**2 + rho_e)")]
POINTS["cylinder_density_epstein"] = dict(
    solver="cylinder_density_coronal", patches=EPSTEIN_PATCH,
    ks=[0.1, 0.5, 1.0, 2.0, 3.2, 4.5], Ws=[0.52, 0.6, 0.75, 0.88, 1.35, 1.6, 1.95, 2.95, 3.3, 4.0, 4.9, 0.45, 5.2])

# a second width per density script (the widths of the shipped root tables)
POINTS["cylinder_density_coronal_w15"] = dict(
    solver="cylinder_density_coronal", overrides={"dr": 1.5},
    ks=[0.3, 1.0, 2.0, 3.2, 4.5], Ws=[0.52, 0.7, 0.88, 1.5, 1.95, 3.3, 4.0, 4.9, 0.45, 5.2])
POINTS["slab_density_coronal_w3"] = dict(
    solver="slab_density_coronal", overrides={"dx": 3.0},
    ks=[0.05, 0.3, 0.75, 1.5, 3.0], Ws=[0.42, 0.6, 0.74, 1.75, 2.0, 2.5, 2.95, 0.398, 3.1])

# rotational flow, a second (linear) rotation law: the shipped root tables of that law are checked too
POINTS["cylinder_rotation_sausage_p1"] = dict(
    solver="cylinder_rotation_sausage", overrides={"v_twist": 0.1, "power": 1.0},
    ks=[0.5, 1.0, 2.0, 3.0, 4.0], Ws=[0.55, 0.8, 0.97, 1.1, 1.2, 1.3, 1.45, 0.49, 1.6])
POINTS["cylinder_rotation_kink_p1"] = dict(
    solver="cylinder_rotation_kink", overrides={"v_twist": 0.1, "power": 1.0},
    ks=[0.5, 1.0, 2.0, 3.0, 4.0], Ws=[0.55, 0.8, 0.97, 1.1, 1.2, 1.3, 1.45, 0.49, 1.6])

# the kink scripts at their OWN shipped rotation laws (power 0.8: the Doppler shift m v_phi/r grows towards
# the axis, so part of every scan crosses a resonance near the axis; "regular_only" keeps the candidate
# points at which neither D nor C3 changes sign inside the layer - tests/helpers.py::rotation_regular -
# plus the skipped ones, because the reference needs minutes per point across a resonance)
POINTS["cylinder_rotation_kink_p08"] = dict(
    solver="cylinder_rotation_kink", overrides={}, regular_only=(0.25, 0.8, 0.001),
    ks=[0.25, 0.31, 0.37, 0.6, 1.0, 1.5, 2.0, 2.5, 3.0, 3.5, 4.0],
    Ws=[0.97, 1.05, 1.1, 1.2, 1.28, 1.33, 1.37, 1.39, 1.45, 0.49, 1.6])
# the sausage script with the same law (its own layer end, r = 0.01)
POINTS["cylinder_rotation_sausage_p08"] = dict(
    solver="cylinder_rotation_sausage", overrides={"v_twist": 0.25, "power": 0.8}, regular_only=(0.25, 0.8, 0.01),
    ks=[0.5, 1.0, 2.0, 3.0, 4.0], Ws=[0.55, 0.8, 0.97, 1.1, 1.2, 1.3, 1.45, 0.49, 1.6])
POINTS["cylinder_rotation_kink_slow_p08"] = dict(
    solver="cylinder_rotation_kink_slow", overrides={}, regular_only=(0.1, 0.8, 0.001),
    ks=[0.05, 0.2, 0.5, 1.0, 2.0, 3.0, 4.0], Ws=[0.95, 1.02, 1.08, 1.15, 1.19, 1.3, 1.45, 0.49, 1.6])

SCANS = {
    # (mode, k, W_lo, W_hi, n)  - intervals that contain a mode
    "cylinder_density_coronal": [("kink", 1.0, 2.95, 4.9, 30), ("sausage", 2.0, 2.95, 4.9, 30),
                                 ("kink", 3.0, 0.52, 0.88, 30)],
    "slab_density_coronal": [("kink", 0.75, 0.42, 0.76, 25), ("sausage", 1.5, 1.75, 2.95, 25)],
    "cylinder_flow_coronal": [("kink", 3.0, 2.95, 4.9, 30), ("sausage", 2.0, 2.95, 4.9, 30),
                              ("sausage", 3.0, -4.9, -2.95, 30), ("kink", 3.0, -4.9, -2.95, 30)],
    "slab_flow_coronal": [("kink", 1.5, 1.25, 2.45, 30), ("sausage", 1.5, 1.25, 2.45, 30),
                          ("kink", 1.5, -2.45, -1.25, 30)],
}

PICKLES = {
    "cyl_coronal": ("Cylinder/Non-uniform density/Coronal/Example data/Cylindrical_coronal_width%s.pickle",
                    {"09": 0.9, "1": 1.0, "125": 1.25, "15": 1.5, "175": 1.75, "3": 3.0, "1e5": 1e5}),
    "slab_coronal": ("Slab/Non uniform density/Coronal/Example data/width%s_coronal.pickle",
                     {"09": 0.9, "15": 1.5, "3": 3.0, "1e5": 1e5}),
    "cyl_photospheric": ("Cylinder/Non-uniform density/Photospheric/Example data/"
                         "Cylindrical_photospheric_width_%s.pickle", {"09": 0.9, "15": 1.5, "3": 3.0, "1e5": 1e5}),
    "slab_photospheric": ("Slab/Non uniform density/Photospheric/Example data/width%s.pickle",
                          {"09": 0.9, "15": 1.5, "3": 3.0, "1e5": 1e5}),
    "cylflow_coronal": ("Cylinder/Non-uniform flow/Coronal/Example data/Cylindrical_coronal_flow_%s.pickle",
                        # produced with U_i0 = 0.05 and xi_tol = 6 % (tests/helpers.py ROOT_CASES)
                        {"1e5": 1e5, "1": 1.0, "06": 0.6}),
    "flow_coronal": ("Slab/Non uniform flow/Example data/flow_width%s_coronal.pickle",
                     # produced with U_i0 = 0.35 (tests/helpers.py ROOT_CASES); the file named
                     # "width125" matches width 2.5 (median mismatch 0.6 %), not 1.25 (15 %)
                     {"1": 1.0, "125": 2.5, "15": 1.5, "3": 3.0, "5": 5.0, "1e5": 1e5}),
}


def _rotation_point_ok(law, mode, k, W, margin=0.03):
    """skipped by the reference (m_e < 0), or regular: no resonance inside the layer at W and W +- margin"""
    sys.path.insert(0, os.path.dirname(HERE))
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    import helpers
    from oracle import reference_path as rp
    md = rp.CYL_PHOTOSPHERIC
    if md.m_e(k, W * k) < 0:
        return True
    vt, pw, s_end = law
    ok = helpers.regular_mask(np.array([W]), helpers.rotation_continua(md, vt, pw, s_end, mode, k), margin)[0]
    for dW in (-margin, 0.0, margin):
        ok = ok and helpers.rotation_regular(md, vt, pw, s_end, mode, k, np.array([W + dW]))[0]
    return bool(ok)


def main():
    only = sys.argv[1:]
    for name, spec in POINTS.items():
        if only and name not in only:
            continue
        t0 = time.time()
        ref = ReferenceSolver(spec.get("solver", name), overrides=spec.get("overrides"), patches=spec.get("patches"))
        rows = []
        for mode_id, mode in ((0, "sausage"), (1, "kink")):
            if mode not in ref.modes:
                continue
            for k in spec["ks"]:
                for W in spec["Ws"]:
                    if spec.get("regular_only") and not _rotation_point_ok(spec["regular_only"], mode_id, k, W):
                        continue
                    rows.append((mode_id, k, W * k, ref.D(mode, k, W * k)))
        a = np.array(rows)
        np.savez(os.path.join(HERE, "ref_D_%s.npz" % name), mode=a[:, 0].astype(np.int32), k=a[:, 1],
                 w=a[:, 2], D=a[:, 3])
        print(name, len(rows), "points in %.0f s" % (time.time() - t0), flush=True)
        out = {}
        for n, (mode, k, lo, hi, num) in enumerate(SCANS.get(name, [])):
            freq = np.linspace(lo * k, hi * k, num)
            ks, ws = ref.roots(mode, k, freq)
            out["scan%d_mode" % n] = np.array([0 if mode == "sausage" else 1])
            out["scan%d_k" % n] = np.array([k])
            out["scan%d_freq" % n] = freq
            out["scan%d_sol_ks" % n] = ks
            out["scan%d_sol_ws" % n] = ws
            print(name, mode, k, "->", ws, flush=True)
        if out:
            np.savez(os.path.join(HERE, "ref_scan_%s.npz" % name), **out)

    if only and "roots" not in only:
        return
    roots = {}
    for fam, (pat, widths) in PICKLES.items():
        for tag, width in widths.items():
            with open(os.path.join(REF_ROOT, pat % tag), "rb") as fh:
                sw, sk, kw, kk = pickle.load(fh, encoding="latin1")
            key = "%s_%s" % (fam, tag)
            roots[key + "_width"] = np.array([width])
            roots[key + "_sausage_w"] = np.real(np.asarray(sw)).astype(np.float64)
            roots[key + "_sausage_k"] = np.real(np.asarray(sk)).astype(np.float64)
            roots[key + "_kink_w"] = np.real(np.asarray(kw)).astype(np.float64)
            roots[key + "_kink_k"] = np.real(np.asarray(kk)).astype(np.float64)
    # rotational flow: [omega, k] per file, every shipped table (54: four rotation amplitudes x powers
    # 0.8, 0.9, 1, 1.25 x the sausage / kink, fast / slow scripts); key rot_v<amp>_p<power>_<file suffix>
    rot_dir = os.path.join(REF_ROOT, "Cylinder/Rotational flow/Photospheric/Example data")
    for path in sorted(glob.glob(os.path.join(rot_dir, "Cylindrical_photospheric_vtwist*_power*_*.pickle"))):
        stem = os.path.basename(path)[len("Cylindrical_photospheric_vtwist"):-len(".pickle")]
        vt, rest = stem.split("_power", 1)
        pw, kind = rest.split("_", 1)
        with open(path, "rb") as fh:
            w, k = pickle.load(fh, encoding="latin1")
        key = "rot_v%s_p%s_%s" % (vt, pw, kind)
        roots[key + "_w"] = np.real(np.asarray(w)).astype(np.float64)
        roots[key + "_k"] = np.real(np.asarray(k)).astype(np.float64)
    np.savez(os.path.join(HERE, "ref_roots.npz"), **roots)
    print("ref_roots.npz:", len(roots), "arrays")


if __name__ == "__main__":
    main()
