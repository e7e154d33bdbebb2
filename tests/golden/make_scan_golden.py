"""Generate tests/golden/ref_scans.npz: the reference's OWN scan + bisection output
(`sol_ks`, `sol_omegas` of sausage()/kink()) over many (script, mode, k, frequency window) cases,
for the bidirectional check of the drop-in's "reference" accept rule.

Runs ONLY in the build container (needs /root/reference):  python tests/golden/make_scan_golden.py
One process per case (the scripts keep module-global state); ~10 minutes on 8 cores.

Per case n the file holds  c<n>_script (key of reference_api.SCRIPTS), c<n>_overrides (repr of the
assignment-line edits applied to the script, mirrored by `DROPIN_OVERRIDES` in the test), c<n>_mode,
c<n>_k, c<n>_freq, c<n>_sol_ws.
"""
from __future__ import annotations

import multiprocessing as mp
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

#: script preset -> (run_reference solver, assignment-line overrides, [(mode, k, W_lo, W_hi, n_freq)])
CASES = {
    "cylinder_density": ("cylinder_density_coronal", {}, [
        ("kink", 0.6, 2.95, 4.9, 40), ("kink", 1.0, 2.95, 4.9, 30), ("kink", 2.0, 2.95, 4.9, 40),
        ("kink", 3.5, 2.95, 4.9, 40), ("sausage", 2.0, 2.95, 4.9, 30), ("sausage", 3.0, 2.95, 4.9, 40),
        ("sausage", 4.2, 2.95, 4.9, 40), ("kink", 3.0, 0.52, 0.88, 30), ("sausage", 3.5, 0.52, 0.88, 40),
        ("kink", 1.5, 4.0, 4.99, 50)]),
    "cylinder_density_photospheric": ("cylinder_density_photospheric", {}, [
        ("kink", 1.0, 0.9, 1.49, 40), ("kink", 2.5, 0.9, 1.49, 40), ("sausage", 1.5, 0.9, 1.49, 40),
        ("sausage", 3.5, 0.9, 1.49, 40), ("kink", 4.0, 0.55, 0.88, 40), ("sausage", 2.0, 0.55, 0.88, 40)]),
    "slab_density": ("slab_density_coronal", {}, [
        ("kink", 0.75, 0.42, 0.76, 25), ("sausage", 1.5, 1.75, 2.95, 25), ("kink", 1.5, 1.75, 2.95, 40),
        ("sausage", 0.6, 1.75, 2.95, 40), ("kink", 3.0, 1.75, 2.95, 40), ("sausage", 3.0, 0.42, 0.76, 40)]),
    "slab_density_photospheric": ("slab_density_photospheric", {"dx": 0.9}, [
        ("kink", 1.0, 1.02, 1.29, 35), ("sausage", 1.0, 1.02, 1.29, 35), ("kink", 2.5, 1.02, 1.29, 35),
        ("sausage", 2.5, 0.3, 0.7, 35)]),
    "slab_flow": ("slab_flow_coronal", {"dx": 1.0}, [
        ("kink", 1.5, 1.25, 2.45, 30), ("sausage", 1.5, 1.25, 2.45, 30), ("kink", 1.5, -2.45, -1.25, 30),
        ("sausage", 3.0, 1.25, 2.45, 40), ("kink", 0.6, 1.25, 2.45, 40), ("sausage", 3.0, -2.45, -1.25, 40)]),
    "slab_flow_photospheric": ("slab_flow_photospheric", {}, [
        ("kink", 1.0, 0.2, 0.58, 40), ("sausage", 1.0, 0.2, 0.58, 40), ("kink", 2.5, 0.2, 0.58, 40),
        ("sausage", 2.5, -0.85, -0.3, 40)]),
    "cylinder_flow": ("cylinder_flow_coronal", {"U_i0": 0.35, "dr": 1.0}, [
        ("kink", 3.0, 2.95, 4.9, 30), ("sausage", 2.0, 2.95, 4.9, 30), ("sausage", 3.0, -4.9, -2.95, 30),
        ("kink", 3.0, -4.9, -2.95, 30), ("kink", 1.0, 2.95, 4.9, 40), ("sausage", 4.0, 2.95, 4.9, 40)]),
    "rotation_sausage": ("cylinder_rotation_sausage", {}, [
        ("sausage", 1.5, 0.9, 1.49, 40), ("sausage", 2.5, 0.9, 1.49, 40), ("sausage", 3.5, 0.9, 1.49, 40)]),
    "rotation_kink": ("cylinder_rotation_kink", {"v_twist": 0.15, "power": 1.25}, [
        ("kink", 1.0, 0.9, 1.49, 40), ("kink", 2.5, 0.9, 1.49, 40), ("kink", 3.5, 0.9, 1.49, 40)]),
    # the two kink scripts exactly as shipped (v_phi = 0.25 r^0.8 / 0.1 r^0.8, layer down to r = 0.001), on
    # windows where no resonance sits next to the axis (tests/helpers.py rotation_regular); "@" = variant tag
    "rotation_kink@p08": ("cylinder_rotation_kink", {}, [
        ("kink", 3.0, 1.25, 1.45, 40), ("kink", 3.5, 1.25, 1.45, 40), ("kink", 4.0, 1.22, 1.45, 40)]),
    "rotation_kink_slow@p08": ("cylinder_rotation_kink_slow", {}, [
        ("kink", 2.0, 1.12, 1.45, 40), ("kink", 3.0, 1.06, 1.45, 40), ("kink", 4.0, 1.06, 1.45, 40)]),
}


def run_case(args):
    script, solver, overrides, (mode, k, lo, hi, num) = args
    from run_reference import ReferenceSolver
    t0 = time.time()
    ref = ReferenceSolver(solver, overrides=overrides or None)
    freq = np.linspace(lo * k, hi * k, num)
    try:
        ks, ws = ref.roots(mode, k, freq)
        ws = np.real(np.asarray(ws)).astype(np.float64)
    except RecursionError:          # the scripts' recursion has no other bound than itt_num > 150
        ws = np.array([np.nan])
    return dict(script=script, overrides=repr(overrides), mode=0 if mode == "sausage" else 1, k=k, freq=freq,
                sol_ws=ws, seconds=time.time() - t0)


def main():
    jobs = [(script, solver, ov, case) for script, (solver, ov, cases) in CASES.items() for case in cases]
    only = sys.argv[1:]
    out, first = {}, 0
    if only:
        # append the named scripts' cases to the existing fixture instead of regenerating all of it
        jobs = [j for j in jobs if j[0] in only]
        old = np.load(os.path.join(HERE, "ref_scans.npz"))
        out = {f: old[f] for f in old.files}
        while "c%d_script" % first in out:
            first += 1
    with mp.get_context("fork").Pool(min(8, os.cpu_count() or 1), maxtasksperchild=1) as pool:
        for n, r in enumerate(pool.imap(run_case, jobs), start=first):
            for key in ("script", "overrides"):
                out["c%d_%s" % (n, key)] = np.array(r[key])
            out["c%d_mode" % n] = np.array([r["mode"]])
            out["c%d_k" % n] = np.array([r["k"]])
            out["c%d_freq" % n] = r["freq"]
            out["c%d_sol_ws" % n] = r["sol_ws"]
            print(n, r["script"], r["mode"], r["k"], "->", r["sol_ws"], "(%.0f s)" % r["seconds"], flush=True)
    np.savez(os.path.join(HERE, "ref_scans.npz"), **out)


if __name__ == "__main__":
    main()
