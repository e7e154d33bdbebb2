"""Drop-in mirror of the reference scripts' solver functions.

The reference exposes, per solver script, two worker functions with the signature

    sausage(wavenumber, sausage_ws, sausage_ks, freq)
    kink(wavenumber, kink_ws, kink_ks, freq)

(Density_cylinder.py:847 and :546; ..._coronal.py:154 and :331) - one wavenumber,
an array of trial frequencies, and two queues that receive the list of solution
k's and the list of solution omega's - plus a driver that loops them over
`wavenumber x speed intervals` in separate processes and pickles
`[sol_omegas1, sol_ks1, sol_omegas_kink1, sol_ks_kink1]` (:1126-1183).

`ReferenceScript` keeps those names, argument meanings and the output format; the
work goes to the GPU through `DispersionSolver`.  `run()` is the driver loop done
as ONE batched GPU sweep per speed interval instead of one process per (k, interval).

`SCRIPTS` lists every solver script of the reference with the parameter values it
ships with (equilibrium, profile, layer end, exterior length, acceptance threshold,
driver grid); `ReferenceScript(name, **overrides)` is "that script with these
assignment lines edited", which is how the reference is used.
"""
from __future__ import annotations

import numpy as np

from .solver import (CYLINDER_CORONAL, CYLINDER_FLOW_CORONAL, CYLINDER_PHOTOSPHERIC, SLAB_CORONAL,
                     SLAB_FLOW_CORONAL, SLAB_PHOTOSPHERIC, DispersionSolver, FlowMedium, GaussianAxialFlow,
                     GaussianDensity, GaussianFlow, PowerLawRotation)

_REF = {
    "cyl_c": "Cylinder/Non-uniform density/Coronal/solvers/Density_cylinder.py",
    "cyl_p": "Cylinder/Non-uniform density/Photospheric/Solvers/Density_cylinder_photospheric.py",
    "slab_c": "Slab/Non uniform density/Coronal/Solvers/multiprocessor_Inhomogeneous_method_coronal.py",
    "slab_p": "Slab/Non uniform density/Photospheric/Solvers/multiprocessor_Inhomogeneous_method.py",
    "flow_c": "Slab/Non uniform flow/Solver/flow_multiprocessor_coronal.py",
    "flow_p": "Slab/Non uniform flow/Solver/flow_multiprocessor.py",
    "cylflow": "Cylinder/Non-uniform flow/Coronal/solvers/Cylinder_method_flow_testing.py",
    "rot_s": "Cylinder/Rotational flow/Photospheric/Solvers/Twisted_photospheric_flow_sausage.py",
    "rot_ss": "Cylinder/Rotational flow/Photospheric/Solvers/Twisted_photospheric_flow_sausage_slow.py",
    "rot_k": "Cylinder/Rotational flow/Photospheric/Solvers/Twisted_photospheric_nonlinear_flow_kink_fast.py",
    "rot_ks": "Cylinder/Rotational flow/Photospheric/Solvers/Twisted_photospheric_nonlinear_flow_kink_slow.py",
}


def _cT_boundary(md, prof):
    """cT_bound of the slab density script (:196): the tube speed at x = -1."""
    rho_b = float(np.asarray(prof(md, np.array([-1.0]))[0])[0])
    beta = md.vA_i0**2 * md.rho_i0
    alpha = md.rho_e * (md.c_e**2 + 0.5 * md.gamma * md.vA_e**2) - 0.5 * md.gamma * beta
    return float(np.sqrt(alpha * beta / ((alpha + beta) * rho_b)))


#: name -> the script's shipped settings.  tol = its acceptance threshold in percent (xi_tol /
#: p_tol / P_tol); accept = what the mismatch is divided by ("max": max(|ext|,|int|), "ext": |ext|);
#: speeds / n_freq / wavenumber = its driver loop (file:line in the comments).
SCRIPTS = {
    # Density_cylinder.py: dr :125, xi_tol :522, speeds :225 (including its `cT_e -c_e` element, a
    # missing comma), wavenumber :1126, test_freq :1145
    "cylinder_density": dict(
        ref=_REF["cyl_c"], kind="cylinder_density", medium=CYLINDER_CORONAL, profile=GaussianDensity(0.95),
        tol=1.0, n_freq=90, wavenumber=(0.01, 4.5, 90),
        speeds=lambda md, prof: sorted([md.c_i0, md.c_e, md.vA_i0, md.vA_e, md.cT_i0, md.cT_e - md.c_e, -md.c_i0,
                                  -md.vA_i0, -md.vA_e, -md.cT_i0, -md.cT_e])),
    # Density_cylinder_photospheric.py: written in r > 0; dr :125, speeds :227, driver :1129,:1148
    "cylinder_density_photospheric": dict(
        ref=_REF["cyl_p"], kind="cylinder_density", medium=CYLINDER_PHOTOSPHERIC, profile=GaussianDensity(0.9),
        solver=dict(coordinate="positive"), tol=1.0, n_freq=100, wavenumber=(0.01, 4.5, 130),
        speeds=lambda md, prof: sorted([md.c_i0, md.cT_i0, 0.5 * (md.c_i0 + md.cT_i0), 0.675, 0.8, 0.7])),
    # ..._method_coronal.py: dx :111, p_tol :378, speeds :202 (cT_bound = cT at the slab boundary), driver :876,:895
    "slab_density": dict(
        ref=_REF["slab_c"], kind="slab_density", medium=SLAB_CORONAL, profile=GaussianDensity(0.9),
        tol=1.0, n_freq=25, wavenumber=(0.001, 0.75, 25),
        speeds=lambda md, prof: sorted([1.0, _cT_boundary(md, prof), 1.2, 1.3, 0.9])),
    # ..._method.py (photospheric): dx :93 (uniform), 7-wavelength exterior, p_tol :275, speeds :182
    "slab_density_photospheric": dict(
        ref=_REF["slab_p"], kind="slab_density", medium=SLAB_PHOTOSPHERIC, profile=GaussianDensity(1e5),
        solver=dict(ext_wavelengths=7.0), tol=3.0, n_freq=35, wavenumber=(0.001, 0.75, 35),
        speeds=lambda md, prof: sorted([md.c_i0, md.cT_i0])),
    # flow_multiprocessor_coronal.py: speeds :180, p_tol :250, dx :119, driver :758,:780
    "slab_flow": dict(
        ref=_REF["flow_c"], kind="slab_flow", medium=SLAB_FLOW_CORONAL, profile=GaussianFlow(1e5),
        tol=1.0, n_freq=100, wavenumber=(0.01, 4.5, 100),
        speeds=lambda md, prof: sorted([-md.vA_e, 0.0, md.c_i, md.c_e, md.vA_i, md.vA_e, md.cT_i, md.cT_e])),
    # flow_multiprocessor.py (photospheric, steady flow): :63-69 speeds, 7 wavelengths :538, p_tol :411,
    # driver :812-838: freq = logspace(0.001, 0.55, 80) - 1 for every k, plus body_freq = linspace(cT_i k, (c_e+U_e) k, 100)
    "slab_flow_photospheric": dict(
        ref=_REF["flow_p"], kind="slab_flow",
        medium=FlowMedium(vA_i=1.0, c_i=2.0 / 3.0, vA_e=0.0, c_e=0.75, U_i0=0.0, U_e=-0.15),
        profile=GaussianFlow(1e5), solver=dict(ext_wavelengths=7.0), tol=1e-6, n_freq=100,
        wavenumber=(0.01, 3.5, 350), speeds=lambda md, prof: sorted([md.cT_i, md.c_e + md.U_e]),
        freq=lambda: np.logspace(0.001, 0.55, 80) - 1.0),       # :813, the same frequencies for every k
    # Cylinder_method_flow_testing.py: xi_tol :530 (6 %), speeds :243, driver :1134,:1153; ships with U_i0 = 0
    "cylinder_flow": dict(
        ref=_REF["cylflow"], kind="cylinder_flow", medium=CYLINDER_FLOW_CORONAL, profile=GaussianAxialFlow(1e5),
        tol=6.0, n_freq=70, wavenumber=(0.01, 4.0, 150),
        speeds=lambda md, prof: sorted([md.c_i0, md.vA_i0, md.vA_e, md.cT_i0, md.c_kink])),
    # Twisted_photospheric_flow_sausage.py: v_twist/power :176-177, ix ends at 0.01, xi_tol :419, speeds :224
    "rotation_sausage": dict(
        ref=_REF["rot_s"], kind="cylinder_rotation", medium=CYLINDER_PHOTOSPHERIC,
        profile=PowerLawRotation(0.15, 1.25), solver=dict(s_end=0.01), tol=1.5, n_freq=40,
        wavenumber=(0.75, 4.0, 110), speeds=lambda md, prof: sorted([md.c_e, md.c_kink, 1.4])),
    # ..._sausage_slow.py: xi_tol :423 = 4.5, speeds :232, wavenumber :752
    "rotation_sausage_slow": dict(
        ref=_REF["rot_ss"], kind="cylinder_rotation", medium=CYLINDER_PHOTOSPHERIC,
        profile=PowerLawRotation(0.15, 1.25), solver=dict(s_end=0.01), tol=4.5, n_freq=40,
        wavenumber=(0.25, 4.0, 100), speeds=lambda md, prof: sorted([1.0, 0.98, 0.96, 0.94, 0.92, 0.9, 0.88])),
    # ..._kink_fast.py: v_twist/power :176-177, P_tol :435, speeds :227, driver :743,:758
    "rotation_kink": dict(
        ref=_REF["rot_k"], kind="cylinder_rotation", medium=CYLINDER_PHOTOSPHERIC,
        profile=PowerLawRotation(0.25, 0.8), tol=2.5, n_freq=50, wavenumber=(0.25, 0.37, 20),
        speeds=lambda md, prof: sorted([md.c_kink, 1.35, 1.4])),
    # ..._kink_slow.py: P_tol :441 = 3, acceptance divided by |xi_e| alone (:586,:722), speeds :229
    "rotation_kink_slow": dict(
        ref=_REF["rot_ks"], kind="cylinder_rotation", medium=CYLINDER_PHOTOSPHERIC,
        profile=PowerLawRotation(0.1, 0.8), tol=3.0, accept="ext", n_freq=60, wavenumber=(0.01, 0.5, 60),
        speeds=lambda md, prof: sorted([md.c_i0, md.c_kink, 1.1, 1.2])),
}


def pickle_layout(script):
    """Which of run()'s four arrays a script pickles, in order (`pickle.dump([...], f)` at the end of every
    solver script): the density / flow scripts all four (Density_cylinder.py:1183), the rotational sausage
    scripts `[sol_omegas1, sol_ks1]` (Twisted_photospheric_flow_sausage.py:786), the rotational kink scripts
    `[sol_omegas_kink1, sol_ks_kink1]` (..._kink_fast.py:782)."""
    if script not in SCRIPTS:
        raise KeyError("unknown solver script %r" % script)
    if script.startswith("rotation_sausage"):
        return (0, 1)
    if script.startswith("rotation_kink"):
        return (2, 3)
    return (0, 1, 2, 3)


def write_root_table(path, results, script="cylinder_density"):
    """The reference's output file: `results` = run()'s [sol_omegas1, sol_ks1, sol_omegas_kink1, sol_ks_kink1]
    pickled as numpy arrays in the layout THAT script writes (pickle_layout), so that the analysis scripts
    which read those files (e.g. Eigenfunctions/analysis_compare_coronal_eigenfunctions_coronal.py:364
    `sol_omegas, sol_ks, sol_omegas_kink, sol_ks_kink = pickle.load(f)`) work on them unchanged."""
    import pickle
    if len(results) < 4:
        raise ValueError("results = [sol_omegas1, sol_ks1, sol_omegas_kink1, sol_ks_kink1]")
    payload = [np.asarray(results[i], dtype=np.float64) for i in pickle_layout(script)]
    with open(path, "wb") as fh:
        pickle.dump(payload, fh)
    return payload


def read_root_table(path, script="cylinder_density"):
    """Inverse of write_root_table; also reads the reference's own Example data pickles (written by
    Python 2 numpy: latin1).  Returns [sol_omegas1, sol_ks1, sol_omegas_kink1, sol_ks_kink1], empty
    arrays for the lists the script does not pickle."""
    import pickle
    with open(path, "rb") as fh:
        data = pickle.load(fh, encoding="latin1")
    layout = pickle_layout(script)
    if len(data) != len(layout):
        raise ValueError("%s holds %d lists, script %r pickles %d" % (path, len(data), script, len(layout)))
    out = [np.zeros(0) for _ in range(4)]
    for i, a in zip(layout, data):
        out[i] = np.real(np.asarray(a)).astype(np.float64)
    return out


class _ListSink:
    """Minimal stand-in for multiprocessing.Queue when the caller has none."""

    def __init__(self):
        self.items = []

    def put(self, x):
        self.items.append(x)


class ReferenceScript:
    """One reference solver script: its equilibrium, profile and settings, on the GPU.

    `script` is a key of SCRIPTS (for the two density kinds the plain kind name, as before).
    Overrides: medium=, profile=, width= (Gaussian profiles), tol=, rule= ("reference" | "converged"),
    and any DispersionSolver keyword (n_steps, scheme, coordinate, ext_wavelengths, s_end, device ...)."""

    def __init__(self, script="cylinder_density", medium=None, width=None, x0=0.0, tol=None, device=0,
                 profile=None, accept=None, kind=None, rule="reference", **solver_kw):
        if kind is not None:            # older spelling: ReferenceScript(kind="slab_density")
            script = kind
        if script not in SCRIPTS:
            raise ValueError("unknown reference script %r (have: %s)" % (script, ", ".join(SCRIPTS)))
        spec = SCRIPTS[script]
        self.script = script
        self.spec = spec
        self.kind = spec["kind"]
        self.medium = medium if medium is not None else spec["medium"]
        if profile is None:
            profile = spec["profile"]
            if width is not None:
                if not hasattr(profile, "width"):
                    raise ValueError("width= applies to the Gaussian profiles")
                profile = type(profile)(width, x0)
        self.tol = spec["tol"] if tol is None else tol       # xi_tol / p_tol / P_tol (percent)
        self.accept = accept or spec.get("accept", "max")
        kw = dict(spec.get("solver", {}))
        kw.update(solver_kw)
        self.solver = DispersionSolver(self.kind, self.medium, profile, device=device, **kw)
        # rule="reference": the script's own scan / bisection rule, point for point (the point sets its
        # pickles hold); "converged": one machine-precision root per sign change
        self.rule = rule
        if rule == "reference" and self.kind.startswith("slab"):
            self.rule = "reference_slab"   # bisects after two points seen, follows both halves
        if self.accept == "ext" and rule == "reference":
            # ..._kink_slow.py divides the mismatch by |xi_e| alone (:586): the library's reference rule uses
            # max(|ext|, |int|); that script is served by the converged rule + its own test in _modes_of
            self.rule = "converged"
        self.solver.set_accept_rule(self.rule)

    # -- the reference's worker signature -------------------------------
    def _modes_of(self, tab):
        """(k, omega) of the entries the script would have appended to sol_ks / sol_omegas."""
        if self.accept == "ext":
            # ..._kink_slow.py:586  |xi_e - xi_i| 100/|xi_e| < P_tol
            pct = np.abs(tab.ext - tab.intq) * 100.0 / np.abs(tab.ext)
            m = pct < self.tol
            return tab.k[m], tab.omega[m]
        return tab.modes()

    def _worker(self, mode, wavenumber, ws_queue, ks_queue, freq):
        freq = np.asarray(freq, dtype=np.float64)
        tab = self.solver.find_roots(mode, [float(wavenumber)], freq, layout="shared",
                                     tol_percent=self.tol)
        ks, ws = self._modes_of(tab)
        ks_queue.put(list(ks))
        ws_queue.put(list(ws))

    def sausage(self, wavenumber, sausage_ws, sausage_ks, freq):
        self._worker(0, wavenumber, sausage_ws, sausage_ks, freq)

    def kink(self, wavenumber, kink_ws, kink_ks, freq):
        self._worker(1, wavenumber, kink_ws, kink_ks, freq)

    def fluting(self, wavenumber, fluting_ws, fluting_ks, freq, m=2):
        if not self.kind.startswith("cylinder"):
            raise ValueError("fluting modes exist for the cylinder only")
        self._worker(int(m), wavenumber, fluting_ws, fluting_ks, freq)

    # -- the reference's driver loop, batched ----------------------------
    def default_speeds(self):
        return self.spec["speeds"](self.medium, self.solver.profile)

    def default_wavenumbers(self):
        lo, hi, n = self.spec["wavenumber"]
        return np.linspace(lo, hi, int(n))

    def run(self, wavenumber=None, speeds=None, n_freq=None, modes=("sausage", "kink")):
        """`for k in wavenumber: for i in range(len(speeds)-1): test_freq = linspace(speeds[i]*k,
        speeds[i+1]*k, n_freq)` (Density_cylinder.py:1142-1145) as one (k x phase-speed) grid per
        interval.  Defaults: the script's own wavenumber / speeds / sample count.

        Returns [sol_omegas1, sol_ks1, sol_omegas_kink1, sol_ks_kink1] like the pickle,
        plus further (omega, k) pairs for any extra modes requested."""
        wavenumber = self.default_wavenumbers() if wavenumber is None else np.asarray(wavenumber, dtype=np.float64)
        speeds = sorted(self.default_speeds() if speeds is None else speeds)
        n_freq = int(self.spec["n_freq"] if n_freq is None else n_freq)
        out = []
        for mode in modes:
            ws, ks = [], []
            for i in range(len(speeds) - 1):
                W = np.linspace(speeds[i], speeds[i + 1], n_freq)
                tab = self.solver.find_roots(mode, wavenumber, W, layout="phase_speed",
                                             tol_percent=self.tol)
                k_ok, w_ok = self._modes_of(tab)
                ks.append(k_ok)
                ws.append(w_ok)
            if "freq" in self.spec and speeds == sorted(self.default_speeds()):
                # a script whose driver also scans one absolute frequency array for every k
                tab = self.solver.find_roots(mode, wavenumber, self.spec["freq"](), layout="shared",
                                             tol_percent=self.tol)
                k_ok, w_ok = self._modes_of(tab)
                ks.append(k_ok)
                ws.append(w_ok)
            out.append(np.concatenate(ws) if ws else np.zeros(0))
            out.append(np.concatenate(ks) if ks else np.zeros(0))
        return out

    def run_and_pickle(self, path, **kw):
        """run() and write the script's own output file (write_root_table)."""
        results = self.run(**kw)
        write_root_table(path, results, self.script)
        return results

    def close(self):
        self.solver.close()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()
