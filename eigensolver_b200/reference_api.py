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
as ONE batched GPU sweep instead of one process per (k, interval).
"""
from __future__ import annotations

import numpy as np

from .solver import (CYLINDER_CORONAL, SLAB_CORONAL, DispersionSolver, GaussianDensity, Medium)


class _ListSink:
    """Minimal stand-in for multiprocessing.Queue when the caller has none."""

    def __init__(self):
        self.items = []

    def put(self, x):
        self.items.append(x)


class ReferenceScript:
    """One reference solver script (geometry + equilibrium + profile width)."""

    def __init__(self, kind="cylinder_density", medium=None, width=None, x0=0.0, tol=1.0, device=0,
                 n_steps=None, scheme="rk8"):
        if medium is None:
            medium = CYLINDER_CORONAL if kind == "cylinder_density" else SLAB_CORONAL
        if width is None:
            width = 0.95 if kind == "cylinder_density" else 0.9   # dr :125 / dx :94
        self.kind = kind
        self.medium = medium
        self.tol = tol            # xi_tol :522 / p_tol :143  (percent)
        self.solver = DispersionSolver(kind, medium, GaussianDensity(width, x0), n_steps=n_steps,
                                       scheme=scheme, device=device)

    # -- the reference's worker signature -------------------------------
    def _worker(self, mode, wavenumber, ws_queue, ks_queue, freq):
        freq = np.asarray(freq, dtype=np.float64)
        tab = self.solver.find_roots(mode, [float(wavenumber)], freq, layout="shared",
                                     tol_percent=self.tol)
        ks, ws = tab.modes()
        ks_queue.put(list(ks))
        ws_queue.put(list(ws))

    def sausage(self, wavenumber, sausage_ws, sausage_ks, freq):
        self._worker(0, wavenumber, sausage_ws, sausage_ks, freq)

    def kink(self, wavenumber, kink_ws, kink_ks, freq):
        self._worker(1, wavenumber, kink_ws, kink_ks, freq)

    def fluting(self, wavenumber, fluting_ws, fluting_ks, freq, m=2):
        if self.kind != "cylinder_density":
            raise ValueError("fluting modes exist for the cylinder only")
        self._worker(int(m), wavenumber, fluting_ws, fluting_ks, freq)

    # -- the reference's driver loop, batched ----------------------------
    def default_speeds(self):
        md = self.medium
        if self.kind == "cylinder_density":
            # Density_cylinder.py:225 (including its `cT_e -c_e` element, a missing comma)
            sp = [md.c_i0, md.c_e, md.vA_i0, md.vA_e, md.cT_i0, md.cT_e - md.c_e, -md.c_i0, -md.vA_i0,
                  -md.vA_e, -md.cT_i0, -md.cT_e]
        else:
            sp = [1.0, 0.9, 1.2, 1.3]
        return sorted(sp)

    def run(self, wavenumber, speeds=None, n_freq=90, modes=("sausage", "kink")):
        """`for k in wavenumber: for i in range(len(speeds)-1): test_freq = linspace(speeds[i]*k,
        speeds[i+1]*k, n_freq)` (:1142-1145) as one (k x phase-speed) grid per interval.

        Returns [sol_omegas1, sol_ks1, sol_omegas_kink1, sol_ks_kink1] like the pickle,
        plus further (omega, k) pairs for any extra modes requested."""
        wavenumber = np.asarray(wavenumber, dtype=np.float64)
        speeds = sorted(self.default_speeds() if speeds is None else speeds)
        out = []
        for mode in modes:
            ws, ks = [], []
            for i in range(len(speeds) - 1):
                W = np.linspace(speeds[i], speeds[i + 1], int(n_freq))
                tab = self.solver.find_roots(mode, wavenumber, W, layout="phase_speed",
                                             tol_percent=self.tol)
                k_ok, w_ok = tab.modes()
                ks.append(k_ok)
                ws.append(w_ok)
            out.append(np.concatenate(ws) if ws else np.zeros(0))
            out.append(np.concatenate(ks) if ks else np.zeros(0))
        return out

    def close(self):
        self.solver.close()
