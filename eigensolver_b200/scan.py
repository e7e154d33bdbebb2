"""Parameter-space scans (BASELINE configs[4]: density contrast x flow amplitude x k).

The reference has no scan driver: a user edits the speeds / profile constants at the top of
a solver script and reruns it.  Here a scan is a list of equilibria; every one is a small
table upload (`DispersionSolver.reconfigure`) followed by the same fused sweep, and the list
shards across GPUs exactly like the k axis does (no data-path collective, root tables
gathered at the end).
"""
from __future__ import annotations

import dataclasses

import numpy as np

from .distributed import shard_bounds
from .solver import DispersionSolver, FlowMedium, GaussianDensity, GaussianFlow, Medium


def medium_for_density_contrast(base: Medium, contrast: float) -> Medium:
    """Equilibrium with rho_e/rho_i0 = contrast: the reference fixes rho_e by total-pressure
    balance, rho_e = rho_i0 (c_i0^2 + g/2 vA_i0^2)/(c_e^2 + g/2 vA_e^2) (Density_cylinder.py:80);
    the exterior Alfven speed is the free constant that realises a requested contrast."""
    g = base.gamma
    vAe2 = ((base.c_i0**2 + 0.5 * g * base.vA_i0**2) / contrast - base.c_e**2) * 2.0 / g
    if vAe2 <= 0:
        raise ValueError("density contrast %g is not reachable with c_e = %g" % (contrast, base.c_e))
    return dataclasses.replace(base, vA_e=float(np.sqrt(vAe2)))


@dataclasses.dataclass
class ScanPoint:
    """One equilibrium of a scan and what was found there."""
    label: dict
    n_brackets: list          # per mode
    n_modes: list             # per mode: accepted roots
    tables: list = None       # RootTable per mode if keep_tables


def _scan_point(solver, p, modes, tol_percent, keep_tables):
    solver.reconfigure(medium=p.get("medium"), profile=p.get("profile"))
    ns = solver.sweep_resident_multi(modes, tol_percent)
    # page-locked views (one packed copy per slot); copied out only if the caller keeps them
    tabs = [solver.download_roots_pinned(slot) for slot in range(len(ns))]
    n_modes = [int(t.accepted.sum()) for t in tabs]
    kept = [dataclasses.replace(t, **{f.name: np.array(getattr(t, f.name)) for f in dataclasses.fields(t)
                                      if isinstance(getattr(t, f.name), np.ndarray)})
            for t in tabs] if keep_tables else None
    return ScanPoint(p.get("label", {}), ns, n_modes, kept)


def parameter_scan(solver: DispersionSolver, points, k, W, modes, layout="phase_speed", tol_percent=1.0,
                   rank=0, world=1, keep_tables=False):
    """Sweep every equilibrium in `points` (list of dicts with optional keys 'medium', 'profile' and
    a free-form 'label') over the same (k, W) grid on this rank's share of the list.

    Returns the list of ScanPoint for THIS rank, in the order of `points` (use
    torch.distributed.all_gather_object or eigensolver_b200.distributed to combine ranks)."""
    lo, hi = shard_bounds(len(points), rank, world)
    solver.upload_axes(k, W, layout)
    return [_scan_point(solver, p, modes, tol_percent, keep_tables) for p in points[lo:hi]]


def density_flow_grid(contrasts, flow_amplitudes, base_density: Medium = None, base_flow: FlowMedium = None,
                      width=0.95, flow_width=1.0):
    """The two families of configs[4]: cylinder density models over `contrasts` and slab flow
    models over `flow_amplitudes` (the reference has no single script with both a density and a
    flow profile in the layer).  Returns (density_points, flow_points) for parameter_scan."""
    from .solver import CYLINDER_CORONAL, SLAB_FLOW_CORONAL
    base_density = base_density or CYLINDER_CORONAL
    base_flow = base_flow or SLAB_FLOW_CORONAL
    dens = [{"medium": medium_for_density_contrast(base_density, c), "profile": GaussianDensity(width),
             "label": {"density_contrast": float(c)}} for c in contrasts]
    flow = [{"medium": dataclasses.replace(base_flow, U_i0=float(a)), "profile": GaussianFlow(flow_width),
             "label": {"flow_amplitude": float(a)}} for a in flow_amplitudes]
    return dens, flow
