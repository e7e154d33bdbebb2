"""Parameter-space scans (BASELINE configs[4]: density contrast x flow amplitude x k).

The reference has no scan driver: a user edits the speeds / profile constants at the top of
a solver script and reruns it.  Here a scan is a list of equilibria handed to the library as ONE
batched job (`DispersionSolver.scan_models` -> esb_scan_models: all tables uploaded once, the sweeps of
all equilibria enqueued back to back without host synchronisation, one compact result table).

Multi-GPU: every rank sweeps ALL equilibria on its strided share of the wavenumbers (rows r, r+N, ...),
so the ranks finish together whatever the cost of the individual equilibria (a list of 40 equilibria of
two different costs split over 8 ranks reached 68 % strong-scaling efficiency in round 1); the only
exchange is the gather of the accepted modes.
"""
from __future__ import annotations

import dataclasses

import numpy as np

from .distributed import shard_k
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
    """One equilibrium of a scan and what was found there (on this rank's wavenumbers)."""
    label: dict
    n_brackets: list          # per mode
    n_modes: list             # per mode: accepted roots
    tables: list = None       # per mode: dict of arrays (k_index, w_index, omega, ext, intq, accepted, iterations)


@dataclasses.dataclass
class ScanResult:
    points: list              # ScanPoint per equilibrium
    table: dict               # the compact table of this rank (numpy views of page-locked buffers)
    k: np.ndarray             # this rank's wavenumbers; global row = k_offset + k_index * k_stride
    k_offset: int
    k_stride: int
    guard: list = None        # guard_ends=True: discretisation-guard reports of the first and the last equilibrium


def _guard_ends(solver, points, modes, tol_percent):
    """The batched job carries no discretisation guard (esb_scan_models).  The two ends of the parameter
    range are swept once more as ordinary sweeps, which do: their reports bound the error of a scan over a
    one-parameter family.  The solver's own equilibrium is restored afterwards."""
    import warnings
    from .solver import DiscretisationWarning
    keep = (solver.spec.medium, solver.spec.profile)
    reports = []
    try:
        for p in (points[0], points[-1]) if len(points) > 1 else (points[0],):
            solver.reconfigure(p.get("medium"), p.get("profile"))
            solver.sweep_resident_multi(modes, tol_percent)
            rep = solver.guard_report()
            reports.append(rep)
            if rep["n_checked"] and rep["worst"] > rep["threshold"]:
                warnings.warn("parameter scan: discretisation error %.1e > %.0e at %s: raise n_steps"
                              % (rep["worst"], rep["threshold"], p.get("label", {})), DiscretisationWarning, stacklevel=3)
    finally:
        solver.reconfigure(*keep)
    return reports


def parameter_scan(solver: DispersionSolver, points, k, W, modes, layout="phase_speed", tol_percent=1.0,
                   rank=0, world=1, keep_tables=False, capacity_per_table=0, download=True,
                   guard_ends=False) -> ScanResult:
    """Sweep every equilibrium in `points` (list of dicts with optional keys 'medium', 'profile' and
    a free-form 'label') over the (k, W) grid: this rank's rows are k[rank::world], all equilibria.
    download=False: the compact table stays on the device (result.table is None; gather_scan_modes_device
    reads it in place).  guard_ends=True: the first and the last equilibrium are also swept on their own
    with the discretisation guard (two extra sweeps; result.guard holds the reports, DiscretisationWarning
    above the threshold)."""
    k_loc, k_off, k_stride = shard_k(np.asarray(k, dtype=np.float64), rank, world, layout="strided")
    solver.upload_axes(k_loc, W, layout)
    guard = _guard_ends(solver, points, modes, tol_percent) if guard_ends else None
    tab, nb = solver.scan_models(points, modes, tol_percent, capacity_per_table, download=download)
    if tab is None:
        pts = [ScanPoint(p.get("label", {}), [int(x) for x in nb[i]], None) for i, p in enumerate(points)]
        return ScanResult(pts, None, k_loc, k_off, k_stride, guard)
    out = []
    n_slots = len(list(modes))
    # entries of (model i, slot m) are contiguous and in this order
    bounds = np.concatenate([[0], np.cumsum(nb.reshape(-1))])
    for i, p in enumerate(points):
        n_modes, tables = [], []
        for m in range(n_slots):
            lo, hi = bounds[i * n_slots + m], bounds[i * n_slots + m + 1]
            n_modes.append(int(tab["accepted"][lo:hi].sum()))
            if keep_tables:
                tables.append({name: np.array(tab[name][lo:hi]) for name in
                               ("k_index", "w_index", "omega", "ext", "intq", "accepted", "iterations")})
        out.append(ScanPoint(p.get("label", {}), [int(x) for x in nb[i]], n_modes, tables if keep_tables else None))
    return ScanResult(out, tab, k_loc, k_off, k_stride, guard)


def gather_scan_modes(result: ScanResult, device, group=None):
    """All-gather the accepted modes of a scan: float64 tensor [total, 4] = (model, mode slot, global k
    row, omega) on `device`, every rank's share concatenated in rank order.  One count exchange, one
    padded payload exchange (NCCL over NVLink on GPUs, gloo in the CPU tests)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    t = result.table
    m = np.asarray(t["accepted"]) == 1
    mine = np.stack([np.asarray(t["model"])[m].astype(np.float64), np.asarray(t["slot"])[m].astype(np.float64),
                     np.asarray(t["k_index"])[m].astype(np.float64) * result.k_stride + result.k_offset,
                     np.asarray(t["omega"])[m]], axis=1) if m.any() else np.zeros((0, 4))
    mine = torch.as_tensor(mine, device=device)
    cnt = torch.tensor([mine.shape[0]], dtype=torch.int64, device=device)
    counts = [torch.zeros_like(cnt) for _ in range(world)]
    dist.all_gather(counts, cnt, group=group)
    counts = [int(c.item()) for c in counts]
    cap = max(max(counts), 1)
    pay = torch.zeros((cap, 4), dtype=torch.float64, device=device)
    pay[: mine.shape[0]] = mine
    bufs = [torch.zeros_like(pay) for _ in range(world)]
    dist.all_gather(bufs, pay, group=group)
    return torch.cat([b[:c] for b, c in zip(bufs, counts)], dim=0)


def gather_scan_modes_device(solver: DispersionSolver, result: ScanResult, device, group=None):
    """The same gather straight from the library's device buffers (no host round trip): NCCL over NVLink
    moves (model, mode slot, global k row, omega) of the accepted modes of every rank."""
    import torch
    import torch.distributed as dist
    from .distributed import _DevArray, _consumer_stream

    world = dist.get_world_size(group)
    info = solver.scan_table_device(stream=_consumer_stream(device))
    n = info["n"]
    if n:
        col = lambda name: torch.as_tensor(_DevArray(info[name][0], n, info[name][1]), device=device)
        m = col("accepted") == 1
        mine = torch.stack((col("model")[m].to(torch.float64), col("slot")[m].to(torch.float64),
                            col("k_index")[m].to(torch.float64) * float(result.k_stride) + float(result.k_offset),
                            col("omega")[m]), dim=1)
    else:
        mine = torch.zeros((0, 4), dtype=torch.float64, device=device)
    cnt = torch.tensor([mine.shape[0]], dtype=torch.int64, device=device)
    counts = torch.empty(world, dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(counts, cnt, group=group)
    counts = counts.tolist()
    cap = max(max(counts), 1)
    pay = torch.zeros((cap, 4), dtype=torch.float64, device=device)
    pay[: mine.shape[0]] = mine
    out = torch.empty((world * cap, 4), dtype=torch.float64, device=device)
    dist.all_gather_into_tensor(out, pay, group=group)
    return torch.cat([out[r * cap: r * cap + c] for r, c in enumerate(counts)], dim=0)


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
