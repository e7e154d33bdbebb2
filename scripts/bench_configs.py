"""Throughput of the full sweep (scan + brackets + refinement) on every BASELINE.json config at its
full size, one GPU.  Not the driver's bench (bench.py measures configs[1]); a record for profiles/."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import eigensolver_b200 as esb
from eigensolver_b200.scan import density_flow_grid, parameter_scan

CONFIGS = [
    ("configs[0] slab density, sausage+kink, 200 k x 2000 omega", "slab_density", {}, [0, 1],
     np.linspace(0.001, 0.75, 200), np.linspace(0.42, 2.95, 2000)),
    ("configs[1] cylinder density, n=0,1,2, 1000 k x 10000 omega", "cylinder_density", {}, [0, 1, 2],
     np.linspace(0.01, 4.5, 1000), np.linspace(0.5, 5.0, 10000)),
    ("configs[2] slab sheared flow, both branches, 2000 k x 20000 omega", "slab_flow",
     dict(medium=esb.FlowMedium(U_i0=0.35), profile=esb.GaussianFlow(1.0)), [0, 1],
     np.linspace(0.01, 4.5, 2000), np.linspace(-2.7, 2.7, 20000)),
    ("configs[3] cylinder rotational flow, n=0..3, 2000 k x 20000 omega", "cylinder_rotation",
     dict(profile=esb.PowerLawRotation(0.15, 1.25), s_end=0.01), [0, 1, 2, 3],
     np.linspace(0.25, 4.0, 2000), np.linspace(0.40, 1.6, 20000)),
]


def sweep(s, modes):
    ns = []
    for i in range(0, len(modes), 3):           # at most 3 modes per fused scan
        ns += s.sweep_resident_multi(modes[i:i + 3])
    return ns


out = []
for name, kind, kw, modes, k, W in CONFIGS:
    with esb.DispersionSolver(kind, **kw) as s:
        s.upload_axes(k, W)
        sweep(s, modes); torch.cuda.synchronize()
        t = time.perf_counter()
        reps = 3
        for _ in range(reps):
            ns = sweep(s, modes)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t) / reps
        evals = len(modes) * len(k) * len(W)
        rec = {"config": name, "n_steps": int(s.model.n_steps), "modes": modes, "evals": evals,
               "ms_per_sweep": 1e3 * dt, "evals_per_sec": evals / dt, "grid_kernel_ms_last": s.last_kernel_ms(),
               "brackets": int(sum(ns))}
        print(json.dumps(rec), flush=True)
        out.append(rec)

# configs[4]: parameter scan, density contrast x flow amplitude x k; 1e9 evaluations in total on 8 GPUs =
# 1.25e8 per GPU: here the one-GPU share (25 cylinder equilibria x 3 modes + 25 slab-flow equilibria x 2 modes)
dens, flow = density_flow_grid(np.linspace(0.12, 0.35, 25), np.linspace(0.05, 0.9, 25))
k = np.linspace(0.01, 4.5, 500)
tot = 0
Wd, Wf = np.linspace(0.5, 5.0, 2000), np.linspace(-2.7, 2.7, 2000)
with esb.DispersionSolver("cylinder_density") as sd, esb.DispersionSolver("slab_flow") as sf:
    parameter_scan(sd, dens[:1], k, Wd, [0, 1, 2]); parameter_scan(sf, flow[:1], k, Wf, [0, 1])    # warm-up
    torch.cuda.synchronize()
    t = time.perf_counter()
    r1 = parameter_scan(sd, dens, k, Wd, [0, 1, 2])
    tot += len(dens) * 3 * len(k) * 2000
    r2 = parameter_scan(sf, flow, k, Wf, [0, 1])
    tot += len(flow) * 2 * len(k) * 2000
    torch.cuda.synchronize()
    dt = time.perf_counter() - t
rec = {"config": "configs[4] parameter scan: 25 density contrasts (cylinder, 3 modes) + 25 flow amplitudes (slab, 2 modes), "
                 "500 k x 2000 omega each, incl. per-point table upload and root-table download",
       "evals": tot, "s_total": dt, "evals_per_sec": tot / dt,
       "modes_found": int(np.asarray(r1.table["accepted"]).sum() + np.asarray(r2.table["accepted"]).sum())}
print(json.dumps(rec), flush=True)
