"""Exploration: lane-per-bracket vs warp-per-bracket refinement as a function of the bracket count."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import eigensolver_b200 as esb

for kind, modes, W, kw in (("cylinder_density", [0, 1, 2], np.linspace(0.5, 5.0, 2000), {}),
                           ("slab_flow", [0, 1], np.linspace(-2.7, 2.7, 2000), dict(profile=esb.GaussianFlow(1.0))),
                           ("cylinder_rotation", [0, 1, 2], np.linspace(0.4, 1.6, 2000),
                            dict(profile=esb.PowerLawRotation(0.15, 1.25), s_end=0.01)),
                           ("slab_density", [0, 1], np.linspace(0.42, 2.95, 2000), {})):
    with esb.DispersionSolver(kind, **kw) as s:
        for nk in (25, 50, 100, 200, 400, 800, 1600):
            k = np.linspace(0.05, 4.5, nk)
            s.upload_axes(k, W)
            row = []
            for mode in ("lane", "warp"):
                s.set_schedule(mode)
                s.sweep_resident_multi(modes); torch.cuda.synchronize()
                ts = []
                for _ in range(3):
                    t = time.perf_counter(); ns = s.sweep_resident_multi(modes); torch.cuda.synchronize()
                    ts.append(1e3 * (time.perf_counter() - t) - s.last_kernel_ms())
                row.append(min(ts))
            print("%-18s nk %4d brackets %7d   lane %.2f ms   warp %.2f ms" % (kind, nk, sum(ns), row[0], row[1]), flush=True)
