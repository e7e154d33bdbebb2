"""Exploration: where does one equilibrium of a small-grid parameter scan spend its time?"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import eigensolver_b200 as esb
from eigensolver_b200.scan import density_flow_grid

dens, flow = density_flow_grid(np.linspace(0.12, 0.35, 6), np.linspace(0.05, 0.9, 6))
k = np.linspace(0.01, 4.5, 500)
for kind, pts, modes, W in (("cylinder_density", dens, [0, 1, 2], np.linspace(0.5, 5.0, 2000)),
                            ("slab_flow", flow, [0, 1], np.linspace(-2.7, 2.7, 2000))):
    with esb.DispersionSolver(kind) as s:
        s.upload_axes(k, W)
        for p in pts:
            t0 = time.perf_counter(); s.reconfigure(medium=p["medium"], profile=p["profile"])
            t1 = time.perf_counter(); ns = s.sweep_resident_multi(modes); torch.cuda.synchronize()
            t2 = time.perf_counter(); tabs = [s.download_roots_pinned(i) for i in range(len(ns))]
            t3 = time.perf_counter()
            its = np.concatenate([t.iterations for t in tabs])
            print("%-17s reconfigure %.2f ms  sweep %.2f ms (grid kernel %.2f)  download %.2f ms  brackets %d  max iters %d mean %.1f" % (
                kind, 1e3 * (t1 - t0), 1e3 * (t2 - t1), s.last_kernel_ms(), 1e3 * (t3 - t2), sum(ns), its.max(), its.mean()))
