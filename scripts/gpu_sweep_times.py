"""Exploration (not a test): sweep time / scan-kernel time of every BASELINE config (lane schedule for the big ones)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import eigensolver_b200 as esb
cases = [
    ("configs[1]", "cylinder_density", {}, [0, 1, 2], np.linspace(0.01, 4.5, 1000), np.linspace(0.5, 5.0, 10000)),
    ("configs[1]/8", "cylinder_density", {}, [0, 1, 2], np.linspace(0.01, 4.5, 1000)[::8], np.linspace(0.5, 5.0, 10000)),
    ("configs[0]", "slab_density", {}, [0, 1], np.linspace(0.001, 0.75, 200), np.linspace(0.41, 2.95, 2000)),
    ("configs[2]", "slab_flow", dict(medium=esb.FlowMedium(U_i0=0.35), profile=esb.GaussianFlow(1.0)), [0, 1], np.linspace(0.01, 4.5, 2000), np.linspace(-2.7, 2.7, 20000)),
    ("configs[2]/16", "slab_flow", dict(medium=esb.FlowMedium(U_i0=0.35), profile=esb.GaussianFlow(1.0)), [0, 1], np.linspace(0.01, 4.5, 1000)[::8], np.linspace(-2.7, 2.7, 10000)),
    ("configs[3]", "cylinder_rotation", dict(profile=esb.PowerLawRotation(0.15, 1.25), s_end=0.01), [0, 1, 2, 3], np.linspace(0.25, 4.0, 2000), np.linspace(0.40, 1.6, 20000)),
]
for name, kind, kw, modes, k, W in cases:
    with esb.DispersionSolver(kind, **kw) as s:
        s.upload_axes(k, W)
        for _ in range(2):
            s.sweep_resident_multi(modes); s.lib.esb_tables_wait(s.ctx, None)
        ts, ks = [], []
        for _ in range(4):
            torch.cuda.synchronize(); t = time.perf_counter()
            ns = s.sweep_resident_multi(modes); s.lib.esb_tables_wait(s.ctx, None)
            ts.append(time.perf_counter() - t); ks.append(s.last_kernel_ms())
        ev = len(modes) * k.size * W.size
        print("%-14s %s sweep %.2f ms grid %.2f ms rest %.2f ms  %.3e eval/s  brackets %d" % (name, s.spec.scheme, 1e3 * np.mean(ts), np.mean(ks), 1e3 * np.mean(ts) - np.mean(ks), ev / np.mean(ts), sum(ns)), flush=True)
