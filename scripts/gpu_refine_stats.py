"""Exploration (not a test): bracket / evaluation statistics of the bench sweep and the time of its parts."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import eigensolver_b200 as esb

k = np.linspace(0.01, 4.5, 1000); W = np.linspace(0.5, 5.0, 10000)
with esb.DispersionSolver("cylinder_density") as s:
    s.upload_axes(k, W)
    for _ in range(3):
        ns = s.sweep_resident_multi([0, 1, 2])
    torch.cuda.synchronize()
    ts, ks = [], []
    for _ in range(5):
        t = time.perf_counter(); s.sweep_resident_multi([0, 1, 2]); s.lib.esb_tables_wait(s.ctx, None)
        ts.append(time.perf_counter() - t); ks.append(s.last_kernel_ms())
    print("sweep %.2f ms, grid kernel %.2f ms, rest (brackets+refine) %.2f ms" % (
        1e3 * np.mean(ts), np.mean(ks), 1e3 * np.mean(ts) - np.mean(ks)))
    tot_it = 0
    for slot, n in enumerate(ns):
        t = s.download_roots(n, slot)
        it = t.iterations
        tot_it += int(it.sum())
        print("slot %d: %d brackets, %d accepted, %d poles from the scan, evaluations: mean %.2f max %d, hist %s" % (
            slot, n, int(t.accepted.sum()), int((it == 0).sum()), it.mean(), it.max(),
            np.bincount(np.minimum(it, 24))))
    print("total refinement evaluations %d = %.3f per grid point" % (tot_it, tot_it / (3 * k.size * W.size)))
