"""Exploration (not a test): time of one resident sweep minus its grid kernel = brackets + refinement,
and a breakdown of the host-API (e2e) call.  ESB_REFINE_MINB selects the refine kernel variant."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import eigensolver_b200 as esb

k = np.linspace(0.01, 4.5, 1000); W = np.linspace(0.5, 5.0, 10000)
with esb.DispersionSolver("cylinder_density") as s:
    s.upload_axes(k, W)
    for _ in range(3):
        s.sweep_resident_multi([0, 1, 2])
    torch.cuda.synchronize()
    ts, ks = [], []
    for _ in range(5):
        t = time.perf_counter(); s.sweep_resident_multi([0, 1, 2]); torch.cuda.synchronize()
        ts.append(time.perf_counter() - t); ks.append(s.last_kernel_ms())
    print("MINB=%s sweep %.2f ms, grid kernel %.2f ms, rest (brackets+refine+syncs) %.2f ms" % (
        os.environ.get("ESB_REFINE_MINB", "4"), 1e3 * np.mean(ts), np.mean(ks), 1e3 * np.mean(ts) - np.mean(ks)))
    if "--e2e" in sys.argv:
        for _ in range(3):
            t0 = time.perf_counter(); s.upload_axes(k, W); torch.cuda.synchronize()
            t1 = time.perf_counter(); ns = s.sweep_resident_multi([0, 1, 2]); torch.cuda.synchronize()
            t2 = time.perf_counter(); tabs = [s.download_roots(n, slot) for slot, n in enumerate(ns)]
            t3 = time.perf_counter()
            print("e2e: upload %.2f ms, sweep %.2f ms, download %.2f ms (%d roots, %.1f MB)" % (
                1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2), sum(ns), sum(ns) * 40 / 1e6))
