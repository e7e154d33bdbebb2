"""Exploration (not a test): what the step count costs on the device - sweep time and scan-kernel time of the
configs[1] grid at several n_steps, with the guard's report (the accuracy side: scripts/host_steps_study.py)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import eigensolver_b200 as esb
k, W, modes = np.linspace(0.01, 4.5, 1000), np.linspace(0.5, 5.0, 10000), [0, 1, 2]
for n in (112, 128, 136, 152, 176):
    with esb.DispersionSolver("cylinder_density", n_steps=n) as s:
        s.upload_axes(k, W)
        for _ in range(2):
            s.sweep_resident_multi(modes); s.lib.esb_tables_wait(s.ctx, None)
        ts, ks = [], []
        for _ in range(5):
            torch.cuda.synchronize(); t = time.perf_counter()
            ns = s.sweep_resident_multi(modes); s.lib.esb_tables_wait(s.ctx, None)
            ts.append(time.perf_counter() - t); ks.append(s.last_kernel_ms())
        rep = s.guard_report()
        ev = len(modes) * k.size * W.size
        print("n_steps %d sweep %.2f ms scan %.2f ms  %.3e eval/s  brackets %d  guard worst %.1e (%d judged)"
              % (n, 1e3 * np.mean(ts), np.mean(ks), ev / np.mean(ts), sum(s.last_n_brackets), rep["worst"], rep["n_checked"]),
              flush=True)
