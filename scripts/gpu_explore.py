"""Exploratory GPU check (not a test): parity of the CUDA path vs the C oracle and timing."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import eigensolver_b200 as esb
from oracle import rk_oracle as ork

def rel(D, e0, i0):
    return np.abs(D - (e0 - i0)) / np.maximum(np.abs(e0), np.abs(i0))

for kind, modes, Ws in (("cylinder_density", (0, 1, 2), [(0.52, 0.88), (1.32, 1.98), (2.95, 4.95)]),
                        ("slab_density", (0, 1), [(0.42, 0.75), (1.72, 2.95)])):
    model = ork.make_model(kind)
    k = np.linspace(0.1, 4.5, 12)
    for scheme, n in (("rk8", None), ("rk8", 256), ("rk4", 2048)):
        with esb.DispersionSolver(kind, scheme=scheme, n_steps=n) as s:
            for mode in modes:
                for (a, b) in Ws:
                    W = np.linspace(a, b, 40)
                    ext, inq = s.dispersion_grid(mode, k, W)
                    e0, i0 = ork.grid(model, mode, k, W)
                    r = rel(ext - inq, e0, i0)
                    re = np.abs(ext - e0) / np.abs(e0)
                    print(kind, scheme, n, 'mode', mode, 'W', (a, b), 'max rel D %.2e  ext %.2e  nan %d' % (np.nanmax(r), np.nanmax(re), np.isnan(r).sum()), flush=True)

# timing on a bigger grid
for kind in ("cylinder_density", "slab_density"):
    with esb.DispersionSolver(kind) as s:
        k = np.linspace(0.01, 4.5, 1000)
        W = np.linspace(0.5, 4.99, 2000) if kind == "cylinder_density" else np.linspace(0.41, 2.99, 2000)
        for it in range(3):
            t = time.time(); ext, inq = s.dispersion_grid(1, k, W); dt = time.time() - t
            print(kind, 'grid 1000x2000: host call %.3f s, kernel %.3f ms, %.3e evals/s (kernel)' % (dt, s.last_kernel_ms(), ext.size / (s.last_kernel_ms() * 1e-3)), flush=True)
        t = time.time(); roots = s.find_roots(1, k, W); dt = time.time() - t
        print(kind, 'find_roots: %.3f s, brackets %d accepted %d, iters max %d mean %.1f' % (dt, roots.n_brackets, roots.accepted.sum(), roots.iterations.max(), roots.iterations.mean()), flush=True)
