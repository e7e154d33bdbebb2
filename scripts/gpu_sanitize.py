"""Small sweeps of every kind through both scan kernels and both refinement kernels: a quick all-kernels exercise
(written for compute-sanitizer, which is closed on this pool)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import eigensolver_b200 as esb

CASES = [("cylinder_density", {}, [0, 1, 2], (0.55, 4.95)),
         ("slab_density", {}, [0, 1], (0.42, 2.95)),
         ("slab_flow", dict(profile=esb.GaussianFlow(1.0)), [0, 1], (-2.6, 2.6)),
         ("cylinder_flow", {}, [0, 1, 2], (-4.9, 4.9)),
         ("cylinder_rotation", dict(profile=esb.PowerLawRotation(0.15, 1.25), s_end=0.01), [0, 1, 2], (0.5, 1.45))]
for kind, kw, modes, (lo, hi) in CASES:
    with esb.DispersionSolver(kind, **kw) as s:
        for nk, nw in ((3, 129), (12, 900)):            # warp-per-point / thread-per-point scan
            k = np.linspace(0.6, 3.5, nk); W = np.linspace(lo, hi, nw)
            for mode in ("lane", "warp"):
                s.set_schedule(mode)
                tabs = s.find_roots_multi(modes, k, W)
                t1 = s.find_roots(modes[-1], k, W, max_roots=4096)
                p = s.download_roots_pinned(0)
            print(kind, nk, nw, [len(t.omega) for t in tabs], int(sum(t.accepted.sum() for t in tabs)), flush=True)
print("done")
