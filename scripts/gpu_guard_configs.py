"""Discretisation-guard report of the full-size BASELINE grids (configs[0..3])."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import eigensolver_b200 as esb

cases = [("configs[0]", "slab_density", {}, [0, 1], np.linspace(0.001, 0.75, 200), np.linspace(0.41, 2.95, 2000)),
         ("configs[1]", "cylinder_density", {}, [0, 1, 2], np.linspace(0.01, 4.5, 1000), np.linspace(0.5, 5.0, 10000)),
         ("configs[2]", "slab_flow", dict(medium=esb.FlowMedium(U_i0=0.35), profile=esb.GaussianFlow(1.0)), [0, 1],
          np.linspace(0.01, 4.5, 2000), np.linspace(-2.7, 2.7, 20000)),
         ("configs[3]", "cylinder_rotation", dict(profile=esb.PowerLawRotation(0.15, 1.25), s_end=0.01), [0, 1, 2, 3],
          np.linspace(0.25, 4.0, 2000), np.linspace(0.40, 1.6, 20000)),
         ("configs[3] kink law", "cylinder_rotation", dict(profile=esb.PowerLawRotation(0.25, 0.8), s_end=0.001), [0, 1, 2, 3],
          np.linspace(0.25, 4.0, 2000), np.linspace(0.40, 1.6, 20000))]
for name, kind, kw, modes, k, W in cases:
    with esb.DispersionSolver(kind, **kw) as s:
        s.upload_axes(k, W)
        s.sweep_resident_multi(modes)
        rep = s.guard_report()
        print(name, {a: rep[a] for a in ("worst", "n_checked", "n_above", "stride", "slot", "k_index", "w_index")},
              "k %.3f W %.4f" % (k[max(rep["k_index"], 0)], W[max(rep["w_index"], 0)]), flush=True)
