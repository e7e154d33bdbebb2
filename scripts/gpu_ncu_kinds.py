"""One full-size scan per kind (for an ncu capture of every scan kernel): 1000 k x 10000 omega."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import eigensolver_b200 as esb

k = np.linspace(0.25, 4.0, 1000)
for kind, kw, modes, W in (("slab_density", {}, [0, 1], np.linspace(0.42, 2.95, 10000)),
                           ("slab_flow", dict(profile=esb.GaussianFlow(1.0)), [0, 1], np.linspace(-2.7, 2.7, 10000)),
                           ("cylinder_flow", {}, [0, 1, 2], np.linspace(-4.9, 4.9, 10000)),
                           ("cylinder_rotation", dict(profile=esb.PowerLawRotation(0.15, 1.25), s_end=0.01), [0, 1, 2],
                            np.linspace(0.4, 1.6, 10000))):
    with esb.DispersionSolver(kind, **kw) as s:
        s.upload_axes(k, W)
        s.sweep_resident_multi(modes)
        print(kind, s.last_kernel_ms(), flush=True)
