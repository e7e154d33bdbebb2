"""Exploration: latency of ONE reference-style worker call, kink(k, ws, ks, freq) with 90 frequencies
(what the reference's own driver issues per process), and of the pieces underneath."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import eigensolver_b200 as esb


class Q:
    def put(self, x):
        self.x = x


for name in ("cylinder_density", "slab_density", "slab_flow", "rotation_sausage"):
    with esb.ReferenceScript(name) as sc:
        sp = sc.default_speeds()
        k = 2.0
        freq = np.linspace(sp[-2] * k, sp[-1] * k, 90)
        fn = sc.sausage
        for _ in range(5):
            fn(k, Q(), Q(), freq)
        n = 50
        t = time.perf_counter()
        for _ in range(n):
            q1, q2 = Q(), Q()
            fn(k, q1, q2, freq)
        dt = (time.perf_counter() - t) / n
        s = sc.solver
        t = time.perf_counter()
        for _ in range(n):
            s.dispersion_grid(0, [k], freq, layout="shared")
        dg = (time.perf_counter() - t) / n
        print("%-18s worker call %.3f ms (%d modes)   grid-only call %.3f ms   grid kernel %.3f ms" % (
            name, 1e3 * dt, len(q1.x), 1e3 * dg, s.last_kernel_ms()), flush=True)
