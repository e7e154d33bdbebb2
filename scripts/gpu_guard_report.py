"""Discretisation-guard report of every test case (bench-like windows) and the cost of the guard pass."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import eigensolver_b200 as esb
from helpers import CASES

for name, case in CASES.items():
    k = np.linspace(0.25, 4.0, 64) if case.kind == "cylinder_rotation" else np.linspace(0.5, 4.5, 64)
    W = np.linspace(case.W[0], case.W[1], 512)
    with case.gpu_solver(esb) as s:
        s.find_roots_multi(list(case.modes)[:2], k, W)
        print(name, s.guard_report(), flush=True)

k = np.linspace(0.01, 4.5, 1000); W = np.linspace(0.5, 5.0, 10000)
for guard in (0, -1, 256, 1024):
    with esb.DispersionSolver("cylinder_density", guard=guard) as s:
        s.upload_axes(k, W)
        for _ in range(3):
            s.sweep_resident_multi([0, 1, 2])
        s.lib.esb_tables_wait(s.ctx, None)
        t = time.perf_counter()
        for _ in range(10):
            s.sweep_resident_multi([0, 1, 2])
        s.lib.esb_tables_wait(s.ctx, None)
        print("guard", guard, "ms/sweep %.3f" % ((time.perf_counter() - t) * 100), s.guard_report(), flush=True)

# the sharp shell of the tests: what the guard sees against the explicit convergence check
k = np.linspace(0.5, 4.5, 64); W2 = np.linspace(4.6, 4.95, 512)
for width in (0.05, 0.04, 0.03):
    with esb.DispersionSolver("cylinder_density", profile=esb.GaussianDensity(width, x0=-0.5)) as s:
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            s.find_roots_multi([0, 1], k, W2)
        rep = s.guard_report()
        err, where = s.convergence_check([0, 1], k, W2)
        print("shell width", width, "guard", rep, "convergence_check", err, where, flush=True)
