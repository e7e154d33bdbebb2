"""Exploration (not a test): Brent iteration statistics of the refine kernel on the bench workload."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import eigensolver_b200 as esb

k = np.linspace(0.01, 4.5, 1000); W = np.linspace(0.5, 5.0, 10000)
with esb.DispersionSolver("cylinder_density") as s:
    tabs = s.find_roots_multi([0, 1, 2], k, W)
for m, t in enumerate(tabs):
    acc = t.accepted == 1
    it = t.iterations
    nan = ~np.isfinite(t.ext)
    print("mode %d: %d brackets, %d modes, %d poles/other, %d nan" % (m, len(it), acc.sum(), (~acc).sum(), nan.sum()))
    for name, sel in (("modes", acc), ("unaccepted", ~acc)):
        if sel.sum():
            h = np.bincount(it[sel], minlength=12)
            print("   %-10s iterations mean %.2f max %d  hist %s" % (name, it[sel].mean(), it[sel].max(), list(h[:40])))
    print("   total evaluations %d = %.3f x grid points" % (it.sum(), it.sum() / (len(k) * len(W))))
