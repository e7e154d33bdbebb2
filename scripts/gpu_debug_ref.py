import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import eigensolver_b200 as esb
g = np.load("tests/golden/ref_scan_cylinder_density_coronal.npz")
freq = g["scan0_freq"]; k = float(g["scan0_k"][0])
with esb.DispersionSolver("cylinder_density") as s:
    e, i = s.dispersion_grid(1, [k], freq, layout="shared")
    pct = np.abs(e-i)*100/np.maximum(abs(e),abs(i))
    print("D", (e-i)[0,:6], "pct", pct[0,:6])
    for rule in ("converged", "reference"):
        s.set_accept_rule(rule)
        for sched in ("auto", "lane"):
            s.set_schedule(sched)
            t = s.find_roots(1, [k], freq, layout="shared")
            print(rule, sched, "n", len(t.omega), "w_index", t.w_index, "omega", t.omega, "acc", t.accepted, "it", t.iterations, "nb", t.n_brackets)
print("ref", g["scan0_sol_ws"])
