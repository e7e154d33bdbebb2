import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import eigensolver_b200 as esb
from test_gpu_parity import DROPIN_OVERRIDES
g = np.load("tests/golden/ref_scans.npz")
n = 0
scripts = {}
while "c%d_script" % n in g.files:
    name = str(g["c%d_script" % n])
    if name not in scripts:
        scripts[name] = esb.ReferenceScript(name, **DROPIN_OVERRIDES[name])
    sc = scripts[name]
    mode = int(g["c%d_mode" % n][0]); k = float(g["c%d_k" % n][0]); freq = g["c%d_freq" % n]; ref = np.sort(g["c%d_sol_ws" % n])
    t = sc.solver.find_roots(mode, [k], freq, layout="shared", tol_percent=sc.tol)
    got = np.sort(t.omega[t.accepted == 1])
    same = len(got) == len(ref) and np.allclose(got, ref, rtol=1e-9, atol=0)
    print(n, name, mode, k, "ref", ref, "got", got, "OK" if same else "MISMATCH", "rule", sc.rule)
    if not same:
        e, i = sc.solver.dispersion_grid(mode, [k], freq, layout="shared")
        D = (e - i)[0]; pct = (np.abs(e - i) * 100 / np.maximum(abs(e), abs(i)))[0]
        sg = np.sign(D)
        ch = np.nonzero(sg[:-1] * sg[1:] < 0)[0]
        print("   sign changes at", ch, "freq", freq[ch], "pct lo", pct[ch], "pct hi", pct[ch + 1], "nan", np.isnan(D).sum())
        print("   in-band grid points", np.nonzero(pct < sc.tol)[0], "tol", sc.tol)
        print("   table: w_index", t.w_index, "omega", t.omega, "acc", t.accepted, "it", t.iterations)
    n += 1
