"""Exploration (not a test): accuracy of the graded mesh (mesh=2) vs its parameters and N.
Reference = sin^2-clustered mesh at N = 1024 on the GPU; max deviation over the regular region."""
import sys, os, warnings, itertools
warnings.filterwarnings("ignore")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import eigensolver_b200 as esb
from helpers import CASES

names = sys.argv[1:] or ["cylinder_density", "cylinder_photospheric", "cylinder_flow", "cylinder_rotation"]
for name in names:
    case = CASES[name]
    k = np.linspace(0.05, 4.5, 40); W = np.linspace(case.W[0], case.W[1], 480)
    modes = list(case.modes)
    with case.gpu_solver(esb, n_steps=1024) as s:
        E0, I0 = s.dispersion_grid_multi(modes, k, W)
    ok = np.array([case.regular(k, W, m) for m in modes]) & ~np.isnan(E0)
    scale = np.maximum(np.abs(E0), np.abs(I0))

    def dev(**kw):
        with case.gpu_solver(esb, **kw) as s:
            E, I = s.dispersion_grid_multi(modes, k, W)
        d = np.abs((E - I) - (E0 - I0)) / scale
        d = np.where(ok, d, 0)
        return [float(np.nanmax(d[i])) for i in range(len(modes))]

    slab = case.kind.startswith("slab")
    for n in (128, 160, 192, 256, 384):
        print("%-22s clustered N=%d  %s" % (name, n, " ".join("%.1e" % v for v in dev(n_steps=n))), flush=True)
    if slab:
        Ns = (128, 160, 192, 224, 256, 288, 320, 352, 384)
        combos = [(0.0, e) for e in ((0, 0), (0.04, 0.1), (0.05, 0.2), (0.1, 0.5))]
    else:
        Ns = (64, 80, 96, 112, 128)
        combos = list(itertools.product((0.08, 0.16, 0.3), ((0, 0), (0.02, 0.1))))
    for n in Ns:
        for ax, (ed, ew) in combos:
            d = dev(n_steps=n, mesh="graded", mesh_params=(ax, ed, ew))
            print("%-22s graded N=%d axis=%.2f edge=(%.2f,%.2f)  %s  max %.1e" % (
                name, n, ax, ed, ew, " ".join("%.1e" % v for v in d), max(d)), flush=True)
