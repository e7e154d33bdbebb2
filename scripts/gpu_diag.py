"""Diagnostic (not a test): where does the GPU path deviate most from the C oracle?"""
import sys, os, warnings
warnings.filterwarnings("ignore")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import eigensolver_b200 as esb
from helpers import CASES, regular_mask
from oracle import rk_oracle as ork

names = sys.argv[1:] or list(CASES)
for name in names:
    case = CASES[name]
    k = np.linspace(0.05, 4.5, 20); W = np.linspace(case.W[0], case.W[1], 240)
    model = case.c_model()
    for mode in case.modes:
        e0, i0 = ork.grid(model, mode, k, W)
        ok = case.regular(k, W, mode) & ~np.isnan(e0)
        for n in (128, 144, 192, 256):
            with case.gpu_solver(esb, n_steps=n) as s:
                e, i = s.dispersion_grid(mode, k, W)
            nanmis = (np.isnan(e) != np.isnan(e0)).sum()
            dev = np.abs((e - i) - (e0 - i0)) / np.maximum(np.abs(e0), np.abs(i0))
            dev = np.where(ok, dev, 0)
            j = np.unravel_index(np.nanargmax(dev), dev.shape)
            print("%-22s mode %d N=%3d  max dev %.2e at k=%.3f W=%.4f  nan-mismatch %d  ext dev %.1e" % (
                name, mode, n, dev[j], k[j[0]], W[j[1]], nanmis,
                np.nanmax(np.where(ok, np.abs(e - e0) / np.abs(e0), 0))), flush=True)
