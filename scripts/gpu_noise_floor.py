"""Noise-floor map of the BASELINE grids: per config and mode, the fraction of evaluated grid
points and of sign-change brackets that lie ABOVE the noise floor, i.e. outside the resonant continua
(tests/helpers.py regular masks) - the region parity with the reference is claimed for.  Rows subsampled
(every `step`-th wavenumber); writes profiles/r02_noise_floor.json."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import eigensolver_b200 as esb
from helpers import CASES
from test_gpu_parity import BASELINE_GRIDS

out = []
for config, (name, modes, k_full, W) in BASELINE_GRIDS.items():
    case = CASES[name]
    step = 40 if name == "cylinder_rotation" else 10 if len(k_full) > 500 else 2
    k = k_full[::step]
    with case.gpu_solver(esb) as s:
        s.set_schedule("lane")
        for m in modes:
            e, i = s.dispersion_grid(m, k, W)
            tab = s.find_roots(m, k, W)
            fin = np.isfinite(e) & np.isfinite(i)
            reg = case.regular(k, W, m)
            ok_iv = reg[:, :-1] & reg[:, 1:]
            sel = ok_iv[tab.k_index, tab.w_index]
            rec = {"config": config, "kind": name, "mode": m, "rows": int(k.size), "columns": int(W.size),
                   "evaluated_fraction": float(fin.mean()),
                   "points_above_floor": float((reg & fin).sum() / max(fin.sum(), 1)),
                   "brackets": int(len(tab.omega)), "brackets_above_floor": int(sel.sum()),
                   "brackets_above_floor_fraction": float(sel.mean()) if len(sel) else None,
                   "accepted_above_floor": int((sel & (tab.accepted == 1)).sum()),
                   "accepted_total": int(tab.accepted.sum())}
            out.append(rec)
            print(rec, flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "r02_noise_floor.json"), "w"), indent=1)
