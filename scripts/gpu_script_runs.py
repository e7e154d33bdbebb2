"""The drop-in at the scripts' own settings: ReferenceScript(name).run() - the whole driver loop of each of the 11
solver scripts (its wavenumber grid, speed intervals, sample count, tolerance; the scripts' own scan / bisection
rule) - timed end to end from host arrays to the four output lists, with the number of D evaluations of the scan
(bisection evaluations not counted) and of solutions found.  The reference spends ~0.15 s per evaluation."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import eigensolver_b200 as esb
from eigensolver_b200.reference_api import SCRIPTS

for name in SCRIPTS:
    with esb.ReferenceScript(name) as script:
        script.run()                                   # warm-up (allocations, first launches)
        t = time.perf_counter()
        out = script.run()
        dt = time.perf_counter() - t
        nk = len(script.default_wavenumbers())
        n_int = len(script.default_speeds()) - 1
        n_eval = 2 * nk * (n_int * script.spec["n_freq"] + (len(script.spec["freq"]()) if "freq" in script.spec else 0))
        print("%-30s %4d k x %2d intervals x %3d freq x 2 modes = %7d scan evaluations  %8.2f ms  solutions %d + %d"
              % (name, nk, n_int, script.spec["n_freq"], n_eval, 1e3 * dt, len(out[0]), len(out[2])), flush=True)
