"""Exploration (not a test): configs[2] / configs[3] sweep times for every library variant variants/lib_*.so
(built with -DESB_ROT_THREADS / -DESB_FLOW_THREADS) against the in-tree library."""
import glob, os, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
dst = os.path.join(ROOT, "eigensolver_b200", "libeigensolver_b200.so")
CODE = r'''
import sys, time; sys.path.insert(0, %r)
import numpy as np, eigensolver_b200 as esb
for name, kind, kw, modes, k, W in (
    ("configs[2]", "slab_flow", dict(medium=esb.FlowMedium(U_i0=0.35), profile=esb.GaussianFlow(1.0)), [0, 1],
     np.linspace(0.01, 4.5, 2000), np.linspace(-2.7, 2.7, 20000)),
    ("configs[3]", "cylinder_rotation", dict(profile=esb.PowerLawRotation(0.15, 1.25), s_end=0.01), [0, 1, 2, 3],
     np.linspace(0.25, 4.0, 2000), np.linspace(0.40, 1.6, 20000))):
    with esb.DispersionSolver(kind, **kw) as s:
        s.upload_axes(k, W); s.sweep_resident_multi(modes); s.lib.esb_tables_wait(s.ctx, None)
        t = time.perf_counter(); ks = []
        for _ in range(2):
            s.sweep_resident_multi(modes); ks.append(s.last_kernel_ms())
        s.lib.esb_tables_wait(s.ctx, None)
        print(name, "%%.2f ms per sweep, scan %%.2f ms" %% ((time.perf_counter() - t) * 500, sum(ks) / 2), flush=True)
''' % ROOT
shutil.copy(dst, dst + ".orig")
try:
    for lib in [dst + ".orig"] + sorted(glob.glob(os.path.join(ROOT, "variants", "lib_*.so"))):
        shutil.copy(lib, dst)
        out = subprocess.run([sys.executable, "-c", CODE], capture_output=True, text=True)
        print(os.path.basename(lib), "|", out.stdout.strip().replace("\n", " | "), out.stderr.strip()[-300:], flush=True)
finally:
    shutil.copy(dst + ".orig", dst)
    os.remove(dst + ".orig")
