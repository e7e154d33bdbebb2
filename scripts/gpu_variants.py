"""Exploration (not a test): time the scan kernel and the rest of the sweep for every library variant
gpurun_out/lib_*.so (built with different -DESB_GRID_MINB / -DESB_REFINE_MINB)."""
import glob, os, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
dst = os.path.join(ROOT, "eigensolver_b200", "libeigensolver_b200.so")
shutil.copy(dst, dst + ".orig")
for lib in sorted(glob.glob(os.path.join(ROOT, "variants", "lib_*.so"))):
    shutil.copy(lib, dst)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "gpu_refine_time.py")], capture_output=True, text=True)
    print(os.path.basename(lib), out.stdout.strip().replace("MINB=4 ", ""), out.stderr.strip()[-200:], flush=True)
shutil.copy(dst + ".orig", dst)
