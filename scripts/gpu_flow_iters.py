import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import eigensolver_b200 as esb
k = np.linspace(0.01, 4.5, 1000)[::8]; W = np.linspace(-2.7, 2.7, 10000)
for U in (0.05, 0.35, 0.9):
    with esb.DispersionSolver("slab_flow", medium=esb.FlowMedium(U_i0=U), profile=esb.GaussianFlow(1.0)) as s:
        s.set_schedule("lane")
        s.upload_axes(k, W)
        for rep in range(2):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            ns = s.sweep_resident_multi([0, 1]); s.lib.esb_tables_wait(s.ctx, None)
            t1 = time.perf_counter()
        print("U=%.2f sweep %.2f ms grid %.2f ms" % (U, 1e3 * (t1 - t0), s.last_kernel_ms()))
        for slot, n in enumerate(ns):
            t = s.download_roots(n, slot)
            it = t.iterations
            print("  slot", slot, "n", n, "hist(>=20)", np.bincount(np.minimum(it, 125))[20:].nonzero()[0] + 20, "max", it.max(), "count it>=40:", int((it >= 40).sum()))
            big = np.nonzero(it >= 40)[0][:6]
            for j in big:
                kk = k[t.k_index[j]]
                print("     k %.3f W %.5f..%.5f omega/k %.6f it %d acc %d ext %.3e int %.3e U-range W-U: %.4f" % (kk, W[t.w_index[j]], W[t.w_index[j]+1], t.omega[j]/kk, it[j], t.accepted[j], t.ext[j], t.intq[j], t.omega[j]/kk - U))
