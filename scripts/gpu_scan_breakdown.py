"""Exploration (not a test): the configs[4] job as rank 0 of N sees it (k[0::N]) on one GPU, N = 1, 2, 4, 8:
time of the whole job and of its host-side parts - what strong scaling can reach."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import eigensolver_b200 as esb
from eigensolver_b200.scan import density_flow_grid, parameter_scan

dens, flow = density_flow_grid(np.linspace(0.1, 0.4, 20), np.linspace(0.05, 0.9, 20))
k = np.linspace(0.01, 4.5, 1000)
Wd = np.linspace(0.5, 5.0, 10000); Wf = np.linspace(-2.7, 2.7, 10000)
with esb.DispersionSolver("cylinder_density") as sd, esb.DispersionSolver("slab_flow") as sf:
    for world in (1, 2, 4, 8):
        for rep in range(2):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            r1 = parameter_scan(sd, dens, k, Wd, [0, 1, 2], rank=0, world=world, download=False)
            torch.cuda.synchronize(); t1 = time.perf_counter()
            r2 = parameter_scan(sf, flow, k, Wf, [0, 1], rank=0, world=world, download=False)
            torch.cuda.synchronize(); t2 = time.perf_counter()
        # the two families from two host threads (two contexts, two streams): the latency-bound tail of one
        # family's refinement overlaps the other's scan
        import threading
        for rep in range(2):
            torch.cuda.synchronize()
            t5 = time.perf_counter()
            th = [threading.Thread(target=parameter_scan, args=(sd, dens, k, Wd, [0, 1, 2]), kwargs=dict(rank=0, world=world, download=False)),
                  threading.Thread(target=parameter_scan, args=(sf, flow, k, Wf, [0, 1]), kwargs=dict(rank=0, world=world, download=False))]
            [t.start() for t in th]; [t.join() for t in th]
            torch.cuda.synchronize(); t6 = time.perf_counter()
        print("      two threads: total %.1f ms" % (1e3 * (t6 - t5)), flush=True)
        # host-side share: building the 20 specs + sampling (python) measured alone
        t3 = time.perf_counter()
        import copy
        for p in dens:
            sp = copy.copy(sd.spec); sp.model = type(sd.spec.model).from_buffer_copy(sd.spec.model)
            sp.replace(p.get("medium"), p.get("profile")); sp.sampled()
        t4 = time.perf_counter()
        print("N=%d: density %.1f ms, flow %.1f ms, total %.1f ms (ideal %.1f); python model set-up of 20 equilibria %.1f ms"
              % (world, 1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t2 - t0), 1316.0 / world, 1e3 * (t4 - t3)), flush=True)
