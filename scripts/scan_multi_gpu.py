"""BASELINE configs[4] as stated: a parameter-space scan of 1e9 D evaluations sharded over the GPUs of one
box.  20 density contrasts (cylinder, n = 0,1,2, 1000 k x 10000 omega = 3e7 evaluations each) + 20 flow
amplitudes (slab, sausage + kink, 1000 k x 10000 omega = 2e7 each) = 1e9.  The list of equilibria shards
across ranks (no data-path collective); the per-equilibrium mode counts are gathered at the end.

    python scripts/scan_multi_gpu.py                                   # one GPU
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 scripts/scan_multi_gpu.py
"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import eigensolver_b200 as esb
from eigensolver_b200.scan import density_flow_grid, parameter_scan

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
dens, flow = density_flow_grid(np.linspace(0.12, 0.35, 20), np.linspace(0.05, 0.9, 20))
k = np.linspace(0.01, 4.5, 1000)
Wd, Wf = np.linspace(0.5, 5.0, 10000), np.linspace(-2.7, 2.7, 10000)
with esb.DispersionSolver("cylinder_density", device=local) as sd, esb.DispersionSolver("slab_flow", device=local) as sf:
    parameter_scan(sd, dens[:1], k[:64], Wd, [0, 1, 2]); parameter_scan(sf, flow[:1], k[:64], Wf, [0, 1])   # warm-up
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t = time.perf_counter()
    # round-robin shares, the two families started half a turn apart: when the lists do not divide by the
    # number of ranks, the ranks that get an extra cylinder equilibrium are not the ones that get an extra slab
    mine = parameter_scan(sd, dens[rank::world], k, Wd, [0, 1, 2]) + \
        parameter_scan(sf, flow[(rank + world // 2) % world::world], k, Wf, [0, 1])
    torch.cuda.synchronize()
    summary = [(p.label, p.n_brackets, p.n_modes) for p in mine]
    if world > 1:
        allp = [None] * world
        dist.all_gather_object(allp, summary)
        summary = [x for part in allp for x in part]
        dist.barrier()
    dt = time.perf_counter() - t
if rank == 0:
    evals = len(dens) * 3 * len(k) * len(Wd) + len(flow) * 2 * len(k) * len(Wf)
    print(json.dumps({"config": "configs[4]: 20 density contrasts x (n=0,1,2) + 20 flow amplitudes x (sausage, kink), "
                                "1000 k x 10000 omega each", "n_gpus": world, "evals": evals, "seconds": dt,
                      "evals_per_sec": evals / dt, "equilibria": len(summary),
                      "modes_found": int(sum(sum(s[2]) for s in summary))}))
if world > 1:
    dist.destroy_process_group()
