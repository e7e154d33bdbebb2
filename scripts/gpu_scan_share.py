"""One rank's share (k[0::world]) of the configs[4] density family, for a launch list under ncu."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import eigensolver_b200 as esb
from eigensolver_b200.scan import density_flow_grid, parameter_scan

world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4
dens, flow = density_flow_grid(np.linspace(0.1, 0.4, 20), np.linspace(0.05, 0.9, 20))
k = np.linspace(0.01, 4.5, 1000)
Wd = np.linspace(0.5, 5.0, 10000)
with esb.DispersionSolver("cylinder_density") as sd:
    for rep in range(2):
        r1 = parameter_scan(sd, dens[:n], k, Wd, [0, 1, 2], rank=0, world=world, download=False)
        torch.cuda.synchronize()
    print(r1.points[0].n_brackets)
