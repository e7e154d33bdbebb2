"""Host study (no GPU; the kernels' arithmetic compiled for the CPU, tests/host_kernel.py): accuracy of one kind's
scan against a 608-step solution as a function of n_steps - D itself relative to max(|ext|, |int|) (maximum, 99.9 %
quantile, points above 1e-9: the ones next to a pole of D) and the guard's projective measure g - per mode, on 60 k x
600 omega random points of the BASELINE window outside the continua.
Usage: python scripts/host_steps_study.py cylinder_density 96,112,128,136,144,152,176  (profiles/r02ag_steps_study_host.log)"""
import os, sys
_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, _ROOT); sys.path.insert(0, os.path.join(_ROOT, 'tests'))
import numpy as np, host_kernel as hk, time
import eigensolver_b200 as esb
from helpers import CASES
from test_host_kernel import spec_of
name=sys.argv[1]; steps=[int(x) for x in sys.argv[2].split(',')]
case=CASES[name]
rng=np.random.default_rng(7)
if case.kind=="cylinder_rotation": k=np.sort(rng.uniform(0.25,4.0,60))
else: k=np.sort(np.concatenate([rng.uniform(0.01,0.05,6),rng.uniform(0.05,4.5,54)]))
W=np.sort(rng.uniform(case.W[0],case.W[1],600))
modes=list(case.modes)+([3] if case.kind.startswith("cylinder") else [])
sp=spec_of(case)
kw=sp.solver_kwargs(); kw["n_steps"]=608 if case.kind!="cylinder_rotation" else 512
ref=esb.ModelSpec(**kw)
E1,I1,D1=hk.grid(ref,modes,k,W)
reg=np.array([case.regular(k,W,m) for m in modes])
for n in steps:
    kw=sp.solver_kwargs(); kw["n_steps"]=n
    s=esb.ModelSpec(**kw)
    E0,I0,D0=hk.grid(s,modes,k,W)
    out=[]
    for j,m in enumerate(modes):
        ok=reg[j]&np.isfinite(E1[j])&np.isfinite(I1[j])&np.isfinite(E0[j])&np.isfinite(I0[j])
        dev=(np.abs((E0[j]-I0[j])-(E1[j]-I1[j]))/np.maximum(np.abs(E1[j]),np.abs(I1[j])))[ok]
        g0=(E0[j]-I0[j])*D0[j]/(np.abs(E0[j]*D0[j])+np.abs(I0[j]*D0[j])); g1=(E1[j]-I1[j])*D1[j]/(np.abs(E1[j]*D1[j])+np.abs(I1[j]*D1[j]))
        gd=np.abs(g0-g1)[ok]
        out.append("m%d: max %.1e q999 %.1e n>1e-9 %d | g max %.1e"%(m,dev.max(),np.quantile(dev,0.999),(dev>1e-9).sum(),gd.max()))
    print(n, s.scheme, " ; ".join(out), flush=True)
