import sys, time
sys.path.insert(0,'/root/repo')
import numpy as np, torch
import opencmp_b200.ngs as ngs
from opencmp_b200.backend import CudaBackend
be = CudaBackend(); ngs.set_backend(be)
from opencmp_b200.workloads import INSTaylorGreen
for N in (16, 32, 64, 128):
  for pre in ("asm",):
    w = INSTaylorGreen(N, order=3, preconditioner=pre)
    w.t.Set(w.t.Get()+w.dt.Get()); w.apply_dirichlet_bcs(); w.assemble()
    for restart in (1000,):
        for tol in (1e-10,):
            x = w.gfu.vec.Copy()
            torch.cuda.synchronize(); t=time.time()
            ngs.solvers.GMRes(A=w.a.mat, b=w.L.vec, pre=w.pre, freedofs=w.fes.FreeDofs(), x=x, tol=tol, maxsteps=1000, restart=restart)
            torch.cuda.synchronize()
            print('N',N,pre,'restart',restart,'tol',tol,'its',be.last_iters,'res %.2e'%be.last_resid,'%.2fs'%(time.time()-t), flush=True)
