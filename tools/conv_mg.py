import sys, time, os
sys.path.insert(0,'/root/repo')
import numpy as np, torch
import opencmp_b200.ngs as ngs
from opencmp_b200.backend import CudaBackend
be = CudaBackend(); ngs.set_backend(be)
from opencmp_b200.workloads import INSTaylorGreen
for N in (32, 64, 128, 256):
    t0=time.time()
    w = INSTaylorGreen(N, order=3, preconditioner='multigrid')
    torch.cuda.synchronize(); ts=time.time()-t0
    w.t.Set(w.t.Get()+w.dt.Get()); w.apply_dirichlet_bcs()
    torch.cuda.synchronize(); t=time.time(); w.assemble(); torch.cuda.synchronize(); ta=time.time()-t
    t=time.time(); w.assemble(); torch.cuda.synchronize(); ta2=time.time()-t
    for tol in (1e-10,):
        x = w.gfu.vec.Copy()
        torch.cuda.synchronize(); t=time.time()
        ngs.solvers.GMRes(A=w.a.mat, b=w.L.vec, pre=w.pre, freedofs=w.fes.FreeDofs(), x=x, tol=tol, maxsteps=300, restart=100)
        torch.cuda.synchronize()
        print('N',N,'ndof',w.ndof,'setup %.1fs'%ts,'assemble+precond %.3fs / %.3fs'%(ta,ta2),'tol',tol,'its',be.last_iters,'res %.2e'%be.last_resid,'solve %.3fs'%(time.time()-t), flush=True)
    if N <= 32:
        # compare with plain asm solution
        x2 = w.gfu.vec.Copy()
        pre2 = ngs.Preconditioner(w.a, 'asm'); pre2.Update()
        ngs.solvers.GMRes(A=w.a.mat, b=w.L.vec, pre=pre2, freedofs=w.fes.FreeDofs(), x=x2, tol=1e-12, maxsteps=1000, restart=1000)
        print('   diff vs asm solve', float((x.a-x2.a).abs().max()/x2.a.abs().max()), 'asm its', be.last_iters, flush=True)
