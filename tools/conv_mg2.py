import sys, time, os
sys.path.insert(0,'/root/repo')
import numpy as np, torch
import opencmp_b200.ngs as ngs
from opencmp_b200.backend import CudaBackend
be = CudaBackend(); ngs.set_backend(be)
from opencmp_b200.workloads import INSTaylorGreen
N = int(sys.argv[1])
w = INSTaylorGreen(N, order=3, preconditioner='multigrid')
w.t.Set(w.t.Get()+w.dt.Get()); w.apply_dirichlet_bcs(); w.assemble()
x = w.gfu.vec.Copy()
torch.cuda.synchronize(); t=time.time()
ngs.solvers.GMRes(A=w.a.mat, b=w.L.vec, pre=w.pre, freedofs=w.fes.FreeDofs(), x=x, tol=1e-10, maxsteps=200, restart=100)
torch.cuda.synchronize()
print('N',N,'N0',os.environ.get('OCMP_MG_N0'),'nu',os.environ.get('OCMP_MG_NU'),'omega',os.environ.get('OCMP_MG_OMEGA'),'levels',w.pre.state.nlevels,'its',be.last_iters,'res %.2e'%be.last_resid,'solve %.3fs'%(time.time()-t), flush=True)
