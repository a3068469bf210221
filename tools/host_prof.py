import sys, time, os, cProfile, pstats
sys.path.insert(0,'/root/repo')
import numpy as np, torch
import opencmp_b200.ngs as ngs
from opencmp_b200.backend import CudaBackend
be = CudaBackend(); ngs.set_backend(be)
from opencmp_b200.workloads import INSTaylorGreen
w = INSTaylorGreen(128, order=3, preconditioner='multigrid')
for _ in range(2): w.step()
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
t=time.perf_counter()
for _ in range(3): w.step()
torch.cuda.synchronize()
print('wall per step', (time.perf_counter()-t)/3)
pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(28)
