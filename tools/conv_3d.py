"""Convergence history of the multigrid-GMRES solves of the 3-D INS-DIM workload (run on the GPU box)."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import opencmp_b200.ngs as ngs
from opencmp_b200.backend import CudaBackend
be = CudaBackend(0); ngs.set_backend(be)
from opencmp_b200.workloads import INSSphereDIM3D
N = int(sys.argv[1]); tol = float(sys.argv[2]) if len(sys.argv) > 2 else 1e-12
maxit = int(os.environ.get('MAXIT', '150'))
t0 = time.time()
w = INSSphereDIM3D(N, linear_tolerance=tol, linear_max_iterations=maxit, n0=int(os.environ.get('N0', '2')),
                   lam=float(os.environ.get('LAM', '0.25')))
torch.cuda.synchronize(); ts = time.time() - t0
hist = []
orig = w.linear_solve
def solve():
    orig(); hist.append(be.krylov_history())
w.linear_solve = solve
for step in range(int(os.environ.get('STEPS', '2'))):
    t0 = time.time(); w.step(); torch.cuda.synchronize(); dt = time.time() - t0
    print(json.dumps({'N': N, 'tol': tol, 'step': step, 's': dt, 'picard': w.picard_iterations, 'its': w.linear_iterations[-w.picard_iterations:],
                      'err': w.errors()[0], 'setup_s': ts, 'mem_GB': torch.cuda.max_memory_allocated() / 1e9,
                      'hist': [['%.1e' % v for v in h[::10]] for h in hist[-w.picard_iterations:]]}), flush=True)
