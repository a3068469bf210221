"""GMRES iteration counts of the 3-D INS-DIM workload when the BOX grows with the mesh at fixed cell size (single GPU):
   python tools/domain_study.py 24   -> (N=24, half-width 1), (N=36, 1.5), (N=48, 2): h = 2/24 throughout"""
import os, sys, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import opencmp_b200.ngs as ngs
from opencmp_b200.backend import CudaBackend
from opencmp_b200.mesh import structured_3d
from opencmp_b200.workloads import INSSphereDIM3D
be = CudaBackend(0); ngs.set_backend(be)
N0 = int(sys.argv[1]) if len(sys.argv) > 1 else 24
for N, half in ((N0, 1.0), (N0 * 3 // 2, 1.5), (2 * N0, 2.0)):
    n, k = N, 0
    while n % 2 == 0 and n > 2:
        n //= 2; k += 1
    mesh = structured_3d([n] * 3, scale=(2.0 * half,) * 3, offset=(half,) * 3)
    for _ in range(k):
        mesh.Refine()
    w = INSSphereDIM3D(N, mesh=mesh, nu=1.0, linear_tolerance=1e-12, periodic=(False,) * 3, wall_period=0.1,
                       nonlinear_max_iterations=2, nonlinear_tolerance=(0.0, 0.0))
    w.step(); t0 = time.time(); w.step(); w.step(); torch.cuda.synchronize()
    print(json.dumps({'N': N, 'half_width': half, 'coarsest': n, 'dofs': w.ndof, 'its': w.linear_iterations,
                      's_per_step': (time.time() - t0) / 2, 'err': w.errors()[0]}), flush=True)
    del w
    torch.cuda.empty_cache()
