"""GMRES iteration counts of the 3-D INS-DIM workload on anisotropically refined boxes (single GPU):
   python tools/aniso_study.py 24   -> meshes (24,24,24), (48,24,24), (48,48,24), (48,48,48) cells"""
import os, sys, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import opencmp_b200.ngs as ngs
from opencmp_b200.backend import CudaBackend
from opencmp_b200.mesh import structured_3d
from opencmp_b200.workloads import INSSphereDIM3D
be = CudaBackend(0); ngs.set_backend(be)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 24
for grid in ((1, 1, 1), (2, 1, 1), (2, 2, 1), (2, 2, 2)):
    n, k = N, 0
    while n % 2 == 0 and n > 2:
        n //= 2; k += 1
    mesh = structured_3d([n * g for g in grid], scale=(2.0,) * 3, offset=(1.0,) * 3)
    for _ in range(k):
        mesh.Refine()
    w = INSSphereDIM3D(N, mesh=mesh, nu=1.0, linear_tolerance=1e-12, periodic=(False,) * 3)
    t0 = time.time(); w.step(); w.step(); torch.cuda.synchronize()
    print(json.dumps({'cells': [N * g for g in grid], 'dofs': w.ndof, 'its': w.linear_iterations, 's_per_step': (time.time() - t0) / 2,
                      'err': w.errors()[0]}), flush=True)
    del w
    torch.cuda.empty_cache()
