"""One assembled 3-D INS-DIM system (default N = 32) and a few operator / smoother / assembly calls — the target of the
ncu --set full captures of round 2 (run on the GPU box): python tools/kern_bench3d.py [N]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault('OCMP_PATCH_STORAGE', 'fp32')
os.environ.setdefault('OCMP_SPMV_FP32', '1')
import torch
import opencmp_b200.ngs as ngs
from opencmp_b200.backend import CudaBackend
from opencmp_b200.workloads import INSSphereDIM3D
be = CudaBackend(0); ngs.set_backend(be)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 32
w = INSSphereDIM3D(N, nu=1.0, linear_tolerance=1e-12, periodic=(False,) * 3, nonlinear_max_iterations=1,
                   nonlinear_tolerance=(0.0, 0.0), wall_period=0.1)
w.step()
def timeit(fn, n=10):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
x = w.gfu.vec.CreateVector(); x.data = w.gfu.vec
y = w.gfu.vec.CreateVector()
ms = timeit(lambda: w.a.mat.Mult(x, y))
by = w.nnz * 12 + w.ndof * 20
print('N %d dofs %d nnz %d | mat*x %.3f ms = %.0f GB/s (CSR bytes nnz*12 + n*20)' % (N, w.ndof, w.nnz, ms, by / ms / 1e6))
os.environ['OCMP_SPMV_GROUPED'] = '0'
print('a.Assemble %.3f ms -> %.0f Mnnz/s ; L.Assemble %.3f ms ; pre.Update %.3f ms' % (
    (ta := timeit(lambda: w.a.Assemble(), 5)), w.nnz / ta / 1e3, timeit(lambda: w.L.Assemble(), 5), timeit(lambda: w.pre.Update(), 3)))
w.step()
print('step ok, its', w.linear_iterations[-2:])
