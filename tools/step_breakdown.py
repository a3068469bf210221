"""Wall / device time of the pieces of one Picard iteration of the 3-D INS-DIM workload (run on the GPU box):
   python tools/step_breakdown.py [N]"""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault('OCMP_PATCH_STORAGE', 'fp32'); os.environ.setdefault('OCMP_SPMV_FP32', '1')
import numpy as np, torch
import opencmp_b200.ngs as ngs
from opencmp_b200.backend import CudaBackend
from opencmp_b200.workloads import INSSphereDIM3D
be = CudaBackend(0); ngs.set_backend(be)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 48
w = INSSphereDIM3D(N, nu=1.0, linear_tolerance=1e-12, periodic=(False,) * 3, nonlinear_max_iterations=2,
                   nonlinear_tolerance=(0.0, 0.0), wall_period=0.1)
for _ in range(3): w.step()
acc = {}
def timed(name, fn):
    torch.cuda.synchronize(); t0 = time.perf_counter(); out = fn(); torch.cuda.synchronize()
    acc[name] = acc.get(name, 0.0) + time.perf_counter() - t0
    return out
steps = 3
u_comp = w.gfu.components[0]
for _ in range(steps):
    w.t.Set(w.t.Get() + w.dt.Get())
    for it in range(2):
        timed('dirichlet Set', w.apply_dirichlet_bcs)
        timed('a.Assemble', w.a.Assemble)
        timed('L.Assemble', w.L.Assemble)
        timed('pre.Update', w.pre.Update)
        timed('linear_solve', w.linear_solve)
        def norms():
            diff = w.W - u_comp
            a = w._integrate(ngs.InnerProduct(diff, diff)); b = w._integrate(ngs.InnerProduct(u_comp, u_comp))
            return a, b
        timed('two Integrate norms', norms)
        def shift():
            w.W.vec.data = u_comp.vec
        timed('wind update', shift)
    def hist():
        w.gfu_0.vec.data = w.gfu.vec
    timed('history shift', hist)
tot = sum(acc.values())
print(json.dumps({'N': N, 's_per_step': tot / steps, 'its': w.linear_iterations[-2:],
                  'ms_per_step': {k: round(v / steps * 1e3, 2) for k, v in acc.items()}}))
