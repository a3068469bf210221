"""Times the band LU behind mat.Inverse on the INS Taylor-Green workload (HDiv-DG order 3) at a few mesh sizes:
   python tools/direct_bench.py 16 32 48      (run on the GPU box)"""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, '.')
import opencmp_b200.ngs as ngs
from opencmp_b200.backend import CudaBackend
from opencmp_b200.direct import BandLU
from opencmp_b200.workloads import INSTaylorGreen

be = CudaBackend()
ngs.set_backend(be)
for N in [int(a) for a in sys.argv[1:]] or [16, 32]:
    w = INSTaylorGreen(N, order=3, linear_solver='direct', preconditioner=None)
    w.apply_dirichlet_bcs()
    w.assemble()
    t0 = time.perf_counter()
    lu = BandLU(be, w.a.mat, w.fes.FreeDofs())
    torch.cuda.synchronize()
    t_first = time.perf_counter() - t0
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    ev[0].record()
    lu.Update()
    ev[1].record()
    r = w.L.vec.a
    out = be.zeros(w.fes.ndof)
    ev[2].record()
    lu.solve(r, out, refine=0)
    ev[3].record()
    torch.cuda.synchronize()
    res = be.zeros(w.fes.ndof)
    be.spmv(w.a.mat, out, res)
    fm = lu.fm
    rel0 = float(((r - res) * fm).norm() / (r * fm).norm())
    lu.solve(r, out, refine=2)
    be.spmv(w.a.mat, out, res)
    rel2 = float(((r - res) * fm).norm() / (r * fm).norm())
    print(dict(N=N, ndof=w.fes.ndof, n=lu.n, kl=lu.kl, ku=lu.ku, ubw=lu.ubw, band_gb=8 * lu.length / 2 ** 30,
               first_s=round(t_first, 3), factor_ms=ev[0].elapsed_time(ev[1]), solve_ms=ev[2].elapsed_time(ev[3]),
               gflops=2e-6 * lu.n * lu.kl * (lu.kl + lu.ku) / ev[0].elapsed_time(ev[1]),
               resid_no_refine=rel0, resid_refined=rel2), flush=True)
