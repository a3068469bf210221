import sys, time, os
sys.path.insert(0,'/root/repo')
import numpy as np, torch
import opencmp_b200.ngs as ngs
from opencmp_b200.backend import CudaBackend
be = CudaBackend(); ngs.set_backend(be)
from opencmp_b200.workloads import INSTaylorGreen
N = int(sys.argv[1]) if len(sys.argv) > 1 else 128
w = INSTaylorGreen(N, order=3, preconditioner='multigrid')
w.t.Set(1e-3); w.apply_dirichlet_bcs(); w.assemble()
st = w.pre.state
lib = be.lib
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0,e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/n
stream = be._stream()
for l in range(st.nlevels-1, 0, -1):
    sm = st.smoothers[l]; n = st.spaces[l].ndof
    r = torch.rand(n, dtype=torch.float64, device='cuda'); z = torch.zeros_like(r)
    esz = {'fp64': 8.0, 'fp32': 4.0, 'bf16': 2.0}[sm.storage]
    f_apply = {'fp64': lib.ocmp_asm_apply, 'fp32': lib.ocmp_asm_apply_f32, 'bf16': lib.ocmp_asm_apply_bf16}[sm.storage]
    f_setup = {'fp64': lib.ocmp_asm_setup, 'fp32': lib.ocmp_asm_setup_f32, 'bf16': lib.ocmp_asm_setup_bf16}[sm.storage]
    ms = timeit(lambda: f_apply(sm.npatch, sm.bs, sm.pdofs.data_ptr(), sm.inv.data_ptr(), sm.inc_ptr.data_ptr(), sm.inc_idx.data_ptr(), sm.ybuf.data_ptr(), r.data_ptr(), z.data_ptr(), n, stream))
    by = esz*sm.npatch*sm.bs*sm.bs + 16.0*n
    mat = w.a.mat if l == st.nlevels-1 else st.mats[l]
    pd = be.pattern_data(st.spaces[l])
    mi = timeit(lambda: f_setup(sm.npatch, sm.bs, sm.pdofs.data_ptr(), pd['rowptr'].data_ptr(), pd['colidx'].data_ptr(), mat.values.data_ptr(), st.masks[l].data_ptr(), sm.inv.data_ptr(), be._patches(st.spaces[l], 'vertex')['pos'].data_ptr(), stream), n=5)
    y = torch.zeros_like(r)
    msp = timeit(lambda: lib.ocmp_spmv(n, pd['rowptr'].data_ptr(), pd['colidx'].data_ptr(), mat.values.data_ptr(), r.data_ptr(), y.data_ptr(), stream))
    bsp = pd['nnz']*12 + n*20
    print(sm.storage, 'level %d: npatch %d bs %d | apply %.3f ms %.0f GB/s | invert %.3f ms (%.1f GFLOP/s) | spmv %.3f ms %.0f GB/s' % (l, sm.npatch, sm.bs, ms, by/ms/1e6, mi, 2.0*sm.npatch*sm.bs**3/mi/1e6, msp, bsp/msp/1e6), flush=True)
ta = timeit(lambda: w.a.Assemble(), n=5); print('a.Assemble fine: %.3f ms  -> %.0f Mnnz/s' % (ta, w.nnz/ta/1e3))
tl = timeit(lambda: w.L.Assemble(), n=5); print('L.Assemble: %.3f ms' % tl)
tb = timeit(lambda: w.apply_dirichlet_bcs(), n=5); print('dirichlet Set: %.3f ms' % tb)
tp = timeit(lambda: w.pre.Update(), n=3); print('pre.Update (all levels): %.3f ms' % tp)
