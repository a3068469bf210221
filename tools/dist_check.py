import os, sys
sys.path.insert(0, '/root/repo')
import numpy as np, torch, torch.distributed as dist
world = int(os.environ.get('WORLD_SIZE', '1')); rank = int(os.environ.get('RANK', '0')); local = int(os.environ.get('LOCAL_RANK', '0'))
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
import opencmp_b200.ngs as ngs
from opencmp_b200.backend import CudaBackend
be = CudaBackend(local); ngs.set_backend(be)
from opencmp_b200.dist_workload import DistributedINS
from opencmp_b200.workloads import INSTaylorGreen
N = int(sys.argv[1]); order = int(sys.argv[2])
d = DistributedINS(N, world, rank, order=order, n0=4, replicate_below=int(os.environ.get('REPL', '100000')))
g = INSTaylorGreen(N, order=order, mesh=d.gmesh, preconditioner='multigrid')
top = d.mg.levels[-1].map
l2g = torch.from_numpy(top.l2g).cuda(); vl2g = torch.from_numpy(d._vmap.l2g).cuda()
d.w.gfu.vec.a.copy_(g.gfu.vec.a[l2g]); d.w.gfu_0.vec.a.copy_(g.gfu_0.vec.a[l2g]); d.w.W.vec.a.copy_(g.W.vec.a[vl2g])
# compare operators after first assembly
for w in (d.w, g):
    w.t.Set(w.t.Get() + w.dt.Get()); w.apply_dirichlet_bcs(); w.assemble()
owned = torch.from_numpy(top.owned).cuda()
def rel(a, b): return float((a - b).abs().max() / b.abs().max())
print(rank, 'bc', rel(d.w.gfu.vec.a, g.gfu.vec.a[l2g]), 'rhs(owned)', rel(d.w.L.vec.a[owned], g.L.vec.a[l2g][owned]), flush=True)
xg = torch.rand(g.ndof, dtype=torch.float64, device='cuda', generator=torch.Generator(device='cuda').manual_seed(1))
yg = torch.zeros_like(xg); be.spmv(g.a.mat, xg, yg)
yl = d.mg.mult(len(d.mg.levels) - 1, xg[l2g].clone())
print(rank, 'spmv', rel(yl, yg[l2g]), flush=True)
# smoother on the finest level vs global smoother
st = g.pre.state
r = xg * st.masks[-1]
zg = torch.zeros_like(r)
sm = st.smoothers[-1]
be.lib.ocmp_asm_apply(sm.npatch, sm.bs, sm.pdofs.data_ptr(), sm.inv.data_ptr(), sm.inc_ptr.data_ptr(), sm.inc_idx.data_ptr(), sm.ybuf.data_ptr(), r.data_ptr(), zg.data_ptr(), zg.numel(), be._stream())
zg = zg * sm.wgt * st.masks[-1]
zl = d.mg.smooth(len(d.mg.levels) - 1, r[l2g].clone())
print(rank, 'smooth', rel(zl, zg[l2g]), flush=True)
# coarse level matrices: compare spmv on level L-1
for l in range(len(d.mg.levels) - 1, -1, -1):
    lv = d.mg.levels[l]
    n_g = lv.map.nglobal
    xg2 = torch.rand(n_g, dtype=torch.float64, device='cuda', generator=torch.Generator(device='cuda').manual_seed(2))
    ll2g = torch.from_numpy(lv.map.l2g).cuda()
    yl2 = d.mg.mult(l, xg2[ll2g].clone())
    mat_g = g.a.mat if l == len(d.mg.levels) - 1 else st.mats[l]
    yg2 = torch.zeros(n_g, dtype=torch.float64, device='cuda'); be.spmv(mat_g, xg2, yg2)
    print(rank, 'level', l, 'spmv', rel(yl2, yg2[ll2g]), flush=True)
# full V-cycle comparison
lt = len(d.mg.levels) - 1
zl = d.mg.vcycle(lt, r[l2g].clone())
zg2 = torch.zeros_like(r)
import ctypes as C
from opencmp_b200.backend import System
sys_ = be._system(g.a.mat, st.fm, st)
work = torch.empty(be.lib.ocmp_krylov_work_len(g.ndof, 2, 1), dtype=torch.float64, device='cuda')
# one Richardson step from x=0 with damping 1: x = P r
it, res = C.c_int(0), C.c_double(0.0)
be.lib.ocmp_krylov(C.byref(sys_), 2, r.data_ptr(), zg2.data_ptr(), 0.0, 1, 1, 1.0, work.data_ptr(), work.numel(), C.byref(it), C.byref(res), be._stream())
print(rank, 'vcycle', rel(zl, zg2[l2g]), flush=True)
for s_ in range(1):
    d.w.t.Set(g.t.Get() - g.dt.Get()); g.t.Set(g.t.Get() - g.dt.Get())
    d.w.linear_iterations = []; g.linear_iterations = []
    d.step(); g.step()
    print(rank, 'its', d.w.linear_iterations, g.linear_iterations, 'sol', rel(d.w.gfu.vec.a[:d.w.V.ndof], g.gfu.vec.a[l2g][:d.w.V.ndof]), flush=True)
dist.destroy_process_group()
