import os, sys, json, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch, torch.distributed as dist
world = int(os.environ.get('WORLD_SIZE', '1')); rank = int(os.environ.get('RANK', '0')); local = int(os.environ.get('LOCAL_RANK', '0'))
torch.cuda.set_device(local)
if world > 1: dist.init_process_group('nccl', device_id=torch.device('cuda', local))
import opencmp_b200.ngs as ngs
from opencmp_b200.backend import CudaBackend
ngs.set_backend(CudaBackend(local))
from opencmp_b200.dist_workload import DistributedINS
N = int(sys.argv[1]) if len(sys.argv) > 1 else 64
t0 = time.time()
d = DistributedINS(N, world, rank, order=3, strips=int(os.environ.get('STRIPS', world)))
torch.cuda.synchronize(); ts = time.time() - t0
for _ in range(2): d.step()
torch.cuda.synchronize()
if world > 1: dist.barrier()
t0 = time.time()
K = 3
its = 0
for _ in range(K):
    d.w.linear_iterations = []
    d.step(); its += sum(d.w.linear_iterations)
torch.cuda.synchronize()
if world > 1: dist.barrier()
sec = (time.time() - t0) / K
eu, ep = d.w.errors()
chk = float(np.sqrt(d.w._integrate(ngs.InnerProduct(d.w.gfu.components[0], d.w.gfu.components[0]))))
if rank == 0:
    print(json.dumps({'world': world, 'N_per_rank': N, 'global_dofs': d.ndof_global, 'local_dofs': d.w.ndof, 's_per_step': sec, 'gmres_its_per_step': its / K, 'picard': d.w.picard_iterations, 'err_u': eu, 'err_p': ep, 'u_norm': repr(chk), 'native': os.environ.get('OCMP_DIST_NATIVE', '1'), 'setup_s': ts}))
if world > 1: dist.destroy_process_group()
