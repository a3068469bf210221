import os, sys, json
sys.path.insert(0, '/root/repo')
import numpy as np, torch, torch.distributed as dist
world = int(os.environ.get('WORLD_SIZE', '1')); rank = int(os.environ.get('RANK', '0')); local = int(os.environ.get('LOCAL_RANK', '0'))
torch.cuda.set_device(local)
if world > 1: dist.init_process_group('nccl', device_id=torch.device('cuda', local))
import opencmp_b200.ngs as ngs
from opencmp_b200.backend import CudaBackend
ngs.set_backend(CudaBackend(local))
from opencmp_b200.dist_workload import DistributedPoisson
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
w = DistributedPoisson(n, 2, world, rank)
ms = w.time_spmv()
t = torch.tensor([ms], dtype=torch.float64, device='cuda')
b = torch.tensor([float(w.spmv_bytes_owned())], dtype=torch.float64, device='cuda')
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX); dist.all_reduce(b)
its, res, sec, x = w.solve(maxit=100)
# check against a few known properties: residual reduction
if rank == 0:
    print(json.dumps({'world': world, 'n_per_rank': n, 'global_dofs': w.nglobal, 'spmv_ms_max': float(t[0]), 'spmv_gbs_total': float(b[0]) / float(t[0]) / 1e6, 'cg_its': its, 'cg_res': res, 'cg_ms_per_it': sec / max(1, its) * 1e3}))
if world > 1: dist.destroy_process_group()
