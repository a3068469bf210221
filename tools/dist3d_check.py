"""Element-partitioned 3-D INS-DIM step: iteration counts of the C driver (ocmp_krylov, native) vs the Python cycle.
   torchrun --nproc-per-node 2 tools/dist3d_check.py 24 sphere"""
import os, sys, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
world = int(os.environ.get('WORLD_SIZE', '1')); rank = int(os.environ.get('RANK', '0')); local = int(os.environ.get('LOCAL_RANK', '0'))
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
import opencmp_b200.ngs as ngs
from opencmp_b200.backend import CudaBackend
be = CudaBackend(local); ngs.set_backend(be)
from opencmp_b200.dist_workload import DistributedINSDIM3D
N = int(sys.argv[1]); layout = sys.argv[2] if len(sys.argv) > 2 else 'sphere'
for native in ('1', '0'):
    os.environ['OCMP_DIST_NATIVE'] = native
    d = DistributedINSDIM3D(N, world, rank, layout='sphere' if layout == 'sphere' else None, linear_max_iterations=150)
    t0 = time.time(); d.step(); torch.cuda.synchronize()
    if rank == 0:
        print(json.dumps({'native': native, 'layout': layout, 'N': N, 'its': d.w.linear_iterations, 's': time.time() - t0,
                          'global_dofs': d.ndof_global}), flush=True)
    del d
    torch.cuda.empty_cache()
dist.destroy_process_group()
