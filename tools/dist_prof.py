import os, sys, json, time, cProfile, pstats
sys.path.insert(0, '/root/repo')
import numpy as np, torch, torch.distributed as dist
world = int(os.environ.get('WORLD_SIZE', '1')); rank = int(os.environ.get('RANK', '0')); local = int(os.environ.get('LOCAL_RANK', '0'))
torch.cuda.set_device(local)
if world > 1: dist.init_process_group('nccl', device_id=torch.device('cuda', local))
import opencmp_b200.ngs as ngs
from opencmp_b200.backend import CudaBackend
ngs.set_backend(CudaBackend(local))
from opencmp_b200.dist_workload import DistributedINS
d = DistributedINS(128, world, rank, order=3)
for _ in range(2): d.step()
torch.cuda.synchronize()
pr = cProfile.Profile()
if rank == 0: pr.enable()
t0 = time.perf_counter()
for _ in range(3): d.step()
torch.cuda.synchronize()
if rank == 0:
    pr.disable()
    print('wall/step', (time.perf_counter() - t0) / 3)
    pstats.Stats(pr).sort_stats('tottime').print_stats(22)
if world > 1: dist.destroy_process_group()
