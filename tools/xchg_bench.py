import os, sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch, torch.distributed as dist
world = int(os.environ.get('WORLD_SIZE', '1')); rank = int(os.environ.get('RANK', '0')); local = int(os.environ.get('LOCAL_RANK', '0'))
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
import opencmp_b200.ngs as ngs
from opencmp_b200.backend import CudaBackend
ngs.set_backend(CudaBackend(local))
from opencmp_b200.dist_workload import DistributedINS
d = DistributedINS(128, world, rank, order=3)
for l, lv in enumerate(d.mg.levels):
    if lv.replicated: continue
    x = torch.rand(lv.n, dtype=torch.float64, device='cuda')
    for name, fn in (('exchange', lambda: lv.map.exchange(x)), ('exchange_sum', lambda: lv.map.exchange_sum(x))):
        for _ in range(20): fn()
        torch.cuda.synchronize(); dist.barrier()
        t0 = time.perf_counter()
        for _ in range(200): fn()
        tc = time.perf_counter() - t0
        torch.cuda.synchronize()
        tg = time.perf_counter() - t0
        if rank == 0: print('level', l, 'n', lv.n, name, 'shared', sum(len(v) for v in lv.map.shared.values()), 'cpu us/call %.1f' % (tc/200*1e6), 'total us/call %.1f' % (tg/200*1e6), flush=True)
dist.destroy_process_group()
