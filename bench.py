#!/usr/bin/env python
"""bench.py — one JSON line for the OpenCMP hot path on B200.

Workload (BASELINE.json metric "INS s/timestep, assembly Mnnz/s, SpMV GB/s"; configs[2] scaled to SURVEY 8(d)'s
throughput size): one time step of the 2-D incompressible Navier-Stokes Taylor-Green problem on a structured
N x N x 2 triangle mesh of [0,pi]^2, HDiv-DG order 3 / L2 order 2, Oseen linearisation, implicit Euler — Dirichlet
projection, full re-assembly of matrix and right-hand side, additive-Schwarz setup, GMRES solve and the two L2-norm
integrals per Picard iteration, exactly the sequence of opencmp/solvers/base_solver.py:521-587 and
opencmp/models/ins.py:323-355 (see opencmp_b200/workloads.py).

`value` = seconds per time step with all inputs resident in HBM; `e2e` = the same step with the previous solution
and wind uploaded from pinned host memory and the new solution read back inside the timed region.
`--impl reference` times the CPU restatement (oracle/, NumPy/SciPy + sparse LU, the reference's own default
`linear_solver = direct`) on a bounded sample of the same workload.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--workload', default='ins2d', choices=['ins2d', 'ins3d_dim'],
                    help='ins2d: 2-D INS Taylor-Green, HDiv-DG order 3 (BASELINE configs[2] scaled up, the default); '
                         'ins3d_dim: 3-D INS with a diffuse-interface sphere, Taylor-Hood Q2/Q1 hexes (configs[4])')
    ap.add_argument('--N', type=int, default=None, help='cells per direction of the structured mesh '
                                                        '(default 256 for ins2d, 32 for ins3d_dim)')
    ap.add_argument('--order', type=int, default=None)
    ap.add_argument('--cpu-N', type=int, default=None, help='mesh size of the bounded CPU sample')
    ap.add_argument('--no-cpu', action='store_true')
    ap.add_argument('--cpu-procs', type=int, default=0, help='replicas of the CPU sample the reference arm runs '
                                                             'concurrently (default: one per host core)')
    ap.add_argument('--layout', default='bricks', choices=['bricks', 'sphere'],
                    help='ins3d_dim on several GPUs. bricks: one N^3 brick with its own sphere per GPU, lined up along x '
                         '(weak scaling, the default); sphere: ONE sphere in [-1,1]^3 meshed with N^3 hexes in total, '
                         'cells split between the GPUs (BASELINE configs[4] as written; use N = 96 on 8 GPUs for the '
                         'per-GPU size of N = 48 on one)')
    ap.add_argument('--full-mg-setup', action='store_true',
                    help='re-assemble and re-invert the coarse multigrid levels on every Preconditioner.Update() '
                         '(OCMP_MG_REUSE_COARSE=0); default: they are rebuilt only when a Parameter they read (dt, t) '
                         'changes — their operators do not see the Oseen wind — while the finest level (the assembled '
                         'system) is always set up again')
    ap.add_argument('--lag-smoother', action='store_true',
                    help='OCMP_MG_LAG=1: keep the finest level\'s patch inverses until a solve needs > 1.25 x + 2 '
                         'iterations of the first solve after the last fresh set-up (opt-in; default: re-invert on '
                         'every Preconditioner.Update())')
    ap.add_argument('--precond-storage', default='fp64', choices=['fp64', 'fp32', 'bf16'],
                    help='storage of the multigrid data (patch inverses, level matrices inside the cycle); arithmetic '
                         'and the Krylov method stay FP64. fp32 = OCMP_PATCH_STORAGE=fp32 OCMP_SPMV_FP32=1; bf16 = '
                         'bfloat16 patch inverses + FP32 level matrices (opt-in until measured on a B200)')
    ap.add_argument('--dist-poisson', action='store_true', help='also run the distributed Poisson CG leg')
    ap.add_argument('--dist-n', type=int, default=512, help='cells per direction and rank of the distributed leg')
    a = ap.parse_args()
    d = {'ins2d': (256, 3, 28), 'ins3d_dim': (32, 2, 8)}[a.workload]
    a.N = d[0] if a.N is None else a.N
    a.order = d[1] if a.order is None else a.order
    a.cpu_N = d[2] if a.cpu_N is None else a.cpu_N
    return a


class ClockSampler:
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.proc = None
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '200'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ''
        sm, mx, reasons = [], [], set()
        for ln in out.strip().splitlines():
            f = [x.strip() for x in ln.split(',')]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for name, val in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), f[3:7]):
                if val.lower().startswith('active'):
                    reasons.add(name)
        hi = [s for s in sm if s >= 0.5 * max(sm)] if sm else []
        return {'sm_mhz': statistics.median(hi) if hi else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(sm)}


def _cpu_workload(N, order, workload):
    """The bench workload on the CPU restatement with the reference's default linear solver (direct,
    base_model.py:918-922); the oracle backend must be the active one."""
    from opencmp_b200.workloads import INSTaylorGreen, INSSphereDIM3D
    if workload != 'ins3d_dim':
        return INSTaylorGreen(N, order=order, linear_solver='direct', preconditioner=None)
    w = INSSphereDIM3D(N, order=order, preconditioner=None, nu=1.0)

    def direct():
        inv = w.a.mat.Inverse(w.fes.FreeDofs())
        r = w.L.vec.CreateVector()
        r.data = w.L.vec - w.a.mat * w.gfu.vec
        w.gfu.vec.data += inv * r
        w.linear_iterations.append(0)
    w.linear_solve = direct
    return w


def cpu_step_seconds(N, order, steps=1, workload='ins2d'):
    """Oracle (CPU restatement, NumPy/SciPy; NOT NGSolve) timed on one Picard-iterated time step."""
    import opencmp_b200.ngs as ngs
    from oracle.backend import OracleBackend
    old = ngs._backend
    ngs.set_backend(OracleBackend())
    try:
        w = _cpu_workload(N, order, workload)
        w.step()                          # warm-up (lowering, tabulation caches)
        t0 = time.perf_counter()
        for _ in range(steps):
            w.step()
        dt = (time.perf_counter() - t0) / steps
        return dt, w.mesh.ne, w.ndof, w.nnz
    finally:
        ngs.set_backend(old)


def _cpu_replica(job):
    """One worker of the reference arm: its own copy of the CPU sample, warm-up step, then ``steps`` timed steps."""
    N, order, workload, steps = job
    for k in ('OMP_NUM_THREADS', 'OPENBLAS_NUM_THREADS', 'MKL_NUM_THREADS'):
        os.environ[k] = '1'                      # one core per replica; set before NumPy / SciPy load their BLAS
    import opencmp_b200.ngs as ngs
    from oracle.backend import OracleBackend
    ngs.set_backend(OracleBackend())
    w = _cpu_workload(N, order, workload)
    w.step()
    out = []
    for _ in range(steps):
        t0 = time.perf_counter()
        w.step()
        out.append(time.perf_counter() - t0)
    return out, w.mesh.ne, w.ndof, w.nnz


def cpu_replicas(N, order, workload, steps, procs):
    """The CPU restatement is a single-threaded NumPy / SciPy program (SuperLU does not thread). To put every host core
    to work, like a TaskManager-threaded NGSolve run would, ``procs`` independent replicas of the sample run at the same
    time, one per core, and share the memory system: aggregate throughput = procs steps per slowest replica's step —
    the figure a perfectly scaling threaded assembly + solve would reach, i.e. an upper bound in the reference's favour.
    Returns (effective seconds per step of the sample, slowest replica's seconds per step, cells, DOFs, nnz)."""
    import multiprocessing as mp
    with mp.get_context('spawn').Pool(procs) as pool:
        res = pool.map(_cpu_replica, [(N, order, workload, steps)] * procs)
    per_step = [max(r[0][i] for r in res) for i in range(steps)]      # all replicas run step i concurrently
    slow = sum(per_step) / len(per_step)
    return slow / procs, slow, res[0][1], res[0][2], res[0][3]


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    # weak scaling: one N x N x 2 strip (2-D) / one N^3 brick (3-D) per GPU
    cells_full = (args.N ** 3 if args.workload == 'ins3d_dim' else 2 * args.N * args.N) * \
        (1 if args.workload == 'ins3d_dim' and args.layout == 'sphere' else max(1, args.gpus))
    procs = args.cpu_procs
    if not procs:
        # one replica per core this process may run on, at most 32, and no more than a quarter of the free memory holds
        # (a replica of the N = 28 sample peaks at ~0.9 GB)
        try:
            procs = len(os.sched_getaffinity(0))
        except AttributeError:
            procs = os.cpu_count() or 1
        try:
            import psutil
            procs = min(procs, int(psutil.virtual_memory().available // (4 << 30)))
        except Exception:
            pass
    procs = max(1, min(procs, 32))
    cores, slow = 1, None
    try:
        if procs == 1:
            raise RuntimeError('one core')
        per, slow, ne, ndof, nnz = cpu_replicas(args.cpu_N, args.order, args.workload, args.steps, procs)
        cores = procs
    except Exception:                            # e.g. no second core / process pool unavailable: the scalar run
        times = []
        ne = ndof = nnz = 0
        for _ in range(max(1, args.warmup // 3)):
            cpu_step_seconds(args.cpu_N, args.order, workload=args.workload)
        for _ in range(args.steps):
            dt, ne, ndof, nnz = cpu_step_seconds(args.cpu_N, args.order, workload=args.workload)
            times.append(dt)
        per = sum(times) / len(times)
    scaled = per * cells_full / ne
    sample = ('one time step (2 Picard iterations: assemble + SciPy SuperLU) at N={} ({} cells, {} DOFs), '
              'scaled linearly by cell count x{:.1f} to {} strip(s) of N={}'.format(args.cpu_N, ne, ndof, cells_full / ne,
                                                                              max(1, args.gpus), args.N))
    if cores > 1:
        sample += ('; {0} single-threaded replicas of the sample ran concurrently, one per host core (slowest replica '
                   '{1:.2f} s per step), value = that / {0}: the throughput of a perfectly scaling threaded run, an '
                   'upper bound in the CPU path\'s favour'.format(cores, slow))
    line = {'impl': 'reference', 'metric': 'INS s/timestep', 'value': scaled, 'unit': 's', 'n_gpus': args.gpus,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': scaled * 1e3, 'higher_is_better': False,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': workload_config(args, 'cpu'),
            'cpu_baseline': {'value': scaled, 'unit': 's', 'cores': cores, 'kind': 'port', 'sample': sample,
                             'sample_value_s': per},
            'e2e': {'value': scaled, 'unit': 's', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(line))


def workload_config(args, where):
    if args.workload == 'ins3d_dim':
        return {'workload': 'INS-DIM 3D (BASELINE configs[4]): structured {0}^3 hexes on [-1,1]^3, Taylor-Hood Q{1}/Q{2}, '
                            'diffuse-interface sphere R=0.5 (erf profile, lambda = 0.25, phi clamped to [1e-10,1]), rotating '
                            'wall as DIM Dirichlet data, Oseen + implicit Euler, dt=1e-2, nu=1 '
                            '(reference models/ins_dim.py forms)'.format(args.N, args.order, args.order - 1),
                'N': args.N, 'order': args.order,
                'linear_solver': 'GMRES(200) + geometric multigrid V(1,1) on the hex hierarchy, open-star vertex-patch '
                                 'additive Schwarz smoother (damping 0.7), coarse-level phase field, tol 1e-12'
                if where == 'gpu' else 'direct (SuperLU)', 'nonlinear_max_iterations': 3,
                'l2': 'inputs larger than L2 (CSR matrix and patch inverses are GBs); no explicit flush',
                'parallelism': ('single' if args.gpus <= 1 else
                                'element-partitioned: ONE sphere in [-1,1]^3, the {0}^3 hexes split into {1} contiguous '
                                'blocks of the refinement-tree cell order, two ghost layers, halo exchange + all-reduce '
                                'over NCCL issued by the C ABI Krylov driver, distributed multigrid-GMRES'
                                .format(args.N, args.gpus) if args.layout == 'sphere' else
                                'element-partitioned: one {0}^3 brick with its own sphere per GPU (domain [-1,{1}] x '
                                '[-1,1]^2), two ghost layers, halo exchange + all-reduce over NCCL issued by the C ABI '
                                'Krylov driver, distributed multigrid-GMRES'.format(args.N, 2 * args.gpus - 1))}
    return {'workload': 'INS Taylor-Green 2D, structured {0}x{0}x2 triangles on [0,pi]^2, HDiv-DG order {1} / L2 order {2}, '
                        'Oseen + implicit Euler, dt=1e-3, nu=1 (examples/INS scaled up)'.format(args.N, args.order,
                                                                                               args.order - 1),
            'N': args.N, 'order': args.order, 'linear_solver': 'GMRES(100) + geometric multigrid V(1,1), vertex-patch additive Schwarz smoother (damping 0.7), tol 1e-10'
            if where == 'gpu' else 'direct (SuperLU)', 'nonlinear_max_iterations': 3,
            'l2': 'inputs larger than L2 (CSR matrix 3.3 GB, patch inverses 9 GB at N=256); no explicit flush',
            'parallelism': ('element-partitioned: one {0}x{0}x2 strip per GPU (domain [0,pi] x [0,{1} pi]), two ghost '
                            'layers, halo exchange + all-reduce over NCCL issued by the C ABI Krylov driver, distributed '
                            'multigrid-GMRES'
                            .format(args.N, args.gpus)) if args.gpus > 1 else 'single'}


def main():
    args = parse()
    if args.impl == 'reference':
        return run_reference(args)
    if args.precond_storage != 'fp64':
        os.environ['OCMP_PATCH_STORAGE'] = args.precond_storage
        os.environ['OCMP_SPMV_FP32'] = '1'
    if args.full_mg_setup:
        os.environ['OCMP_MG_REUSE_COARSE'] = '0'
    if args.lag_smoother:
        os.environ['OCMP_MG_LAG'] = '1'
    import numpy as np
    import torch
    import torch.distributed as dist
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    import opencmp_b200.ngs as ngs
    from opencmp_b200.backend import CudaBackend, patch_storage, read_profile
    from opencmp_b200.workloads import INSTaylorGreen
    be = CudaBackend(local)
    ngs.set_backend(be)
    lib = be.lib
    t_setup = time.perf_counter()
    if world > 1 and args.workload == 'ins3d_dim':
        # element-partitioned 3-D INS-DIM step: rank r owns the brick [-1 + 2r, 1 + 2r] x [-1,1]^2 at N^3 hexes
        from opencmp_b200.dist_workload import DistributedINSDIM3D
        dins = DistributedINSDIM3D(args.N, world, rank, order=args.order,
                                   bricks=1 if args.layout == 'sphere' else None)
        w = dins.w
    elif world > 1:
        # element-partitioned INS step: rank r owns the strip [0,pi] x [r pi, (r+1) pi] at N x N x 2 triangles
        from opencmp_b200.dist_workload import DistributedINS
        dins = DistributedINS(args.N, world, rank, order=args.order)
        w = dins.w
    elif args.workload == 'ins3d_dim':
        from opencmp_b200.workloads import INSSphereDIM3D
        dins = None
        w = INSSphereDIM3D(args.N, order=args.order, nu=1.0, linear_tolerance=1e-12)
    else:
        dins = None
        w = INSTaylorGreen(args.N, order=args.order)
    torch.cuda.synchronize()
    t_setup = time.perf_counter() - t_setup
    ndof, nnz, ne = w.ndof, w.nnz, w.mesh.ne

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        w.step()
    # ---- timed region: K steps, device-resident ------------------------------------------------------------
    lib.ocmp_profile_reset()
    lib.ocmp_profile_enable(1)
    sampler = ClockSampler(local) if rank == 0 else None
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = lib.ocmp_launch_count()
    e0.record()
    picard = lin_its = 0
    for _ in range(args.steps):
        w.linear_iterations = []
        w.step()
        picard += w.picard_iterations
        lin_its += sum(w.linear_iterations)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if sampler is not None else None
    launches = int(lib.ocmp_launch_count() - l0)
    prof = read_profile(lib)
    lib.ocmp_profile_enable(0)
    # ---- e2e: host buffers in, solution out, copies inside the timed region -------------------------------------
    h_prev = torch.empty(ndof, dtype=torch.float64).pin_memory()
    h_wind = torch.empty(w.V.ndof, dtype=torch.float64).pin_memory()
    h_out = torch.empty(ndof, dtype=torch.float64).pin_memory()
    h_prev.copy_(w.gfu.vec.a)
    h_wind.copy_(w.W.vec.a)
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for _ in range(args.steps):
        w.gfu_0.vec.a.copy_(h_prev, non_blocking=True)
        w.gfu.vec.a.copy_(h_prev, non_blocking=True)
        w.W.vec.a.copy_(h_wind, non_blocking=True)
        w.step()
        h_out.copy_(w.gfu.vec.a, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        h_prev.copy_(h_out)
        h_wind.copy_(h_out[:w.V.ndof])
    f1.record()
    barrier()
    ms_e2e = f0.elapsed_time(f1)
    eu, ep = w.errors()
    # ---- element-partitioned leg (N > 1): halo-exchange SpMV + all-reduced Jacobi-CG on a distributed Poisson problem
    multi = None
    if world > 1 and args.dist_poisson:
        from opencmp_b200.dist_workload import DistributedPoisson
        dp = DistributedPoisson(args.dist_n, 2, world, rank)
        sp_ms = dp.time_spmv()
        its, res, sec_cg, _ = dp.solve(maxit=100)
        tt = torch.tensor([sp_ms, sec_cg / max(1, its) * 1e3], dtype=torch.float64, device='cuda')
        bb = torch.tensor([float(dp.spmv_bytes_owned())], dtype=torch.float64, device='cuda')
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dist.all_reduce(bb)
        multi = {'workload': 'Poisson H1 order 2, structured {0}x{1}x2 triangles, one cell block per rank, one ghost '
                             'layer'.format(args.dist_n, args.dist_n * world),
                 'global_dofs': dp.nglobal, 'spmv_ms_max_over_ranks': float(tt[0]),
                 'spmv_gbs_aggregate': float(bb[0]) / float(tt[0]) / 1e6, 'cg_ms_per_iteration': float(tt[1]),
                 'collectives': 'halo exchange: batched isend/irecv (NCCL); dot products: all_reduce'}
    t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = float(t[0]), float(t[1])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    sec = ms / 1e3 / args.steps
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    peak = float(peaks.get('hbm_gbs', 6650.0))
    peak_src = 'measured (MEASURED_PEAKS.json hbm_gbs)' if 'hbm_gbs' in peaks else 'fallback 6650 GB/s'
    sp = prof['spmv']
    spmv_bytes = nnz * 12 + ndof * 12 + ndof * 8
    spmv_ms = sp['ms'] / max(1, sp['count'])
    spmv_gbs = spmv_bytes / (spmv_ms * 1e-3) / 1e9 if sp['count'] else 0.0
    asm_ms = prof['coef']['ms'] + prof['contract_matrix']['ms'] + prof['contract_vector']['ms']
    n_asm = max(1, picard)
    asm_mnnz = nnz / (asm_ms / n_asm * 1e-3) / 1e6 if asm_ms > 0 else 0.0
    share = {k: round(v['ms'] / ms, 4) for k, v in prof.items()}
    # dominant kernel of the step: the smoother application k_patch_apply (HBM bound: streams the patch inverses)
    ap = prof['asm_apply']
    ap_ms = ap['ms'] / max(1, ap['count'])
    ap_bytes = ap['bytes'] / max(1, ap['count'])
    ap_gbs = ap_bytes / (ap_ms * 1e-3) / 1e9 if ap['count'] else 0.0
    roofline = {'kernel': {'fp64': 'k_patch_apply (8 bs^2 bytes per patch',
                           'fp32': 'k_patch_apply_f32 (FP32-stored patch inverses, 4 bs^2 bytes per patch',
                           'bf16': 'k_patch_apply_bf16 (bfloat16-stored patch inverses, 2 bs^2 bytes per patch'
                           }[patch_storage()] +
                          '; additive-Schwarz smoother, all multigrid levels; launch-weighted mean)',
                'bound': 'hbm', 'achieved': ap_gbs, 'peak': peak, 'unit': 'GB/s', 'frac': ap_gbs / peak,
                'peak_source': peak_src, 'bytes_per_launch': ap_bytes, 'launches': ap['count'],
                'avg_launch_ms': ap_ms, 'share_of_step': share['asm_apply'],
                # ncu --set full of the fine-level launch at N=128 (16 641 patches, 2.330 GB algorithmic): dram read
                # 2.368 GB + write 0.018 GB (profiles/r1_ncu_kernels.md) — 1.024 x the algorithmic bytes, scaled here
                'traffic': 1.024 * ap_bytes if args.workload == 'ins2d' and args.precond_storage == 'fp64' else None,
                'traffic_note': 'per launch, from the measured dram/algorithmic ratio 1.024 of the ncu --set full capture '
                                'at N=128 (dram read 2.368 GB + write 0.018 GB vs 2.330 GB algorithmic, '
                                'profiles/r1_ncu_kernels.md)'}
    line = {
        'metric': 'INS s/timestep', 'value': sec, 'unit': 's', 'n_gpus': world, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': ms / args.steps, 'higher_is_better': False, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': dict(workload_config(args, 'gpu'), coarse_levels=(
            'rebuilt on every update' if os.environ.get('OCMP_MG_REUSE_COARSE', '1') == '0' else
            'rebuilt when a Parameter they read changes (constant dt: once); finest level on every update'),
            finest_level_smoother=('lagged: re-inverted when GMRES iterations grow (OCMP_MG_LAG=1)'
                                   if os.environ.get('OCMP_MG_LAG', '0') == '1' else 're-inverted on every update'),
            precond_storage={
            'patch_inverses': patch_storage(),
            'level_matrices_in_cycle': 'fp32' if os.environ.get('OCMP_SPMV_FP32', '0') == '1' else 'fp64'}),
        'problem': {'cells': ne, 'dofs': ndof, 'nnz': nnz, 'global_dofs': dins.ndof_global if dins else ndof,
                    'ranks': world, 'picard_per_step': picard / args.steps,
                    'gmres_its_per_step': lin_its / args.steps, 'l2_err_u': eu, 'l2_err_p': ep,
                    'setup_s': t_setup},
        'assembly_mnnz_per_s': asm_mnnz, 'spmv_gbs': spmv_gbs,
        'roofline': roofline,
        'roofline_spmv': {'kernel': 'k_spmv<16> fine level (CSR FP64 values + int32 columns)', 'bound': 'hbm',
                          'achieved': spmv_gbs, 'peak': peak, 'unit': 'GB/s', 'frac': spmv_gbs / peak,
                          'frac_of_8000_nominal': spmv_gbs / 8000.0, 'bytes_per_launch': spmv_bytes,
                          'launches': sp['count'], 'avg_launch_ms': spmv_ms},
        'kernel_time_share': share,
        'e2e': {'value': ms_e2e / 1e3 / args.steps, 'unit': 's', 'h2d_bytes_per_step': int(8 * (2 * ndof + w.V.ndof)),
                'd2h_bytes_per_step': int(8 * ndof)},
        'gpu_launches': launches, 'clocks': clocks,
    }
    if multi is not None:
        line['multi_gpu'] = multi
    if not args.no_cpu and world == 1:
        try:
            per, cne, cnd, _ = cpu_step_seconds(args.cpu_N, args.order, workload=args.workload)
            line['cpu_baseline'] = {
                'value': per * ne / cne, 'unit': 's', 'cores': 1, 'kind': 'port',
                'sample': 'CPU restatement (NumPy/SciPy, not NGSolve): one time step at N={} ({} cells, {} DOFs) took '
                          '{:.2f} s; scaled linearly by cell count x{:.1f}'.format(args.cpu_N, cne, cnd, per, ne / cne)}
        except Exception as exc:                                    # pragma: no cover
            line['cpu_baseline'] = {'value': None, 'unit': 's', 'cores': 1, 'kind': 'port', 'sample': repr(exc)}
    # ---- matrix-free application of the same operator (BilinearForm.Apply: k_coef + k_lin, no CSR values read) next
    # to the CSR SpMV — outside the step's timed region, its own CUDA events. Runs LAST, when every other number is
    # already on the host: its GPU cases have not run on a B200 yet, and it must never cost the bench line
    matfree = None
    try:
        if world == 1:
            xv = w.gfu.vec.CreateVector()
            xv.data = w.gfu.vec
            yv = w.gfu.vec.CreateVector()
            for _ in range(2):
                w.a.Apply(xv, yv)
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            g0.record()
            for _ in range(5):
                w.a.Apply(xv, yv)
            g1.record()
            torch.cuda.synchronize()
            ycsr = w.a.mat * xv
            den = float(ycsr.Norm())
            matfree = {'ms_per_apply': g0.elapsed_time(g1) / 5,
                       'rel_diff_vs_csr_spmv': float((yv - ycsr).Norm()) / den if den > 0 else None,
                       'note': 'BilinearForm.Apply(x, y): trial rows become field slots of k_coef, k_lin contracts with '
                               'the test rows; FP64-compute bound (quadrature), reads no matrix'}
    except Exception as exc:                                    # pragma: no cover
        matfree = {'error': repr(exc)}
    if matfree is not None:
        line['matrix_free_apply'] = matfree
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
