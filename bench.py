#!/usr/bin/env python
"""bench.py — one JSON line for the OpenCMP hot path (assemble + solve per time step) on B200.

Headline workload (BASELINE.json metric "INS s/timestep ... at 1/2/4/8 B200", configs[4]): one time step of the 3-D
incompressible Navier-Stokes problem with a diffuse-interface sphere (reference opencmp/models/ins_dim.py), Taylor-Hood
Q2/Q1 on structured hexes of [-1,1]^3, Oseen linearisation + implicit Euler — per Picard iteration: Dirichlet
projection, full re-assembly of matrix and right-hand side, multigrid set-up, GMRES solve and the two L2-norm integrals,
the sequence of opencmp/solvers/base_solver.py:521-587 and opencmp/models/ins.py:323-355 (opencmp_b200/workloads.py).
N GPUs: ONE sphere, the mesh refined with the rank count, every rank
owning one compact brick of cells: 48^3, 64^3, 80^3, 96^3 isotropic hexes on 1, 2, 4, 8 GPUs (110.6 k, 131 k, 128 k,
110.6 k cells per GPU; 2.86 M DOFs on one GPU, 22.6 M on 8 — weak scaling, per-GPU load within +19 % of N = 1).

At N = 1 the line also carries `ins2d`: the 2-D INS Taylor-Green step (configs[2] scaled to SURVEY 8(d)'s throughput
size, HDiv-DG order 3 on 256 x 256 x 2 triangles, 2.6 M DOFs) — the workload the SpMV / smoother / assembly rooflines of
round 1 were quoted on.

`value`  = seconds per time step, inputs resident in HBM, timed with CUDA events around a loop WITHOUT per-launch
           profiling; the per-kernel shares and rooflines come from a second, profiled loop of the same steps.
`e2e`    = the same step with the previous solution and wind uploaded from pinned host memory and the new solution
           read back inside the timed region.
`--impl reference` times the CPU restatement (oracle/: NumPy assembly + SciPy SuperLU, the reference's default
`linear_solver = direct`) on a sample of the same workload small enough to finish, reports that MEASURED time as
`value` with the sample's own size in `config`, runs the GPU path on the same sample in the same invocation
(`gpu_same_config`, `measured_ratio`) and keeps the linear extrapolation to the full size as a labelled extra.
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

DEFAULTS = {'ins3d_dim': dict(N=48, order=2), 'ins2d': dict(N=256, order=3)}
# Every time step runs exactly two Picard (Oseen) iterations — assemble, set up the preconditioner, solve, two norms —
# instead of iterating to a tolerance: the flows relax towards a steady state, so a tolerance-driven loop does 3, 2,
# then 1 iteration per step and "seconds per step" would depend on which steps the timed region happens to hold.
PICARD = dict(nonlinear_max_iterations=2, nonlinear_tolerance=(0.0, 0.0))
# 3-D workload: the diffuse wall's angular velocity oscillates, omega(t) = 1 + 0.5 sin(2 pi t / (10 dt)) — with a steady
# wall the flow relaxes to rigid rotation within a few steps, the previous solution becomes an almost exact initial
# guess and GMRES can no longer reduce the residual by 1e-12 relative to the initial one (workloads.INSSphereDIM3D)
WALL = dict(wall_period=0.1, wall_amp=0.5)
# dram (read + write) bytes / algorithmic bytes of k_patch_apply_stream, fine-level launch, ncu --set full on a B200
# (profiles/r2_ncu_kernels.md): FP64 2.368 / 2.330 GB, FP32 1.207 / 1.162 GB; bfloat16 was not captured (estimate)
PATCH_TRAFFIC_RATIO = {'fp64': 1.016, 'fp32': 1.038, 'bf16': 1.05}


def parse(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--workload', default='ins3d_dim', choices=['ins3d_dim', 'ins2d'],
                    help='ins3d_dim: 3-D INS with a diffuse-interface sphere, Taylor-Hood Q2/Q1 hexes (BASELINE '
                         'configs[4], the default and the scaling workload); ins2d: 2-D INS Taylor-Green, HDiv-DG '
                         'order 3 (configs[2] scaled up; also reported as the `ins2d` object of the default line)')
    ap.add_argument('--N', type=int, default=None, help='cells per direction (and per GPU) of the structured mesh '
                                                        '(default 48 for ins3d_dim, 256 for ins2d)')
    ap.add_argument('--order', type=int, default=None)
    ap.add_argument('--cpu-N', type=int, default=None, help='mesh size of the bounded CPU sample')
    ap.add_argument('--no-cpu', action='store_true', help='skip the in-line cpu_baseline leg')
    ap.add_argument('--no-secondary', action='store_true', help='skip the ins2d object of the default line')
    ap.add_argument('--profile-steps', type=int, default=2, help='steps of the second, per-launch profiled loop')
    ap.add_argument('--layout', default='sphere', choices=['sphere', 'bricks'],
                    help='ins3d_dim on several GPUs. sphere (default): ONE sphere in [-1,1]^3, mesh refined with the '
                         'rank count, one compact N^3 brick of cells per rank; bricks: one [-1,1]^3 brick with its own '
                         'sphere per GPU, lined up along x')
    ap.add_argument('--full-mg-setup', action='store_true',
                    help='re-assemble and re-invert the coarse multigrid levels on every Preconditioner.Update() '
                         '(OCMP_MG_REUSE_COARSE=0); default: only when a Parameter they read changes')
    ap.add_argument('--lag-smoother', action='store_true',
                    help='OCMP_MG_LAG=1: keep the finest level\'s patch inverses while GMRES iteration counts hold')
    ap.add_argument('--precond-storage', default='fp32', choices=['fp64', 'fp32', 'bf16'],
                    help='storage of the multigrid data (patch inverses, level matrices inside the cycle); arithmetic '
                         'and the Krylov method around the cycle stay FP64. Default fp32: same GMRES iteration counts '
                         'as fp64 on both workloads (measured, profiles/r2_bench_results.md)')
    a = ap.parse_args(argv)
    d = DEFAULTS[a.workload]
    a.N = d['N'] if a.N is None else a.N
    a.order = d['order'] if a.order is None else a.order
    return a


def cpu_sample_size(workload, steps, cpu_N=None):
    """Mesh size of the CPU sample: about 10-15 s (3-D N = 8) / 10 s (2-D N = 28) of oracle work per step when few
    steps are asked for, a smaller one when the driver asks for many, so that the whole run ends within minutes."""
    if cpu_N:
        return cpu_N
    if workload == 'ins3d_dim':
        return 8 if steps <= 6 else 6
    return 28 if steps <= 10 else 20


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


# ---- CPU restatement (oracle) ---------------------------------------------------------------------------------------
def _cpu_workload(N, order, workload):
    """The bench workload on the CPU restatement with the reference's default linear solver (direct,
    base_model.py:918-922); the oracle backend must be the active one."""
    from opencmp_b200.workloads import INSTaylorGreen, INSSphereDIM3D
    if workload != 'ins3d_dim':
        return INSTaylorGreen(N, order=order, linear_solver='direct', preconditioner=None, **PICARD)
    w = INSSphereDIM3D(N, order=order, preconditioner=None, nu=1.0, periodic=(False, False, False), **PICARD, **WALL)

    def direct():
        inv = w.a.mat.Inverse(w.fes.FreeDofs())
        r = w.L.vec.CreateVector()
        r.data = w.L.vec - w.a.mat * w.gfu.vec
        w.gfu.vec.data += inv * r
        w.linear_iterations.append(0)
    w.linear_solve = direct
    return w


def cpu_step_seconds(N, order, steps=1, workload='ins2d', budget_s=None):
    """Oracle (CPU restatement, NumPy/SciPy; NOT NGSolve) timed on Picard-iterated time steps: one untimed warm-up
    step (lowering, tabulation caches), then up to ``steps`` timed ones (fewer when ``budget_s`` runs out). Returns
    (mean seconds per step, steps timed, cells, DOFs, nnz)."""
    import opencmp_b200.ngs as ngs
    from oracle.backend import OracleBackend
    old = ngs._backend
    ngs.set_backend(OracleBackend())
    try:
        w = _cpu_workload(N, order, workload)
        w.step()
        times = []
        t_all = time.perf_counter()
        for _ in range(steps):
            t0 = time.perf_counter()
            w.step()
            times.append(time.perf_counter() - t0)
            if budget_s is not None and time.perf_counter() - t_all + times[-1] > budget_s:
                break
        return sum(times) / len(times), len(times), w.mesh.ne, w.ndof, w.nnz
    finally:
        ngs.set_backend(old)


def workload_config(args, where, N=None, gpus=None):
    N = args.N if N is None else N
    gpus = args.gpus if gpus is None else gpus
    if args.workload == 'ins3d_dim':
        if gpus <= 1:
            par = 'single'
        elif args.layout == 'sphere':
            from opencmp_b200.dist import brick_grid
            from opencmp_b200.dist_workload import sphere_total_cells
            g, nt = brick_grid(gpus), sphere_total_cells(N, gpus)
            par = ('element-partitioned: ONE sphere in [-1,1]^3 meshed with {0}^3 isotropic hexes (N, 4N/3, 5N/3, 2N per '
                   'direction on 1, 2, 4, 8 GPUs), split into {1} x {2} x {3} compact bricks of {4} cells, one per GPU '
                   '({5:.2f} x the single-GPU cell count), two ghost layers, halo exchange + all-reduce over NCCL issued '
                   'by the C ABI Krylov driver, distributed multigrid-GMRES'
                   .format(nt, g[0], g[1], g[2], nt ** 3 // gpus, nt ** 3 / gpus / N ** 3))
        else:
            par = ('element-partitioned: one {0}^3 brick with its own sphere per GPU (domain [-1,{1}] x [-1,1]^2), two '
                   'ghost layers, halo exchange + all-reduce over NCCL issued by the C ABI Krylov driver, distributed '
                   'multigrid-GMRES'.format(N, 2 * gpus - 1))
        return {'workload': 'INS-DIM 3D (BASELINE configs[4]): structured hexes on [-1,1]^3 ({0}^3 on one GPU), Taylor-Hood '
                            'Q{1}/Q{2}, diffuse-interface sphere R=0.5 (erf profile, lambda = 0.25, phi clamped to '
                            '[1e-10,1]), rotating wall as DIM Dirichlet data with omega(t) = 1 + 0.5 sin(2 pi t / 0.1), Oseen + implicit '
                            'Euler, dt=1e-2, nu=1 '
                            '(reference models/ins_dim.py forms)'.format(N, args.order, args.order - 1),
                'N': N, 'order': args.order,
                'linear_solver': 'GMRES(200) + geometric multigrid V(1,1) on the hex hierarchy, open-star vertex-patch '
                                 'additive Schwarz smoother (damping 0.7), coarse-level phase field, coarse operators with the fine '
                                 'level\'s volume penalty (OCMP_MG_COARSE_H=fine), tol 1e-12'
                if where == 'gpu' else 'direct (SciPy SuperLU), the reference\'s default linear_solver',
                'picard_iterations_per_step': 2,
                'l2': 'inputs larger than L2 (CSR matrix and patch inverses are GBs); no explicit flush',
                'parallelism': par}
    return {'workload': 'INS Taylor-Green 2D, structured {0}x{0}x2 triangles on [0,pi]^2, HDiv-DG order {1} / L2 order '
                        '{2}, Oseen + implicit Euler, dt=1e-3, nu=1 (examples/INS scaled up)'
                        .format(N, args.order, args.order - 1),
            'N': N, 'order': args.order,
            'linear_solver': 'GMRES(100) + geometric multigrid V(1,1), vertex-patch additive Schwarz smoother (damping '
                             '0.7), tol 1e-10' if where == 'gpu' else
                             'direct (SciPy SuperLU), the reference\'s default linear_solver',
            'picard_iterations_per_step': 2,
            'l2': 'inputs larger than L2 (CSR matrix and patch inverses are GBs at N=256); no explicit flush',
            'parallelism': ('element-partitioned: one {0}x{0}x2 strip per GPU (domain [0,pi] x [0,{1} pi]), two ghost '
                            'layers, halo exchange + all-reduce over NCCL issued by the C ABI Krylov driver, distributed '
                            'multigrid-GMRES'.format(N, gpus)) if gpus > 1 else 'single'}


# ---- GPU side ---------------------------------------------------------------------------------------------------------
def set_storage_env(args):
    os.environ['OCMP_PATCH_STORAGE'] = args.precond_storage
    os.environ['OCMP_SPMV_FP32'] = '0' if args.precond_storage == 'fp64' else '1'
    if args.full_mg_setup:
        os.environ['OCMP_MG_REUSE_COARSE'] = '0'
    if args.lag_smoother:
        os.environ['OCMP_MG_LAG'] = '1'


def make_gpu_workload(workload, N, order, world=1, rank=0, layout='sphere'):
    """(workload object, distributed wrapper or None)"""
    if world > 1 and workload == 'ins3d_dim':
        from opencmp_b200.dist_workload import DistributedINSDIM3D
        d = DistributedINSDIM3D(N, world, rank, order=order, layout='sphere' if layout == 'sphere' else None, **PICARD,
                                **WALL)
        return d.w, d
    if world > 1:
        from opencmp_b200.dist_workload import DistributedINS
        d = DistributedINS(N, world, rank, order=order, **PICARD)
        return d.w, d
    if workload == 'ins3d_dim':
        from opencmp_b200.workloads import INSSphereDIM3D
        return INSSphereDIM3D(N, order=order, nu=1.0, linear_tolerance=1e-12, periodic=(False, False, False),
                              **PICARD, **WALL), None
    from opencmp_b200.workloads import INSTaylorGreen
    return INSTaylorGreen(N, order=order, **PICARD), None


class GpuTimer:
    """The three timed loops of one workload: clean (value), e2e (host buffers), profiled (shares / rooflines)."""

    def __init__(self, torch, lib, w, barrier):
        self.torch, self.lib, self.w, self.barrier = torch, lib, w, barrier

    def clean(self, steps):
        torch, w = self.torch, self.w
        self.lib.ocmp_profile_enable(0)
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = self.lib.ocmp_launch_count()
        e0.record()
        picard = its = 0
        per_step = []
        for _ in range(steps):
            w.linear_iterations = []
            w.step()
            picard += w.picard_iterations
            its += sum(w.linear_iterations)
            per_step.append(list(w.linear_iterations))
        e1.record()
        self.barrier()
        return dict(ms=e0.elapsed_time(e1), picard=picard, its=its, launches=int(self.lib.ocmp_launch_count() - l0),
                    its_per_solve=per_step)

    def e2e(self, steps):
        torch, w = self.torch, self.w
        ndof = w.ndof
        h_prev = torch.empty(ndof, dtype=torch.float64).pin_memory()
        h_wind = torch.empty(w.V.ndof, dtype=torch.float64).pin_memory()
        h_out = torch.empty(ndof, dtype=torch.float64).pin_memory()
        h_prev.copy_(w.gfu.vec.a)
        h_wind.copy_(w.W.vec.a)
        self.barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        per_step = []
        for _ in range(steps):
            w.linear_iterations = []
            w.gfu_0.vec.a.copy_(h_prev, non_blocking=True)
            w.gfu.vec.a.copy_(h_prev, non_blocking=True)
            w.W.vec.a.copy_(h_wind, non_blocking=True)
            w.step()
            h_out.copy_(w.gfu.vec.a, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            h_prev.copy_(h_out)
            h_wind.copy_(h_out[:w.V.ndof])
            per_step.append(list(w.linear_iterations))
        f1.record()
        self.barrier()
        return dict(ms=f0.elapsed_time(f1), h2d=int(8 * (2 * ndof + w.V.ndof)), d2h=int(8 * ndof),
                    its_per_solve=per_step)

    def profiled(self, steps):
        from opencmp_b200.backend import read_profile
        torch, w, lib = self.torch, self.w, self.lib
        lib.ocmp_profile_reset()
        lib.ocmp_profile_enable(1)
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        picard = 0
        for _ in range(steps):
            w.step()
            picard += w.picard_iterations
        e1.record()
        self.barrier()
        prof = read_profile(lib)
        lib.ocmp_profile_enable(0)
        return dict(ms=e0.elapsed_time(e1), picard=picard, prof=prof)


def fp64_peak_tflops(torch, n=8192, reps=6):
    """cuBLAS DGEMM n^3 through torch.matmul (a library GEMM used as the FP64 yardstick only — BASELINE.md section 2
    asks for this calibration before an assembly roofline is quoted): best of ``reps`` and the mean of ~1 s back to back."""
    a = torch.randn(n, n, dtype=torch.float64, device='cuda')
    b = torch.randn(n, n, dtype=torch.float64, device='cuda')
    c = torch.empty_like(a)
    torch.matmul(a, b, out=c)
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b, out=c)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k = max(3, int(1000.0 / best))
    e0.record()
    for _ in range(k):
        torch.matmul(a, b, out=c)
    e1.record()
    torch.cuda.synchronize()
    fl = 2.0 * n ** 3
    return {'burst': fl / best / 1e9, 'sustained': fl * k / e0.elapsed_time(e1) / 1e9,
            'how': 'torch.matmul float64 {0}^3 (cuBLAS DGEMM), best of {1} / mean of {2} back to back'.format(n, reps, k)}


def rooflines(be, w, prof, ms_prof, picard, peak, peak_src, storage, fp64_peak):
    """Roofline objects from one profiled loop. Algorithmic bytes: the fine-level CSR SpMV moves nnz (8 + 4) + nrows
    (4 + 8) + ncols 8 bytes; the smoother application bs^2 x (8 | 4 | 2) bytes per patch + 16 per dof
    (counted per launch inside the library); assembly: flops of the B^T D B contraction from the launch plans."""
    nnz, ndof = w.nnz, w.ndof
    share = {k: round(v['ms'] / ms_prof, 4) for k, v in prof.items()}
    sp = prof['spmv']
    spmv_bytes = nnz * 12 + ndof * 12 + ndof * 8
    spmv_ms = sp['ms'] / max(1, sp['count'])
    spmv_gbs = spmv_bytes / (spmv_ms * 1e-3) / 1e9 if sp['count'] else 0.0
    ap = prof['asm_apply']
    ap_ms = ap['ms'] / max(1, ap['count'])
    ap_bytes = ap['bytes'] / max(1, ap['count'])
    ap_gbs = ap_bytes / (ap_ms * 1e-3) / 1e9 if ap['count'] else 0.0
    kname = {'fp64': 'k_patch_apply_stream<double> (8 bs^2 bytes per patch',
             'fp32': 'k_patch_apply_stream<float> (FP32-stored patch inverses, 4 bs^2 bytes per patch',
             'bf16': 'k_patch_apply_stream<bf16> (bfloat16-stored patch inverses, 2 bs^2 bytes per patch'}[storage]
    r_patch = {'kernel': kname + '; additive-Schwarz smoother, all multigrid levels: total bytes / total time)',
               'bound': 'hbm', 'achieved': ap_gbs, 'peak': peak, 'unit': 'GB/s', 'frac': ap_gbs / peak,
               'peak_source': peak_src, 'bytes_per_launch': ap_bytes, 'launches': ap['count'], 'avg_launch_ms': ap_ms,
               'share_of_step': share['asm_apply'],
               # dram__bytes_read + write of the fine-level launch in the ncu --set full captures, relative to the
               # algorithmic bytes (profiles/r2_ncu_kernels.md): FP64 2.368 / 2.330 GB, FP32 see the same file
               'traffic': PATCH_TRAFFIC_RATIO.get(storage, 1.0) * ap_bytes if ap['count'] else None,
               'traffic_note': 'per launch: algorithmic bytes x the dram / algorithmic ratio of the ncu --set full capture '
                               'of the fine-level launch ({:.3f}; profiles/r2_ncu_kernels.md)'
                               .format(PATCH_TRAFFIC_RATIO.get(storage, 1.0))}
    r_spmv = {'kernel': 'k_spmv<16, double> fine level (CSR FP64 values + int32 columns; Krylov operator)',
              'bound': 'hbm', 'achieved': spmv_gbs, 'peak': peak, 'unit': 'GB/s', 'frac': spmv_gbs / peak,
              'frac_of_8000_nominal': spmv_gbs / 8000.0, 'peak_source': peak_src, 'bytes_per_launch': spmv_bytes,
              'launches': sp['count'], 'avg_launch_ms': spmv_ms, 'share_of_step': share['spmv'], 'traffic': None}
    mg = prof['spmv_multigrid']
    mg_gbs = mg['bytes'] / (mg['ms'] * 1e-3) / 1e9 if mg['ms'] > 0 else 0.0
    r_mg = {'kernel': 'k_spmv / k_spmv_vec inside the multigrid cycle (level residuals with fused epilogue, restriction, '
                      'prolongation; FP32-stored level matrices when precond_storage != fp64): total bytes / total time',
            'bound': 'hbm', 'achieved': mg_gbs, 'peak': peak, 'unit': 'GB/s', 'frac': mg_gbs / peak,
            'peak_source': peak_src, 'bytes_per_launch': mg['bytes'] / max(1, mg['count']), 'launches': mg['count'],
            'avg_launch_ms': mg['ms'] / max(1, mg['count']), 'share_of_step': share['spmv_multigrid'], 'traffic': None,
            'bytes_note': 'CSR byte count nnz (value + 4) + 20 rows per product; the node-grouped kernel reads the '
                          'column indices once per nc x nc values, so its DRAM traffic is lower than this count'}
    asm_ms = prof['coef']['ms'] + prof['contract_matrix']['ms'] + prof['contract_vector']['ms']
    n_asm = max(1, picard)
    asm_mnnz = nnz / (asm_ms / n_asm * 1e-3) / 1e6 if asm_ms > 0 else 0.0
    flops = be.matrix_flops(w.a.program())
    cm_ms = prof['contract_matrix']['ms'] / n_asm
    tfl = flops / (cm_ms * 1e-3) / 1e12 if cm_ms > 0 else 0.0
    r_asm = {'kernel': 'k_contract (B^T D B local matrices + scatter), fine-level matrix assembly', 'bound': 'fp64',
             'achieved': tfl, 'peak': fp64_peak['burst'] if fp64_peak else None, 'unit': 'TFLOP/s',
             'frac': (tfl / fp64_peak['burst']) if fp64_peak else None,
             'peak_source': fp64_peak['how'] if fp64_peak else 'not calibrated in this run',
             'flops_per_assembly': flops, 'ms_per_assembly': cm_ms, 'assembly_mnnz_per_s': asm_mnnz,
             'share_of_step': round(share['contract_matrix'] + share['coef'] + share['contract_vector'], 4),
             'note': 'ms_per_assembly also holds the re-discretised coarse multigrid levels when they are rebuilt'}
    return share, r_patch, r_spmv, r_asm, asm_mnnz, spmv_gbs, r_mg


def measure_workload(args, torch, be, w, dins, steps, warmup, profile_steps, barrier, peaks, fp64_peak, sampler_rank):
    lib = be.lib
    for _ in range(warmup):
        w.step()
    gt = GpuTimer(torch, lib, w, barrier)
    sampler = ClockSampler(sampler_rank) if sampler_rank is not None else None
    clean = gt.clean(steps)
    clocks = sampler.stop() if sampler is not None else None
    e2e = gt.e2e(steps)
    prof = gt.profiled(profile_steps) if profile_steps > 0 else None
    eu, ep = w.errors()
    return dict(clean=clean, e2e=e2e, prof=prof, clocks=clocks, errs=(eu, ep))


def run_b200(args):
    set_storage_env(args)
    import torch
    import torch.distributed as dist
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    import opencmp_b200.ngs as ngs
    from opencmp_b200.backend import CudaBackend, patch_storage
    be = CudaBackend(local)
    ngs.set_backend(be)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    peak = float(peaks.get('hbm_gbs', 6650.0))
    peak_src = 'measured (MEASURED_PEAKS.json hbm_gbs)' if 'hbm_gbs' in peaks else 'fallback 6650 GB/s'
    fp64_peak = fp64_peak_tflops(torch) if rank == 0 else None

    t_setup = time.perf_counter()
    w, dins = make_gpu_workload(args.workload, args.N, args.order, world, rank, args.layout)
    torch.cuda.synchronize()
    t_setup = time.perf_counter() - t_setup
    ndof, nnz, ne = w.ndof, w.nnz, w.mesh.ne
    m = measure_workload(args, torch, be, w, dins, args.steps, args.warmup, args.profile_steps, barrier, peaks,
                         fp64_peak, local if rank == 0 else None)
    t = torch.tensor([m['clean']['ms'], m['e2e']['ms']], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = float(t[0]), float(t[1])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    storage = patch_storage()
    pr = m['prof']
    share, r_patch, r_spmv, r_asm, asm_mnnz, spmv_gbs, r_mg = rooflines(be, w, pr['prof'], pr['ms'], pr['picard'], peak,
                                                                        peak_src, storage, fp64_peak)
    # the dominant kernel of the step leads the line
    by_share = sorted(((share['asm_apply'], r_patch), (share['spmv'], r_spmv), (share['spmv_multigrid'], r_mg)),
                      key=lambda x: -x[0])
    sec = ms / 1e3 / args.steps
    line = {
        'metric': 'INS s/timestep', 'value': sec, 'unit': 's', 'n_gpus': world, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': ms / args.steps, 'higher_is_better': False, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': dict(workload_config(args, 'gpu', gpus=world), coarse_levels=(
            'rebuilt on every update' if os.environ.get('OCMP_MG_REUSE_COARSE', '1') == '0' else
            'rebuilt when a Parameter they read changes (constant dt: once); finest level on every update'),
            finest_level_smoother=('lagged: re-inverted when GMRES iterations grow (OCMP_MG_LAG=1)'
                                   if os.environ.get('OCMP_MG_LAG', '0') == '1' else 're-inverted on every update'),
            precond_storage={'patch_inverses': storage,
                             'level_matrices_in_cycle': 'fp32' if os.environ.get('OCMP_SPMV_FP32', '0') == '1'
                             else 'fp64'}),
        'problem': {'cells': ne, 'dofs': ndof, 'nnz': nnz, 'global_dofs': dins.ndof_global if dins else ndof,
                    'ranks': world, 'owned_cells_per_gpu': (dins.gmesh.ne // world) if dins else ne,
                    'picard_per_step': m['clean']['picard'] / args.steps,
                    'gmres_its_per_step': m['clean']['its'] / args.steps,
                    'ms_per_gmres_iteration': ms / max(1, m['clean']['its']),
                    'gmres_its_per_solve': m['clean']['its_per_solve'],
                    'gmres_its_per_solve_e2e': m['e2e']['its_per_solve'], 'l2_err_u': m['errs'][0],
                    'l2_err_p': m['errs'][1], 'setup_s': t_setup},
        'timing': 'value: CUDA events around {0} steps with per-launch profiling OFF; kernel shares and rooflines: a '
                  'second loop of {1} steps with an event pair around every launch'.format(args.steps,
                                                                                          args.profile_steps),
        'assembly_mnnz_per_s': asm_mnnz, 'spmv_gbs': spmv_gbs,
        'roofline': by_share[0][1], 'roofline_patch_apply': r_patch, 'roofline_spmv': r_spmv,
        'roofline_spmv_multigrid': r_mg,
        'roofline_assembly': r_asm, 'fp64_peak_tflops': fp64_peak,
        'kernel_time_share': share, 'profiled_ms_per_step': pr['ms'] / max(1, args.profile_steps),
        'e2e': {'value': ms_e2e / 1e3 / args.steps, 'unit': 's', 'h2d_bytes_per_step': m['e2e']['h2d'],
                'd2h_bytes_per_step': m['e2e']['d2h']},
        'gpu_launches': m['clean']['launches'], 'clocks': m['clocks'],
    }
    if world == 1:
        # ---- the GPU path on the CPU sample's size (what --impl reference measures on the host) -----------------------
        cN = cpu_sample_size(args.workload, args.steps, args.cpu_N)
        try:
            ws, _ = make_gpu_workload(args.workload, cN, args.order)
            for _ in range(2):
                ws.step()
            small = GpuTimer(torch, be.lib, ws, barrier).clean(3)
            line['same_size_as_cpu_sample'] = {'N': cN, 'dofs': ws.ndof, 'value': small['ms'] / 3e3, 'unit': 's',
                                               'gmres_its_per_step': small['its'] / 3}
            del ws
        except Exception as exc:                                    # pragma: no cover
            line['same_size_as_cpu_sample'] = {'N': cN, 'error': repr(exc)}
        if not args.no_cpu:
            try:
                per, nst, cne, cnd, _ = cpu_step_seconds(cN, args.order, steps=2, workload=args.workload, budget_s=30.0)
                line['cpu_baseline'] = {
                    'value': per * ne / cne, 'unit': 's', 'cores': 1, 'kind': 'port',
                    'sample': 'CPU restatement (NumPy/SciPy, not NGSolve): {} time step(s) at N={} ({} cells, {} DOFs) took '
                              '{:.2f} s each (measured); value = that scaled linearly by cell count x{:.1f} (an '
                              'extrapolation)'.format(nst, cN, cne, cnd, per, ne / cne),
                    'sample_value_s': per, 'sample_N': cN}
                ss = line.get('same_size_as_cpu_sample', {})
                if 'value' in ss:
                    line['cpu_baseline']['measured_ratio_at_sample_size'] = per / ss['value']
            except Exception as exc:                                # pragma: no cover
                line['cpu_baseline'] = {'value': None, 'unit': 's', 'cores': 1, 'kind': 'port', 'sample': repr(exc)}
        # ---- secondary workload of the default line: 2-D INS at the round-1 throughput size -------------------------
        if args.workload == 'ins3d_dim' and not args.no_secondary:
            try:
                del w
                torch.cuda.empty_cache()
                a2 = parse(['--workload', 'ins2d'])
                w2, _ = make_gpu_workload('ins2d', a2.N, a2.order)
                m2 = measure_workload(a2, torch, be, w2, None, 3, 2, 2, barrier, peaks, fp64_peak, None)
                p2 = m2['prof']
                sh2, rp2, rs2, ra2, mn2, sg2, rm2 = rooflines(be, w2, p2['prof'], p2['ms'], p2['picard'], peak, peak_src,
                                                              storage, fp64_peak)
                line['ins2d'] = {'config': workload_config(a2, 'gpu', gpus=1), 'value': m2['clean']['ms'] / 3e3,
                                 'unit': 's', 'steps': 3, 'warmup': 2,
                                 'e2e': {'value': m2['e2e']['ms'] / 3e3, 'unit': 's',
                                         'h2d_bytes_per_step': m2['e2e']['h2d'], 'd2h_bytes_per_step': m2['e2e']['d2h']},
                                 'problem': {'cells': w2.mesh.ne, 'dofs': w2.ndof, 'nnz': w2.nnz,
                                             'gmres_its_per_step': m2['clean']['its'] / 3,
                                             'picard_per_step': m2['clean']['picard'] / 3,
                                             'l2_err_u': m2['errs'][0], 'l2_err_p': m2['errs'][1]},
                                 'gpu_launches': m2['clean']['launches'], 'assembly_mnnz_per_s': mn2, 'spmv_gbs': sg2,
                                 'roofline_patch_apply': rp2, 'roofline_spmv': rs2, 'roofline_spmv_multigrid': rm2,
                                 'roofline_assembly': ra2,
                                 'kernel_time_share': sh2}
            except Exception as exc:                                # pragma: no cover
                line['ins2d'] = {'error': repr(exc)}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# ---- reference arm -----------------------------------------------------------------------------------------------------
def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cN = cpu_sample_size(args.workload, args.steps, args.cpu_N)
    # ---- the CPU path, measured: K steps of the sample (fewer only if ~4 minutes would not hold them) -----------------
    for _ in range(max(0, args.warmup // 3 - 1)):
        cpu_step_seconds(cN, args.order, steps=1, workload=args.workload)
    per, nst, ne, ndof, nnz = cpu_step_seconds(cN, args.order, steps=args.steps, workload=args.workload, budget_s=240.0)
    full_cells = (args.N ** 3 if args.workload == 'ins3d_dim' else 2 * args.N * args.N) * max(1, args.gpus)
    sample = ('{} time step(s) (Picard-iterated: assemble with NumPy, solve with SciPy SuperLU) of the same workload at '
              'N={} ({} cells, {} DOFs, {} non-zeros), measured on this host, one process'.format(nst, cN, ne, ndof, nnz))
    ref_args = argparse.Namespace(**vars(args))
    line = {'impl': 'reference', 'metric': 'INS s/timestep', 'value': per, 'unit': 's', 'n_gpus': args.gpus,
            'steps': nst, 'warmup': 1, 'ms_per_step': per * 1e3, 'higher_is_better': False, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': dict(workload_config(ref_args, 'cpu', N=cN, gpus=1),
                           sample_of='the b200 arm\'s N={} per GPU on {} GPU(s); the CPU path cannot hold that size '
                                     '(sparse LU of {} cells)'.format(args.N, max(1, args.gpus), full_cells)),
            'cpu_baseline': {'value': per, 'unit': 's', 'cores': 1, 'kind': 'port', 'sample': sample},
            'e2e': {'value': per, 'unit': 's', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'scaled_to_full_config': {'value': per * full_cells / ne, 'unit': 's',
                                      'note': 'EXTRAPOLATION, not a measurement: the sample time scaled linearly by '
                                              'cell count x{:.1f} (sparse LU grows faster than linearly, so this '
                                              'flatters the CPU path)'.format(full_cells / ne)}}
    # ---- the GPU path on the SAME sample, same invocation: a measured same-config ratio -------------------------------
    try:
        import torch
        if torch.cuda.is_available():
            set_storage_env(args)
            import opencmp_b200.ngs as ngs
            from opencmp_b200.backend import CudaBackend
            torch.cuda.set_device(0)
            be = CudaBackend(0)
            ngs.set_backend(be)
            ws, _ = make_gpu_workload(args.workload, cN, args.order)
            for _ in range(2):
                ws.step()
            g = GpuTimer(torch, be.lib, ws, torch.cuda.synchronize).clean(3)
            line['gpu_same_config'] = {'value': g['ms'] / 3e3, 'unit': 's', 'N': cN, 'dofs': ws.ndof,
                                       'linear_solver': workload_config(ref_args, 'gpu', N=cN, gpus=1)['linear_solver'],
                                       'gmres_its_per_step': g['its'] / 3}
            line['measured_ratio'] = per / (g['ms'] / 3e3)
            ngs.set_backend(None)
    except Exception as exc:                                        # pragma: no cover
        line['gpu_same_config'] = {'error': repr(exc)}
    print(json.dumps(line))


def main():
    args = parse()
    if args.impl == 'reference':
        return run_reference(args)
    return run_b200(args)


if __name__ == '__main__':
    main()
