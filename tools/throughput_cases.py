"""Assembly Mnnz/s and SpMV GB/s of the SURVEY 8(d) throughput cases that bench.py does not time as a whole step
(run on the GPU box): python tools/throughput_cases.py [scale]

  Poisson H1 P2 / P3 (1024 x 1024 x 2 triangles), Stokes Taylor-Hood P3/P2 and Stokes / INS-Oseen HDiv-DG order 3
  (256 x 256 x 2), species transport DG order 2 with an Oseen wind (MCINS convection-diffusion block, 512 x 512 x 2),
  3-D Stokes Q2/Q1 hexes (32^3). Per case: a.Assemble() (deterministic two-phase assembly), L.Assemble(), mat * x,
  and — where the case is SPD — CG + multigrid iterations on the refined hierarchy.
"""
import os, sys, json, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
import opencmp_b200.ngs as ngs
from opencmp_b200.backend import CudaBackend
from opencmp_b200.mesh import structured_2d
import cases
be = CudaBackend(0); ngs.set_backend(be)
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0


def timeit(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def refined(n, levels, **kw):
    m = structured_2d([n >> levels, n >> levels], **kw)
    for _ in range(levels):
        m.Refine()
    return m


def channel(n, levels=0):
    m = structured_2d([n >> levels, n >> levels], scale=(1.0, 0.2))
    m.bnd_names = ['wall', 'outlet', 'wall', 'inlet']           # bottom, right, top, left
    for _ in range(levels):
        m.Refine()
    return m


N2 = int(1024 * scale) // 16 * 16
N1 = int(256 * scale) // 16 * 16
CASES = [
    ('Poisson H1 P2', lambda: cases.poisson(refined(N2, 4), 2, False), 'cg_mg'),
    ('Poisson H1 P3', lambda: cases.poisson(refined(N2, 4), 3, False), 'cg_mg'),
    ('Poisson L2-DG P2', lambda: cases.poisson(refined(N2 // 2, 4), 2, True, family='L2'), None),
    ('Stokes Taylor-Hood P3/P2', lambda: cases.stokes(channel(N1), 3, False), None),
    ('INS Oseen HDiv-DG order 3', lambda: cases.stokes(channel(N1), 3, True, wind=lambda n: cases.random_wind(n),
                                                       dt_val=0.01, mass=True), None),
    ('species DG P2 + wind (MCINS block)', lambda: cases.species(structured_2d([N1 * 2, N1 * 2]), 2,
                                                                 lambda n: cases.random_wind(n, 11)), None),
    ('Stokes 3-D Q2/Q1 hexes', lambda: cases.stokes_3d('hex', 2, n=int(32 * scale)), None),
]
for name, build, solve in CASES:
    t0 = time.time()
    c = build()
    a, L, fes = c['a'], c['L'], c['fes']
    a.Assemble(); L.Assemble()
    torch.cuda.synchronize()
    setup = time.time() - t0
    nnz, n = fes.pattern().nnz, fes.ndof
    ta = timeit(lambda: a.Assemble(), 3)
    tl = timeit(lambda: L.Assemble(), 3)
    x = a.mat.CreateColVector(); x.a.copy_(torch.rand(n, dtype=torch.float64, device='cuda'))
    y = a.mat.CreateColVector()
    ts = timeit(lambda: a.mat.Mult(x, y), 20)
    out = {'case': name, 'cells': c['mesh'].ne, 'dofs': n, 'nnz': nnz, 'setup_s': round(setup, 2),
           'assemble_ms': ta, 'assembly_mnnz_per_s': nnz / ta / 1e3, 'rhs_ms': tl, 'spmv_ms': ts,
           'spmv_gbs': (nnz * 12 + n * 20) / ts / 1e6}
    if solve == 'cg_mg':
        try:
            c['gfu'].components[0].Set(c['exact'], definedon=c['mesh'].Boundaries(c['dnames']))
            pre = ngs.Preconditioner(a, 'multigrid'); pre.Update()
            torch.cuda.synchronize(); t1 = time.time()
            ngs.solvers.CG(mat=a.mat, rhs=L.vec, pre=pre, sol=c['gfu'].vec, tol=1e-10, maxsteps=200, initialize=False)
            torch.cuda.synchronize()
            out['cg_multigrid'] = {'iterations': be.last_iters, 'solve_s': time.time() - t1}
        except Exception as exc:
            out['cg_multigrid'] = {'error': repr(exc)[:200]}
    print(json.dumps(out), flush=True)
    del c, a, L, fes, x, y
    torch.cuda.empty_cache()
