"""Generate golden fixtures by running the UNMODIFIED reference package (/root/reference/opencmp) through the
opencmp_b200 front end with the NumPy oracle as backend. Run in the build container only:

    python tests/golden/make_golden.py

Each fixture stores the mesh arrays (so it travels to the GPU box, where neither /root/reference nor its .vol files
exist), the DOF vector the reference's own model + solver code produced, and the error norms it printed.
"""
import contextlib
import io
import os
import shutil
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = '/root/reference'
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

import opencmp_b200.compat as compat          # noqa: E402
compat.install_as_ngsolve()
import opencmp_b200.ngs as ngs                # noqa: E402
from oracle.backend import OracleBackend      # noqa: E402
ngs.set_backend(OracleBackend())

from opencmp.config_functions import ConfigParser   # noqa: E402
from opencmp.models import get_model_class          # noqa: E402
from opencmp.solvers import get_solver_class        # noqa: E402
from opencmp.post_processing import run_post_processing  # noqa: E402


def run_reference(config_path, overrides):
    cfg = ConfigParser(config_path)
    for (sec, key), val in overrides.items():
        cfg[sec][key] = val
    dim_used = cfg.get_item(['DIM', 'diffuse_interface_method'], bool, quiet=True)
    model_class = get_model_class(cfg.get_item(['OTHER', 'model'], str), dim_used)
    solver_class = get_solver_class(cfg)
    out = io.StringIO()
    with contextlib.redirect_stdout(out):
        solver = solver_class(model_class, cfg)
        sol = solver.solve()
        run_post_processing(cfg, solver, sol)
    errs = {}
    for ln in out.getvalue().splitlines():
        if ' in ' in ln and ':' in ln and ('norm' in ln or 'divergence' in ln):
            k, v = ln.rsplit(':', 1)
            try:
                errs[k.strip()] = float(v)
            except ValueError:
                pass
        elif ln.startswith('divergence of'):
            k, v = ln.rsplit(':', 1)
            errs[k.strip()] = float(v)
    return solver, sol, errs


def mesh_arrays(mesh):
    return dict(points=mesh.points, cells=mesh.cells, bnd_facets=mesh.facets[mesh.bnd_facets],
                bnd_region=mesh.bnd_region, bnd_names=np.array(mesh.bnd_names))


def main():
    here = os.path.dirname(os.path.abspath(__file__))
    work = tempfile.mkdtemp()
    os.chdir(work)
    shutil.copytree(REF + '/pytests', 'pytests')
    cases = {
        'stokes_pipe_dg': ('pytests/full_system/stokes/stationary_pipe/config',
                           {('DG', 'DG'): 'True', ('FINITE ELEMENT SPACE', 'elements'): 'u -> HDiv\np -> L2'}),
        'stokes_pipe_cg': ('pytests/full_system/stokes/stationary_pipe/config', {}),
        # pytests/full_system/ins/test_ins.py:79-87 (Oseen, implicit Euler, HDiv-DG), shortened to three time steps
        'ins_sinusoidal_dg_3steps': ('pytests/full_system/ins/sinusoidal_transient/config',
                                     {('DG', 'DG'): 'True', ('FINITE ELEMENT SPACE', 'elements'): 'u -> HDiv\np -> L2',
                                      ('TRANSIENT', 'time_range'): '0.0, 0.003'}),
    }
    for name, (cfg, ov) in cases.items():
        if len(sys.argv) > 1 and name not in sys.argv[1:]:
            continue
        solver, sol, errs = run_reference(cfg, ov)
        m = solver.model.mesh
        np.savez_compressed(os.path.join(here, name + '.npz'), vec=np.asarray(sol.vec.NumPy()),
                            err_names=np.array(list(errs.keys())), err_vals=np.array(list(errs.values())),
                            **mesh_arrays(m))
        print(name, len(sol.vec), errs)
    # ---- lowered form programs of the reference's own model classes (replayed by tests/test_golden_programs.py) ---
    from opencmp_b200.serialize import dump
    from oracle import fem
    dg_ins = {('DG', 'DG'): 'True', ('FINITE ELEMENT SPACE', 'elements'): 'u -> HDiv\np -> L2'}
    short = {('TRANSIENT', 'time_range'): '0.0, 0.002'}
    prog_cases = {
        # INS (models/ins.py) — Oseen + Crank-Nicolson: Stokes terms, Oseen convection, DG facet upwinding, and the
        # "bilinear form applied to a known field" right-hand side of time_integration_schemes.py:135-190
        'prog_ins_dg_oseen_cn': ('pytests/full_system/ins/sinusoidal_transient/config',
                                 {**dg_ins, **short, ('TRANSIENT', 'scheme'): 'crank nicolson'}),
        # INS IMEX / CNLF (ins.py:300-321 explicit convection, schemes :249-304), Taylor-Hood CG
        'prog_ins_cg_imex_cnlf': ('pytests/full_system/ins/sinusoidal_transient/config',
                                  {**short, ('SOLVER', 'linearization_method'): 'IMEX', ('TRANSIENT', 'scheme'): 'CNLF',
                                   ('TRANSIENT', 'time_range'): '0.0, 0.004'}),
        # MultiComponentINS (models/multi_component_ins.py): HDiv-DG velocity + two L2 species, diffusion-convection
        'prog_mcins_dg_diffusion_convection': ('pytests/full_system/mcins/diffusion_convection/config',
                                               {('DG', 'DG'): 'True', ('FINITE ELEMENT SPACE', 'elements'):
                                                'u -> HDiv\np -> L2\na -> L2\nb -> L2', **short}),
        # MultiComponentINS CG with coupled first-order reactions (trial functions inside source terms, :245-256)
        'prog_mcins_cg_1st_rxn_coupled': ('pytests/full_system/mcins/1st_rxn_coupled/config', dict(short)),
        # Poisson DG transient (models/poisson.py:71-158 incl. interior-facet SIP and Nitsche boundary terms)
        'prog_poisson_dg_transient': ('pytests/full_system/poisson/transient_coarse/config',
                                      {('DG', 'DG'): 'True', **short}),
        # stationary INS in a pipe, stress boundary conditions, Oseen + Anderson mixing (ins.py:208-223)
        'prog_ins_cg_stationary_stress': ('pytests/full_system/ins/pressure_flow_in_pipe_stress/config', {}),
        # PoissonDIM (models/poisson_dim.py) with the phase field and the two boundary masks the reference ships as
        # NGSolve-binary .sol files (pytests/full_system/dim/dim_poisson_2: 39 x 39 quads, H1 order 3, DIM Dirichlet
        # data on 'top' / 'bottom' of a diffuse disc) — DIM.get_DIM_gridfunctions load_method = file (dim.py:348-378)
        'prog_poisson_dim_2': ('pytests/full_system/dim/dim_poisson_2/config', {}),
    }
    # INSDIM (models/ins_dim.py: the form set of BASELINE configs[4]) — the reference ships no INS-DIM case, so its
    # StokesDIM test (pytests/full_system/dim/dim_stokes_1: HDiv-DG order 2 / L2 order 1 on 29 x 29 quadrilaterals,
    # diffuse circle generated from circle_nd.stl by the reference's own pre-processing) is switched to the INS model:
    # Oseen linearisation, implicit Euler, two time steps, nu = 0.1
    src, dst = 'pytests/full_system/dim/dim_stokes_1', 'pytests/full_system/dim/dim_ins_1'
    if not os.path.isdir(dst):
        shutil.copytree(src, dst)
        cfg_txt = open(dst + '/config').read().replace('dim_stokes_1', 'dim_ins_1').replace('model = Stokes', 'model = INS')
        cfg_txt = cfg_txt.replace('transient = False', 'transient = True\nscheme = implicit euler\n'
                                  'time_range = 0.0, 0.02\ndt = 1e-2')
        cfg_txt = cfg_txt.replace('[SOLVER]', '[SOLVER]\nlinearization_method = Oseen\nnonlinear_max_iterations = 3\n'
                                  'nonlinear_tolerance = relative -> 1e-6\n                      absolute -> 1e-8')
        open(dst + '/config', 'w').write(cfg_txt)
        mc = open(dst + '/model_dir/model_config').read().replace('all -> 0.001', 'all -> 0.1')
        open(dst + '/model_dir/model_config', 'w').write(mc)
        ic = open(dst + '/ic_dir/ic_config').read().replace('[STOKES]', '[INS]')
        open(dst + '/ic_dir/ic_config', 'w').write(ic)
    prog_cases['prog_ins_dim_dg_quad'] = (dst + '/config', {})
    only = set(sys.argv[1:])
    if only:
        prog_cases = {k: v for k, v in prog_cases.items() if k in only}
    rng = np.random.default_rng(7)
    for name, (cfg, ov) in prog_cases.items():
        try:
            solver, sol, errs = run_reference(cfg, ov)
        except SystemExit:
            pass
        a, L = solver.a[0], solver.L[0]
        a.Assemble()
        L.Assemble()
        A = fem.csr_matrix(a.space, a.mat.values)
        X = rng.uniform(-1.0, 1.0, (3, a.space.ndof))
        Y = np.stack([A @ x for x in X])
        dump(os.path.join(here, name + '.npz'), {'a': a.program(), 'L': L.program()},
             dict(X=X, Y=Y, rhs=np.asarray(L.vec.NumPy()), diag=A.diagonal(), absmax=np.abs(a.mat.values).max()))
        print(name, a.space.ndof, A.nnz, [i.kind for i in a.program().integrals], os.path.getsize(
            os.path.join(here, name + '.npz')))
    shutil.rmtree(work, ignore_errors=True)


if __name__ == '__main__':
    main()
