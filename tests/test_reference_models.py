"""Run the UNMODIFIED reference package through the drop-in boundary (opencmp_b200.compat.install_as_ngsolve) with
the oracle as backend: its own model / solver / config code and its own pytest cases. Only possible where
/root/reference is mounted (the build container); skipped elsewhere. The full-suite outcome of this round is
recorded in tests/golden/reference_suite_results.md."""
import os
import shutil
import subprocess
import sys
import textwrap

import pytest

REF = '/root/reference'
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.skipif(not os.path.isdir(REF + '/opencmp'), reason='reference tree not mounted')

CONFTEST = textwrap.dedent('''
    import sys
    sys.path.insert(0, {root!r})
    import opencmp_b200.compat as c
    c.install_as_ngsolve()
    import opencmp_b200.ngs as ngs
    from oracle.backend import OracleBackend
    ngs.set_backend(OracleBackend())
''')


@pytest.fixture(scope='module')
def ref_tree(tmp_path_factory):
    d = tmp_path_factory.mktemp('ref')
    shutil.copytree(REF + '/pytests', d / 'pytests')
    shutil.copytree(REF + '/examples/Poisson', d / 'examples' / 'Poisson')
    os.symlink(REF + '/opencmp', d / 'opencmp')
    (d / 'conftest.py').write_text(CONFTEST.format(root=ROOT))
    return d


def _pytest(tree, *args):
    cmd = [sys.executable, '-m', 'pytest', '-q', '-p', 'no:cacheprovider', *args]
    return subprocess.run(cmd, cwd=tree, capture_output=True, text=True, timeout=1500)


def test_reference_stokes_suite_passes_through_the_boundary(ref_tree):
    """pytests/full_system/stokes: Taylor-Hood P3/P2 and HDiv-DG/L2 Poiseuille, 7 golden error norms each."""
    r = _pytest(ref_tree, 'pytests/full_system/stokes')
    assert r.returncode == 0, r.stdout[-3000:]
    assert '2 passed' in r.stdout


def test_reference_ins_transient_cases_pass_through_the_boundary(ref_tree):
    """pytests/full_system/ins: Oseen implicit Euler (CG and DG) and IMEX (CNLF) on the Taylor-Green problem."""
    r = _pytest(ref_tree, 'pytests/full_system/ins/test_ins.py', '-k',
                'implicit_euler or CNLF')
    assert r.returncode == 0, r.stdout[-3000:]
    assert 'passed' in r.stdout and 'failed' not in r.stdout


def test_reference_poisson_example_h_convergence(ref_tree):
    """examples/Poisson/config as shipped: H1 order 3, five uniform refinements, rate p + 1 = 4."""
    script = textwrap.dedent('''
        import sys
        sys.path.insert(0, {root!r}); sys.path.insert(0, {tree!r})
        import conftest
        from opencmp.run import run
        run('config')
    ''').format(root=ROOT, tree=str(ref_tree))
    r = subprocess.run([sys.executable, '-c', script], cwd=ref_tree / 'examples' / 'Poisson', capture_output=True,
                       text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    rows = [ln.split() for ln in r.stdout.splitlines() if ln.startswith('1/')]
    rates = [float(x[-1]) for x in rows]
    assert len(rates) == 5 and all(abs(x - 4.0) < 0.1 for x in rates[1:]), r.stdout[-1500:]


def test_reference_ins_example_with_all_error_metrics(ref_tree, tmp_path):
    """examples/INS (BASELINE configs[2]) as shipped — Taylor-Green, HDiv-DG order 3, Oseen, adaptive two step — with its
    ref_sol_config, which switches on every metric of helpers/error.py incl. ``facet_jumps``
    (``Integrate(... * dx(element_boundary=True))``, helpers/error.py:146). Only the end time is shortened."""
    d = tmp_path / 'INS'
    shutil.copytree(REF + '/examples/INS', d)
    cfg = (d / 'config').read_text().replace('time_range = 0.0, 0.1', 'time_range = 0.0, 0.02')
    assert 'time_range = 0.0, 0.02' in cfg
    (d / 'config').write_text(cfg)
    script = textwrap.dedent('''
        import sys
        sys.path.insert(0, {root!r}); sys.path.insert(0, {tree!r})
        import conftest
        from opencmp.run import run
        run('config')
    ''').format(root=ROOT, tree=str(ref_tree))
    r = subprocess.run([sys.executable, '-c', script], cwd=d, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    vals = {}
    for ln in r.stdout.splitlines():
        if ':' in ln and any(k in ln for k in ('norm in', 'divergence of', 'jump of')):
            key, val = ln.rsplit(':', 1)
            vals[key.strip()] = float(val)
    assert len(vals) == 9, r.stdout[-1500:]
    assert vals['l2 norm in u'] < 1e-3 and vals['divergence of u'] < 1e-9          # HDiv: pointwise divergence free
    assert 0.0 < vals['magnitude of jump of u facets'] < 1e-3                       # tangential jumps only, small
    assert 1e-3 < vals['magnitude of jump of p facets'] < 1.0                        # L2 pressure is discontinuous


def test_reference_stokes_example_as_shipped(ref_tree, tmp_path):
    """examples/Stokes (BASELINE configs[1]) exactly as shipped: channel_3bcs.vol, HDiv order 3 / L2 order 2, DG,
    direct solve, Poiseuille reference solution — errors at round-off level like the reference's golden values."""
    d = tmp_path / 'Stokes'
    shutil.copytree(REF + '/examples/Stokes', d)
    script = textwrap.dedent('''
        import sys
        sys.path.insert(0, {root!r}); sys.path.insert(0, {tree!r})
        import conftest
        from opencmp.run import run
        run('config')
    ''').format(root=ROOT, tree=str(ref_tree))
    r = subprocess.run([sys.executable, '-c', script], cwd=d, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    vals = {}
    for ln in r.stdout.splitlines():
        if ' norm in ' in ln or ln.startswith('divergence of'):
            key, val = ln.rsplit(':', 1)
            vals[key.strip()] = float(val)
    assert vals['l2 norm in u'] < 1e-8 and vals['l2 norm in p'] < 1e-8 and vals['divergence of u'] < 1e-8


def test_reference_mcins_example(ref_tree, tmp_path):
    """examples/MCINS (BASELINE configs[3]): MultiComponentINS, VectorH1/H1 order 3 + two H1 species with zeroth-order
    sources, IMEX / euler IMEX, .vtu output. The shipped config lacks the nonlinear-solver keys that
    solvers/base_solver.py:286-292 reads for every INS-type model (the reference stops with its own ValueError on it), so
    they are added; the end time is shortened. da/dt = -0.1, db/dt = +0.1 from a = 1, b = 0 is reproduced exactly."""
    d = tmp_path / 'MCINS'
    shutil.copytree(REF + '/examples/MCINS', d)
    cfg = (d / 'config').read_text()
    cfg = cfg.replace('linearization_method = IMEX', 'linearization_method = IMEX\nnonlinear_solver = default\n'
                      'nonlinear_tolerance = relative -> 1e-4\n                      absolute -> 1e-6\n'
                      'nonlinear_max_iterations = 3')
    cfg = cfg.replace('time_range = 0.0, 10', 'time_range = 0.0, 0.03')
    (d / 'config').write_text(cfg)
    script = textwrap.dedent('''
        import sys
        sys.path.insert(0, {root!r}); sys.path.insert(0, {tree!r})
        import conftest
        from opencmp.run import run
        run('config')
    ''').format(root=ROOT, tree=str(ref_tree))
    r = subprocess.run([sys.executable, '-c', script], cwd=d, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    vals = {}
    for ln in r.stdout.splitlines():
        if ln.startswith('l2 norm in') or ln.startswith('divergence of'):
            key, val = ln.rsplit(':', 1)
            vals[key.strip()] = float(val)
    # reference "solution" of the example is the state at t = 10 (a = 0, b = 1): at t = 0.03 the L2 distance on the
    # unit square is 1 - 0.1 * 0.03 for both species
    assert abs(vals['l2 norm in a'] - 0.997) < 1e-9 and abs(vals['l2 norm in b'] - 0.997) < 1e-9
    assert vals['l2 norm in u'] < 1e-12 and vals['divergence of u'] < 1e-12


def test_reference_vtu_post_processing_through_the_boundary(ref_tree):
    """`save_type = .vtu`: the reference's own post-processing (post_processing/output_conversions.py — a
    multiprocessing.Pool of workers loading every .sol file and calling VTKOutput(...).Do(), then writing the .pvd
    collection) runs unmodified on our GridFunction / VTKOutput."""
    case = ref_tree / 'pytests' / 'full_system' / 'restart' / 'transient_poisson'
    cfg = (case / 'config').read_text()
    cfg = cfg.replace('save_type = .sol', 'save_type = .vtu').replace('time_range = 0.0, 0.1', 'time_range = 0.0, 0.03')
    cfg = cfg.replace('run_dir = .', 'run_dir = pytests/full_system/restart/transient_poisson')
    (case / 'config_vtu').write_text(cfg)
    script = textwrap.dedent('''
        import sys
        sys.path.insert(0, {root!r}); sys.path.insert(0, {tree!r})
        import conftest
        from opencmp.run import run
        run('pytests/full_system/restart/transient_poisson/config_vtu')
    ''').format(root=ROOT, tree=str(ref_tree))
    r = subprocess.run([sys.executable, '-c', script], cwd=ref_tree, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    out = case / 'output'
    vtus = sorted((out / 'poisson_vtu').glob('*.vtu'))
    assert len(vtus) == 4                                   # t = 0, 0.01, 0.02, 0.03
    pvd = (out / 'poisson_transient.pvd').read_text()
    assert pvd.count('<DataSet') == 4 and 'poisson_vtu/poisson_0.01.vtu' in pvd
    import xml.etree.ElementTree as ET
    piece = ET.parse(vtus[1]).getroot().find('UnstructuredGrid/Piece')
    names = [d.get('Name') for d in piece.findall('PointData/DataArray')]
    assert names == ['u'] and int(piece.get('NumberOfCells')) > 0


def test_reference_dim_poisson_with_shipped_phase_field(ref_tree):
    """pytests/full_system/dim/test_dim.py::test_DIM_poisson_2 — the reference's DIM class loads the phase field and
    the two boundary masks it ships as NGSolve-binary .sol files (H1 order 3 on 39 x 39 quads; their high-order part
    vanishes, so the vertex values determine them, see GridFunction._from_ngsolve_binary), PoissonDIM builds the
    phi-weighted forms (models/poisson_dim.py) and the stationary solver runs. Then the phi-weighted L2 error against
    the manufactured solution of its ref_sol_config is measured: 1.1e-3 relative."""
    r = _pytest(ref_tree, 'pytests/full_system/dim/test_dim.py', '-k', 'poisson_2')
    assert r.returncode == 0, r.stdout[-3000:]
    assert '1 passed' in r.stdout
    script = textwrap.dedent('''
        import sys
        sys.path.insert(0, {root!r}); sys.path.insert(0, {tree!r})
        import conftest
        import numpy as np
        import opencmp_b200.ngs as ngs
        from opencmp.config_functions import ConfigParser
        from opencmp.models import get_model_class
        from opencmp.solvers import get_solver_class
        cfg = ConfigParser('pytests/full_system/dim/dim_poisson_2/config')
        solver = get_solver_class(cfg)(get_model_class('Poisson', True), cfg)
        sol = solver.solve()
        m = solver.model
        phi = m.DIM_solver.phi_gfu
        u = sol.components[0] if sol.components else sol
        ref = ngs.sin(np.pi * ngs.x) * ngs.cos(np.pi * ngs.y)
        err = np.sqrt(ngs.Integrate((u - ref) * (u - ref) * phi, m.mesh))
        nrm = np.sqrt(ngs.Integrate(ref * ref * phi, m.mesh))
        print('RESULT', type(m).__name__, sorted(m.DIM_solver.mask_gfu_dict), err / nrm, ngs.Integrate(phi, m.mesh))
    ''').format(root=ROOT, tree=str(ref_tree))
    r = subprocess.run([sys.executable, '-c', script], cwd=ref_tree, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith('RESULT')][0].split()
    assert line[1] == 'PoissonDIM' and "'bottom'," in line[2] + line[3]
    assert float(line[-2]) < 3e-3                    # phi-weighted relative L2 error
    assert abs(float(line[-1]) - 2.865) < 0.01       # area of the diffuse disc


def test_ngsolve_binary_sol_with_high_order_part_is_refused(ref_tree):
    """dim_stokes_2/phi.sol carries non-zero edge / cell DOFs in NGSolve's basis: refused loudly, not misread."""
    script = textwrap.dedent('''
        import sys
        sys.path.insert(0, {root!r}); sys.path.insert(0, {tree!r})
        import conftest
        import opencmp_b200.ngs as ngs
        m = ngs.Mesh('pytests/full_system/dim/dim_stokes_2/dim_dir/mesh.vol')
        g = ngs.GridFunction(ngs.H1(m, order=2))
        try:
            g.Load('pytests/full_system/dim/dim_stokes_2/dim_dir/phi.sol')
        except NotImplementedError as e:
            print('REFUSED', e)
    ''').format(root=ROOT, tree=str(ref_tree))
    r = subprocess.run([sys.executable, '-c', script], cwd=ref_tree, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and 'REFUSED' in r.stdout, r.stdout + r.stderr[-2000:]


def test_reference_dim_generation_pipeline_through_the_boundary(ref_tree):
    """pytests/full_system/dim/test_dim.py, the `load_method = generate` cases: the reference's own DIM pre-processing
    (structured mesh built through the netgen.meshing builder calls of mesh_helpers.get_Netgen_nonconformal, STL
    boundary -> ray tracing -> Euclidean distance transform (`edt`) -> erf profile -> VoxelCoefficient ->
    GridFunction, boundary masks, and — dim_poisson_5 — a phase field generated on a 59 x 59 grid projected onto the
    19 x 19 simulation mesh) followed by the PoissonDIM solve, all unmodified."""
    r = _pytest(ref_tree, 'pytests/full_system/dim/test_dim.py', '-k', 'poisson or stokes_1')
    assert r.returncode == 0, r.stdout[-3000:]
    assert '7 passed' in r.stdout        # dim_stokes_1: StokesDIM, HDiv-DG order 2 / L2 order 1 on quadrilaterals


def test_workload_forms_equal_reference_insdim(ref_tree):
    """The bench workload of BASELINE configs[4] restates the INSDIM forms (opencmp_b200/workloads.ins_dim_cg_forms).
    Here the reference's own INSDIM class — conforming Taylor-Hood elements on the quadrilateral DIM mesh, circle from
    circle_nd.stl, Oseen + implicit Euler, non-zero DIM Dirichlet data — assembles its system for the second time step,
    and the restated forms, given the same mesh, phase field, mask, wind and previous solution, must give the same
    matrix and right-hand side entry by entry (1e-12 of the largest entry)."""
    import re
    src = ref_tree / 'pytests' / 'full_system' / 'dim' / 'dim_stokes_1'
    dst = ref_tree / 'pytests' / 'full_system' / 'dim' / 'dim_ins_cg'
    if not dst.exists():
        shutil.copytree(src, dst)
        cfg = (dst / 'config').read_text().replace('dim_stokes_1', 'dim_ins_cg').replace('model = Stokes', 'model = INS')
        cfg = cfg.replace('u -> HDiv', 'u -> VectorH1').replace('p -> L2', 'p -> H1').replace('DG = True', 'DG = False')
        cfg = cfg.replace('transient = False', 'transient = True\nscheme = implicit euler\ntime_range = 0.0, 0.02\n'
                          'dt = 1e-2')
        cfg = cfg.replace('[SOLVER]', '[SOLVER]\nlinearization_method = Oseen\nnonlinear_max_iterations = 2\n'
                          'nonlinear_tolerance = relative -> 1e-6\n                      absolute -> 1e-8')
        (dst / 'config').write_text(cfg)
        mc = (dst / 'model_dir' / 'model_config').read_text().replace('all -> 0.001', 'all -> 0.1')
        (dst / 'model_dir' / 'model_config').write_text(mc)
        (dst / 'ic_dir' / 'ic_config').write_text((dst / 'ic_dir' / 'ic_config').read_text().replace('[STOKES]', '[INS]'))
        bc = (dst / 'dim_dir' / 'bc_dir' / 'dim_bc_config').read_text()
        (dst / 'dim_dir' / 'bc_dir' / 'dim_bc_config').write_text(bc.replace('[0.0, 0.0]', '[0.3*y, -0.3*x]'))
    script = textwrap.dedent('''
        import sys
        sys.path.insert(0, {root!r}); sys.path.insert(0, {tree!r})
        import conftest
        import numpy as np
        import opencmp_b200.ngs as ngs
        from opencmp_b200.workloads import ins_dim_cg_forms
        from opencmp.config_functions import ConfigParser
        from opencmp.models import get_model_class
        from opencmp.solvers import get_solver_class
        cfg = ConfigParser('pytests/full_system/dim/dim_ins_cg/config')
        solver = get_solver_class(cfg)(get_model_class('INS', True), cfg)
        solver.solve()
        m = solver.model
        assert type(m).__name__ == 'INSDIM' and m.mesh.cell_type == 'quad'
        a_ref, L_ref = solver.a[0], solver.L[0]
        a_ref.Assemble(); L_ref.Assemble()
        k = m.interp_ord
        x, y = ngs.x, ngs.y
        phi, mask = m.DIM_solver.phi_gfu, m.DIM_solver.mask_gfu_dict['all']
        w = m.W[0] if isinstance(m.W, (list, tuple)) else m.W
        gfu_0 = solver.gfu_0_list[0] if hasattr(solver, 'gfu_0_list') else solver.gfu_0
        h = ngs.specialcf.mesh_size
        alpha = (10.0 * k ** 2) / h
        dt = solver.dt_param[0] if hasattr(solver, 'dt_param') else ngs.Parameter(1e-2)
        D = m.DIM_solver
        a, L = ins_dim_cg_forms(m.fes, phi, D.grad_phi_gfu, D.mag_grad_phi_gfu, mask, w, gfu_0,
                                ngs.CoefficientFunction((0.3 * y, -0.3 * x)),
                                ngs.CoefficientFunction((0.0, 0.0)), ngs.CoefficientFunction(0.1), alpha, dt)
        a.Assemble(); L.Assemble()
        A, B = np.asarray(a.mat.values), np.asarray(a_ref.mat.values)
        print('RESULT', m.fes.ndof, len(A), np.abs(A - B).max() / np.abs(B).max(),
              np.abs(L.vec.NumPy() - L_ref.vec.NumPy()).max() / np.abs(L_ref.vec.NumPy()).max(),
              float(np.abs(w.vec.NumPy()).max()))
    ''').format(root=ROOT, tree=str(ref_tree))
    r = subprocess.run([sys.executable, '-c', script], cwd=ref_tree, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith('RESULT')][0].split()
    assert float(line[3]) < 1e-12 and float(line[4]) < 1e-12, line
    assert float(line[5]) > 1e-3          # the Oseen wind is not trivially zero


def test_use_device_dim_patches_the_reference_pre_processing():
    """compat.use_device_dim(): the reference's own get_binary_2d / get_phi / gridfunction_rigid_body_motion are
    replaced by the device versions of the same name and signature (subprocess: the patch is process-wide)."""
    code = textwrap.dedent('''
        import sys, inspect
        sys.path.insert(0, {root!r}); sys.path.insert(0, {ref!r})
        import opencmp_b200.compat as c
        c.install_as_ngsolve()
        import opencmp
        from opencmp.diffuse_interface import interface
        from opencmp.helpers import ngsolve_ as helpers
        before = (inspect.signature(interface.get_phi), inspect.signature(interface.get_binary_2d),
                  inspect.signature(helpers.gridfunction_rigid_body_motion))
        c.use_device_dim()
        c.use_device_dim()                                  # idempotent
        from opencmp_b200 import dimgen
        assert interface.get_phi is dimgen.get_phi and interface.get_binary_2d is dimgen.get_binary_2d
        after = (inspect.signature(interface.get_phi), inspect.signature(interface.get_binary_2d),
                 inspect.signature(helpers.gridfunction_rigid_body_motion))
        for b, a in zip(before, after):
            assert list(b.parameters) == list(a.parameters), (b, a)
        try:
            interface.get_phi(None, 0.1, [4, 4], [1.0, 1.0], [0.0, 0.0], 2)
        except RuntimeError as exc:
            assert 'CUDA backend' in str(exc)               # no CPU fallback behind the device entry points
        else:
            raise AssertionError('dimgen must refuse to run without the CUDA backend')
        print('ok')
    ''').format(root=ROOT, ref=REF)
    p = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0 and p.stdout.strip().endswith('ok'), p.stderr[-2000:]
