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
