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
        solver, sol, errs = run_reference(cfg, ov)
        m = solver.model.mesh
        np.savez_compressed(os.path.join(here, name + '.npz'), vec=np.asarray(sol.vec.NumPy()),
                            err_names=np.array(list(errs.keys())), err_vals=np.array(list(errs.values())),
                            **mesh_arrays(m))
        print(name, len(sol.vec), errs)
    shutil.rmtree(work, ignore_errors=True)


if __name__ == '__main__':
    main()
