"""Install the opencmp_b200 front end under the import names OpenCMP uses:

    import opencmp_b200.compat as c; c.install_as_ngsolve()
    import opencmp            # the unmodified reference package now runs on the B200 backend

Aliases: ``ngsolve`` (+ ``ngsolve.comp``, ``ngsolve.solvers``, ``ngsolve.config``), ``pyngcore`` / ``ngsolve.ngstd``
(TaskManager, SetNumThreads, BitArray), ``netgen.meshing`` / ``netgen.read_gmsh`` (mesh readers only). This is the
reference-side binding a maintainer would add in ``opencmp/__init__.py``.
"""
from __future__ import annotations

import sys
import types


def install_as_ngsolve(force: bool = False, device_mixing: bool = False) -> None:
    """``device_mixing``: also register opencmp_b200.mixing as ``opencmp.solvers.nonlinear_mixing`` (same interface as
    the reference module, history vectors kept on the device). Must happen before ``import opencmp``."""
    if device_mixing:
        from . import mixing
        sys.modules['opencmp.solvers.nonlinear_mixing'] = mixing
    if 'ngsolve' in sys.modules and not force and getattr(sys.modules['ngsolve'], '__b200__', False):
        return
    from . import ngs
    from .mesh import load_mesh, read_msh
    ngs.__b200__ = True
    sys.modules['ngsolve'] = ngs
    comp = types.ModuleType('ngsolve.comp')
    for name in ('ProxyFunction', 'FESpace', 'DifferentialSymbol', 'GridFunction', 'BilinearForm', 'LinearForm',
                 'Preconditioner', 'Mesh', 'Region', 'SumOfIntegrals'):
        setattr(comp, name, getattr(ngs, name))
    ngs.comp = comp
    sys.modules['ngsolve.comp'] = comp
    solv = types.ModuleType('ngsolve.solvers')
    for name in ('CG', 'MinRes', 'GMRes', 'PreconditionedRichardson'):
        setattr(solv, name, getattr(ngs.solvers, name))
    sys.modules['ngsolve.solvers'] = solv
    core = types.ModuleType('pyngcore')
    core.TaskManager, core.SetNumThreads, core.BitArray = ngs.TaskManager, ngs.SetNumThreads, ngs.BitArray
    sys.modules['pyngcore'] = core
    ngs.ngstd = core
    sys.modules['ngsolve.ngstd'] = core
    ngs.ngsglobals = ngs.ngsglobals
    netgen = types.ModuleType('netgen')
    meshing = types.ModuleType('netgen.meshing')
    # ``ngmsh.Mesh(dim)`` + ``.Load(filename)`` (reference helpers/io.py:94-99) and the builder calls of
    # diffuse_interface/mesh_helpers.py:494-690
    from . import netgen_shim
    for name in ('Mesh', 'MeshPoint', 'Pnt', 'Element1D', 'Element2D', 'Element3D', 'FaceDescriptor', 'PointId'):
        setattr(meshing, name, getattr(netgen_shim, name))
    read_gmsh = types.ModuleType('netgen.read_gmsh')
    read_gmsh.ReadGmsh = lambda filename: read_msh(filename if filename.endswith('.msh') else filename + '.msh')
    if 'edt' not in sys.modules:
        try:
            import edt  # noqa: F401
        except ImportError:
            edt_mod = types.ModuleType('edt')
            edt_mod.edt = netgen_shim.edt
            sys.modules['edt'] = edt_mod
    netgen.meshing, netgen.read_gmsh = meshing, read_gmsh
    sys.modules['netgen'] = netgen
    sys.modules['netgen.meshing'] = meshing
    sys.modules['netgen.read_gmsh'] = read_gmsh
    if 'pyparsing' not in sys.modules:
        try:
            import pyparsing  # noqa: F401
        except ImportError:
            try:
                from pip._vendor import pyparsing as _pp
                sys.modules['pyparsing'] = _pp
            except ImportError:
                pass


def use_device_dim() -> None:
    """Route the voxel pipeline of the reference's diffuse-interface pre-processing to the device: after
    ``import opencmp``, ``opencmp.diffuse_interface.interface.get_binary_2d`` and ``.get_phi`` (interface.py:31-57,
    :137-180) are replaced by the functions of the same name and signature in ``opencmp_b200.dimgen`` (ray tracing,
    erosion, exact distance transform and erf profile in csrc/ocmp_dim.cu). Needs the CUDA backend to be active when
    they are called. ``edt.edt`` alone is routed to the device already by ``install_as_ngsolve()`` whenever the CUDA
    backend is active."""
    import importlib
    from . import dimgen
    interface = importlib.import_module('opencmp.diffuse_interface.interface')
    interface.get_binary_2d = dimgen.get_binary_2d
    interface.get_phi = dimgen.get_phi
    # the per-time-step node loop of moving interfaces (helpers/ngsolve_.py:212-296) -> one launch where it applies
    helpers = importlib.import_module('opencmp.helpers.ngsolve_')
    original = helpers.gridfunction_rigid_body_motion
    if not getattr(original, '__b200__', False):
        def moved(t, orig_gfu, gfu, inv_R, mesh, N, scale, offset):
            return dimgen.gridfunction_rigid_body_motion(t, orig_gfu, gfu, inv_R, mesh, N, scale, offset,
                                                         fallback=original)
        moved.__b200__ = True
        helpers.gridfunction_rigid_body_motion = moved
