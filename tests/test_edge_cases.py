"""Edge cases of the assembly path: smallest meshes, integrals over empty regions, forms without integrals, lowest
orders, a space without Dirichlet boundary. CPU: the oracle against closed-form values, and the CUDA backend's host
path on a null device (tests/test_gpu_paths_dry.py). GPU (collected in tests/test_zz_gpu_late_additions.py):
CUDA vs oracle."""
import numpy as np
import pytest

from opencmp_b200.mesh import structured_2d, structured_3d


def _with_backend(be, fn):
    import opencmp_b200.ngs as ngs
    old = ngs._backend
    ngs.set_backend(be)
    try:
        return fn(ngs)
    finally:
        ngs.set_backend(old)


def build_edge_case(ngs, name):
    """Returns (a, L, extra integrals) of one edge case; every builder is backend independent."""
    x, y = ngs.x, ngs.y
    if name == 'single_square_two_triangles_dg':
        m = ngs.Mesh(structured_2d([1, 1]))
        fes = ngs.FESpace([ngs.L2(m, order=1, dgjumps=True)], dgjumps=True)
    elif name == 'single_quad_q1':
        m = ngs.Mesh(structured_2d([1, 1], cell='quad'))
        fes = ngs.FESpace([ngs.H1(m, order=1)])
    elif name == 'single_hex_q2':
        m = ngs.Mesh(structured_3d([1, 1, 1]))
        fes = ngs.FESpace([ngs.H1(m, order=2, dirichlet='left')])
    elif name == 'l2_order0_dg':
        m = ngs.Mesh(structured_2d([3, 2]))
        fes = ngs.FESpace([ngs.L2(m, order=0, dgjumps=True)], dgjumps=True)
    elif name == 'pure_neumann_p2':
        m = ngs.Mesh(structured_2d([3, 3]))
        fes = ngs.FESpace([ngs.H1(m, order=2)])
    elif name == 'empty_boundary_region':
        m = ngs.Mesh(structured_2d([3, 3]))
        fes = ngs.FESpace([ngs.H1(m, order=2, dirichlet='left')])
    else:
        raise KeyError(name)
    u, v = fes.TrialFunction()[0], fes.TestFunction()[0]
    a = ngs.BilinearForm(fes)
    a += (ngs.InnerProduct(ngs.Grad(u), ngs.Grad(v)) + u * v) * ngs.dx
    L = ngs.LinearForm(fes)
    L += (1.0 + x) * v * ngs.dx
    if fes.dgjumps:
        n = ngs.specialcf.normal(m.dim)
        h = ngs.specialcf.mesh_size
        ju, jv = u - u.Other(), v - v.Other()
        a += (4.0 / h) * ju * jv * ngs.dx(skeleton=True)
        a += (4.0 / h) * u * v * ngs.ds(skeleton=True)
        L += (x + y) * v * ngs.ds(skeleton=True)
    elif name == 'empty_boundary_region':
        # a marker no facet carries: the integral contributes nothing and must not break the launch logic
        a += 3.0 * u * v * ngs.ds(definedon=m.Boundaries('no_such_marker'))
        L += 2.0 * v * ngs.ds(definedon=m.Boundaries('no_such_marker'))
        L += y * v * ngs.ds(definedon=m.Boundaries('right'))
    else:
        L += y * v * ngs.ds
    return m, fes, a, L


EDGE_CASES = ['single_square_two_triangles_dg', 'single_quad_q1', 'single_hex_q2', 'l2_order0_dg', 'pure_neumann_p2',
              'empty_boundary_region']


def assemble_edge_case(ngs, name):
    m, fes, a, L = build_edge_case(ngs, name)
    a.Assemble()
    L.Assemble()
    be = ngs.get_backend()
    vals = np.array(be.to_numpy(a.mat.values), dtype=np.float64).copy()
    rhs = L.vec.NumPy().copy()
    ones = ngs.BaseVector(be.from_numpy(np.ones(fes.ndof)))
    y1 = (a.mat * ones).NumPy().copy()
    vol = ngs.Integrate(ngs.CoefficientFunction(1.0), m)
    empty = ngs.LinearForm(fes)
    empty.Assemble()                                   # a form without integrals is a zero vector
    return dict(vals=vals, rhs=rhs, y1=y1, vol=vol, empty=empty.vec.NumPy().copy(), ndof=fes.ndof, dim=m.dim)


@pytest.mark.parametrize('name', EDGE_CASES)
def test_oracle_edge_cases_closed_form(name):
    from oracle.backend import OracleBackend
    r = _with_backend(OracleBackend(), lambda ngs: assemble_edge_case(ngs, name))
    assert r['rhs'].shape == (r['ndof'],) and np.all(r['empty'] == 0.0)
    assert abs(r['vol'] - 1.0) < 1e-13
    if name == 'single_quad_q1':
        # Q1 on the unit square: sum of all entries of stiffness + mass = area = 1; rhs sums to int (1 + x) + int_bnd y
        assert abs(r['y1'].sum() - 1.0) < 1e-13
        assert abs(r['rhs'].sum() - (1.5 + 2.0)) < 1e-13                 # int_bnd y ds = 0 + 1 + 1/2 + 1/2
    if name == 'l2_order0_dg':
        # piecewise constants: grad = 0, jumps of the constant vanish; only mass + boundary penalty remain
        assert r['vals'].shape[0] > 0 and abs(r['rhs'].sum() - (1.5 + 4.0)) < 1e-12   # int (1+x) + int_bnd (x + y)


def test_dry_cuda_host_path_handles_edge_cases():
    """Plan construction and launch loops of the CUDA backend for the same cases (empty item lists, one-cell meshes,
    forms without integrals) on the null device."""
    from test_gpu_paths_dry import DryCudaBackend
    import torch
    inv = torch.linalg.inv
    try:
        for name in EDGE_CASES:
            be = DryCudaBackend()
            r = _with_backend(be, lambda ngs: assemble_edge_case(ngs, name))
            assert r['rhs'].shape == (r['ndof'],)
            assert 'ocmp_contract_matrix' in be.lib.calls
    finally:
        torch.linalg.inv = inv
