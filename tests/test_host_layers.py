"""CPU tests of the host export layer: meshes, quadrature, bases, DOF numbering, sparsity, symbolic lowering."""
import os
from math import factorial as f

import numpy as np
import pytest

from opencmp_b200.basis import make_basis
from opencmp_b200.mesh import delaunay_rectangle, read_vol, structured_2d, structured_3d
from opencmp_b200.quadrature import cell_rule, facet_rule_in_cell
from opencmp_b200 import space as sp

REF = '/root/reference'
have_ref = os.path.isdir(REF)


def test_structured_generators_follow_reference_numbering():
    # reference diffuse_interface/mesh_helpers.py:524-549: node (i, j) -> i*(Nx+1)+j, names bottom,right,top,left
    m = structured_2d([4, 3], scale=(4.0, 3.0), offset=(2.0, 1.0), cell='quad')
    assert m.ne == 12 and m.nv == 20
    assert np.allclose(m.points[1 * 5 + 2], [-2.0 + 2.0, -1.0 + 1.0])
    assert m.bnd_names == ['bottom', 'right', 'top', 'left']
    assert np.bincount(m.bnd_region).tolist() == [4, 3, 4, 3]
    m.check_affine()
    h = structured_3d([3, 2, 2])
    assert h.ne == 12 and h.bnd_names == ['back', 'left', 'front', 'right', 'bottom', 'top']
    assert np.bincount(h.bnd_region).tolist() == [4, 6, 4, 6, 6, 6]
    t = structured_3d([2, 2, 2], cell='tet')
    assert t.ne == 48 and abs(np.abs(np.linalg.det(t.jacobians())).sum() / 6 - 1.0) < 1e-14


@pytest.mark.parametrize('ct,deg', [('seg', 7), ('tri', 6), ('tet', 5), ('quad', 5), ('hex', 3)])
def test_quadrature_exactness(ct, deg):
    p, w = cell_rule(ct, deg)
    if ct == 'tri':
        err = max(abs((w * p[:, 0] ** a * p[:, 1] ** b).sum() - f(a) * f(b) / f(a + b + 2))
                  for a in range(deg + 1) for b in range(deg + 1 - a))
    elif ct == 'tet':
        err = max(abs((w * p[:, 0] ** a * p[:, 1] ** b * p[:, 2] ** c).sum() - f(a) * f(b) * f(c) / f(a + b + c + 3))
                  for a in range(deg + 1) for b in range(deg + 1 - a) for c in range(deg + 1 - a - b))
    else:
        d = p.shape[1]
        err = max(abs((w * np.prod(p ** a, axis=1)).sum() - (1.0 / (a + 1)) ** d) for a in range(deg + 1))
    assert err < 1e-14


@pytest.mark.parametrize('fam,ct,order,ndof', [('H1', 'tri', 2, 6), ('H1', 'tri', 3, 10), ('L2', 'tri', 2, 6),
                                                ('HDiv', 'tri', 3, 20), ('HDiv', 'tri', 1, 6), ('H1', 'hex', 2, 27),
                                                ('H1', 'quad', 3, 16), ('H1', 'tet', 2, 10), ('L2', 'tet', 1, 4)])
def test_local_dof_counts_and_gradients(fam, ct, order, ndof):
    b = make_basis(fam, ct, order)
    assert b.ndof == ndof                       # P2 = 6, P3 = 10, BDM3 = 20, L2 P2 = 6
    dim = b.dim
    rng = np.random.default_rng(0)
    pts = rng.uniform(0.1, 0.3, (4, dim))
    h = 1e-5
    T = b.tabulate(pts)
    nv = 1 if b.kind == 'scalar' else dim
    for a in range(dim):
        e = np.zeros(dim)
        e[a] = h
        fd = (b.tabulate(pts + e)[:, :nv] - b.tabulate(pts - e)[:, :nv]) / (2 * h)
        for c in range(nv):
            row = nv + c * dim + a
            assert np.abs(fd[:, c] - T[:, row]).max() < 1e-6 * max(1.0, np.abs(T[:, row]).max())


def test_h1_partition_of_unity_and_l2_orthogonality():
    for ct in ('tri', 'quad', 'tet', 'hex'):
        b = make_basis('H1', ct, 1)
        p, w = cell_rule(ct, 3)
        assert np.allclose(b.tabulate(p)[:, 0, :].sum(axis=1), 1.0)
    b = make_basis('L2', 'tri', 3)
    p, w = cell_rule('tri', 8)
    T = b.tabulate(p)[:, 0, :]
    M = T.T @ (w[:, None] * T)
    assert np.abs(M - np.diag(np.diag(M))).max() < 1e-14


def test_global_dof_counts_match_survey():
    # Stokes example 14712 = 10356 HDiv + 4356 L2 on 726 triangles; formulas on structured meshes
    m = structured_2d([6, 5])
    V, E, T = m.nv, m.nf, m.ne
    assert sp.H1(m, order=2).ndof == V + E
    assert sp.H1(m, order=3).ndof == V + 2 * E + T
    assert sp.L2(m, order=2).ndof == 6 * T
    assert sp.HDiv(m, order=3).ndof == 4 * E + 8 * T
    assert sp.VectorH1(m, order=3).ndof == 2 * (V + 2 * E + T)
    h = structured_3d([3, 3, 3])
    assert sp.H1(h, order=2).ndof == 7 ** 3
    if have_ref:
        c = read_vol(REF + '/examples/Stokes/channel_3bcs.vol')
        X = sp.FESpace([sp.HDiv(c, order=3, dgjumps=True), sp.L2(c, order=2, dgjumps=True)], dgjumps=True)
        assert (c.ne, X.blocks[0].ndof, X.blocks[1].ndof, X.ndof) == (726, 10356, 4356, 14712)
        p = read_vol(REF + '/examples/Poisson/unit_square_coarse.vol')
        for _ in range(5):
            p.Refine()
        assert p.ne == 6144 and sp.H1(p, order=3).ndof == 28033


def test_dirichlet_mask_and_pattern():
    m = delaunay_rectangle(6, 2)
    X = sp.FESpace([sp.VectorH1(m, order=2, dirichlet='left|top'), sp.H1(m, order=1)])
    free = X.FreeDofs()
    assert free.dtype == bool and free.sum() < X.ndof
    # every vertex on the left boundary is constrained in both velocity components, never in the pressure
    left = np.nonzero(np.abs(m.points[:, 0]) < 1e-12)[0]
    assert not free[left].any() and not free[X.blocks[0].ndof + left].any()
    assert free[2 * X.blocks[0].ndof:].all()
    pat = X.pattern()
    import scipy.sparse as sps
    A = sps.csr_matrix((np.ones(pat.nnz), pat.colidx, pat.rowptr), shape=(pat.n, pat.n))
    assert (A - A.T).nnz == 0                                     # structurally symmetric
    assert np.all(np.diff(pat.rowptr) > 0)
    assert np.all(pat.colidx[pat.diag] == np.arange(pat.n))
    # scatter map addresses the right (row, col)
    cd = X.cell_dofs
    e, i, j = 3, 4, 11
    pos = pat.cell2nnz[e, i * X.nloc + j]
    assert pat.colidx[pos] == cd[e, j] and pat.rowptr[cd[e, i]] <= pos < pat.rowptr[cd[e, i] + 1]


def test_dgjumps_pattern_couples_facet_neighbours():
    m = structured_2d([3, 3])
    X = sp.FESpace([sp.L2(m, order=1, dgjumps=True)], dgjumps=True)
    pat = X.pattern()
    nb = np.zeros(m.ne, int)
    for c0, c1 in m.facet_cells[m.interior_facets]:
        nb[c0] += 1
        nb[c1] += 1
    assert pat.nnz == (3 * 3 * (1 + nb)).sum()
    assert pat.facet2nnz.shape == (len(m.interior_facets), 2, 9)


def test_hdiv_normal_continuity():
    """Contravariant Piola with signed det(J) and vertex-sorted cells gives single-valued normal fluxes."""
    m = delaunay_rectangle(5, 4)
    V = sp.HDiv(m, order=2)
    b = V.blocks[0].basis
    fp, fw = facet_rule_in_cell('tri', 4)
    tabs = b.tabulate_facets(4)
    J = m.jacobians()
    det = np.linalg.det(J)
    rng = np.random.default_rng(3)
    coef = rng.uniform(-1, 1, V.ndof)
    for f_ in m.interior_facets[:20]:
        vals = []
        t = m.points[m.facets[f_, 1]] - m.points[m.facets[f_, 0]]
        nrm = np.array([t[1], -t[0]])
        for s in (0, 1):
            c, lf = m.facet_cells[f_, s], m.facet_local[f_, s]
            u_ref = tabs[lf][:, :2, :] @ coef[V.cell_dofs[c]]
            u = (J[c] @ u_ref.T).T / det[c]
            vals.append(u @ nrm)
        assert np.abs(vals[0] - vals[1]).max() < 1e-11 * max(1.0, np.abs(vals[0]).max())


def test_symbolic_lowering_entries_and_bytecode(oracle_backend):
    ngs = oracle_backend
    m = ngs.Mesh(structured_2d([2, 2]))
    X = ngs.FESpace([ngs.VectorH1(m, order=2), ngs.H1(m, order=1)])
    (u, p), (v, q) = X.TrialFunction(), X.TestFunction()
    dt = ngs.Parameter(0.5)
    a = ngs.BilinearForm(X)
    a += dt * (2.0 * ngs.InnerProduct(ngs.Grad(u), ngs.Grad(v)) - ngs.div(u) * q - ngs.div(v) * p) * ngs.dx
    prog = a.program()
    assert len(prog.integrals) == 1
    ent = prog.integrals[0].entries
    # grad-grad: 4 entries, div-q: 2, div-p: 2
    assert ent.shape == (8, 3)
    rows = set(map(tuple, ent[:, :2]))
    assert (1, 1) in rows and (2, 2) in rows and (4, 4) in rows and (5, 5) in rows
    assert (6, 1) in rows and (6, 5) in rows and (1, 6) in rows and (5, 6) in rows
    assert str(u).find('trial-function') >= 0 and str(v).find('trial-function') < 0
    with pytest.raises(TypeError):
        _ = (u * u) * v                                     # two trial functions in one integrand
    # IfPos / Norm / Max-Min style trees evaluate like NumPy (reference helpers/math.py:134-221)
    cf = ngs.IfPos(ngs.x - 0.5, ngs.sin(ngs.x) * ngs.y, ngs.Norm(ngs.CoefficientFunction((ngs.x, ngs.y)))) * dt
    val = ngs.Integrate(cf, m, order=6)
    pts, w = cell_rule('tri', 6)
    tot = 0.0
    for c in range(m.ne):
        P = m.points[m.cells[c, 0]] + pts @ m.jacobians()[c].T
        g = np.where(P[:, 0] - 0.5 > 0, np.sin(P[:, 0]) * P[:, 1], np.hypot(P[:, 0], P[:, 1])) * 0.5
        tot += abs(np.linalg.det(m.jacobians()[c])) * (w * g).sum()
    assert abs(val - tot) < 1e-13


def test_gridfunction_set_projects_boundary_data(oracle_backend):
    ngs = oracle_backend
    m = ngs.Mesh(delaunay_rectangle(6, 5))
    X = ngs.FESpace([ngs.H1(m, order=3, dirichlet='left|bottom')])
    g = ngs.GridFunction(X)
    cf = 1.0 + ngs.x * ngs.x * ngs.y + ngs.y ** 3          # cubic: reproduced exactly on the boundary
    g.components[0].Set(cf, definedon=m.Boundaries('left|bottom'))
    mask = ~X.FreeDofs()
    assert np.abs(g.vec.NumPy()[~mask]).max() == 0.0
    full = ngs.GridFunction(X)
    full.components[0].Set(cf)
    assert np.abs(full.vec.NumPy()[mask] - g.vec.NumPy()[mask]).max() < 1e-12
    err = ngs.Integrate((full.components[0] - cf) ** 2, m)
    assert err < 1e-24


def test_pattern_keys_sorted_in_slices_equal_the_plain_result():
    """space._unique_and_locate in slices of the key range (used above 2^30 keys, e.g. the 3-D N = 64 pattern with
    2.08e9 keys) gives the same unique keys and the same positions as the one-shot NumPy / torch paths."""
    import torch
    from opencmp_b200.space import _unique_and_locate
    rng = np.random.default_rng(5)
    keys = [rng.integers(0, 5000, 20000).astype(np.int64) * 7 + 3, rng.integers(100, 900, 5000).astype(np.int64) * 7 + 3]
    queries = keys + [np.unique(keys[0])[::3].copy()]
    ref_u, ref_m = _unique_and_locate(keys, queries)                       # NumPy path (small, no GPU here)
    for nslices_keys in (4096, 1000, 25000):
        u, m = _unique_and_locate(keys, queries, device=torch.device('cpu'), slice_keys=nslices_keys)
        assert np.array_equal(u, ref_u)
        for a, b in zip(m, ref_m):
            assert np.array_equal(a, b)
        for q, a in zip(queries, m):
            assert np.array_equal(u[a], q)                                  # every query sits at its position
    one, mone = _unique_and_locate(keys, queries, device=torch.device('cpu'), slice_keys=1 << 40)   # torch, one shot
    assert np.array_equal(one, ref_u) and all(np.array_equal(a, b) for a, b in zip(mone, ref_m))


def test_integrate_over_element_boundaries(oracle_backend):
    """``Integrate(f * dx(element_boundary=True), mesh)`` — the facet-jump metric of the reference
    (helpers/error.py:131-148, switched on in examples/INS/ref_sol_dir/ref_sol_config): every interior facet is seen
    from both cells, every boundary facet once; a continuous field has no jumps; the outward normal turns with the
    cell (divergence theorem cell by cell)."""
    ngs = oracle_backend
    from opencmp_b200.mesh import structured_2d
    m = ngs.Mesh(structured_2d([3, 2]))
    one = ngs.CoefficientFunction(1.0)
    skel = ngs.Integrate(one * ngs.dx(skeleton=True), m)
    assert abs(ngs.Integrate(one * ngs.dx(element_boundary=True), m) - (2 * skel + 4.0)) < 1e-12
    g = ngs.GridFunction(ngs.H1(m, order=2))
    g.Set(ngs.x * ngs.y + ngs.sin(ngs.x))
    assert ngs.Integrate((g - g.Other()) ** 2 * ngs.dx(element_boundary=True), m) < 1e-24
    L = ngs.L2(m, order=0)
    q = ngs.GridFunction(L)
    q.vec.data = ngs.BaseVector(ngs.get_backend().from_numpy(np.arange(L.ndof, dtype=float)))
    jumps = ngs.Integrate((q - q.Other()) ** 2 * ngs.dx(element_boundary=True), m)
    assert abs(jumps - 2 * ngs.Integrate((q - q.Other()) ** 2 * ngs.dx(skeleton=True), m)) < 1e-10 and jumps > 1.0
    w = ngs.GridFunction(ngs.VectorH1(m, order=1))
    w.Set(ngs.CoefficientFunction((ngs.x, ngs.y)))
    n = ngs.specialcf.normal(2)
    assert abs(ngs.Integrate((w * n) * ngs.dx(element_boundary=True), m) - 2.0) < 1e-12
    u, v = ngs.H1(m, order=1).TnT()
    with pytest.raises(NotImplementedError):
        a = ngs.BilinearForm(ngs.H1(m, order=1))
        a += u * v * ngs.dx(element_boundary=True)
        a.Assemble()
