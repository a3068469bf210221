"""Host-side pieces of the diffuse-interface pre-processing path: VoxelCoefficient, Set from data
living on another mesh, the netgen.meshing builder shim and the edt stand-in."""
import numpy as np
import pytest


@pytest.fixture()
def ngs():
    import opencmp_b200.ngs as ngs
    from oracle.backend import OracleBackend
    old = ngs._backend
    ngs.set_backend(OracleBackend())
    yield ngs
    ngs.set_backend(old)


def test_voxel_coefficient_on_matching_quad_grid_is_nodal(ngs):
    from opencmp_b200.mesh import structured_2d
    m = ngs.Mesh(structured_2d([6, 4], scale=(3.0, 2.0), offset=(1.5, 1.0), cell='quad'))
    xs, ys = np.linspace(-1.5, 1.5, 7), np.linspace(-1.0, 1.0, 5)
    data = np.sin(xs)[None, :] * (1.0 + ys ** 2)[:, None]                      # indexed [y][x]
    g = ngs.GridFunction(ngs.H1(m, order=3))
    g.Set(ngs.VoxelCoefficient((-1.5, -1.0), (1.5, 1.0), data, linear=True))
    v = g.vec.NumPy()
    assert np.abs(v[m.nv:]).max() == 0.0                                        # multilinear part only
    assert np.abs(v[:m.nv] - np.sin(m.points[:, 0]) * (1.0 + m.points[:, 1] ** 2)).max() < 1e-14
    # the bilinear interpolant is reproduced inside the cells
    val = g(m(0.3, 0.2))
    i, j = np.searchsorted(xs, 0.3) - 1, np.searchsorted(ys, 0.2) - 1
    tx, ty = (0.3 - xs[i]) / (xs[i + 1] - xs[i]), (0.2 - ys[j]) / (ys[j + 1] - ys[j])
    ref = (data[j, i] * (1 - tx) * (1 - ty) + data[j, i + 1] * tx * (1 - ty) + data[j + 1, i] * (1 - tx) * ty
           + data[j + 1, i + 1] * tx * ty)
    assert abs(val - ref) < 1e-13


def test_voxel_coefficient_on_other_grids_is_projected(ngs):
    """Non-matching grid (and a triangle mesh): local L2 projection of the interpolated data. Linear data are
    reproduced exactly by either route."""
    from opencmp_b200.mesh import structured_2d
    xs, ys = np.linspace(0.0, 1.0, 12), np.linspace(0.0, 1.0, 9)
    data = 2.0 * xs[None, :] - 3.0 * ys[:, None] + 0.5
    for cell in ('quad', 'tri'):
        m = ngs.Mesh(structured_2d([5, 5], cell=cell) if cell == 'quad' else structured_2d([5, 5]))
        g = ngs.GridFunction(ngs.H1(m, order=2))
        g.Set(ngs.VoxelCoefficient((0.0, 0.0), (1.0, 1.0), data, linear=True))
        err = ngs.Integrate((g - (2.0 * ngs.x - 3.0 * ngs.y + 0.5)) ** 2, m)
        assert err < 1e-24


def test_set_from_gridfunction_on_finer_mesh(ngs):
    """dim.py:393-409: phi, grad(phi) and |grad(phi)| of a fine-mesh GridFunction projected onto the simulation mesh.
    For a polynomial that both spaces contain the projection is exact; it equals the device-path Set of the same
    expression given analytically."""
    from opencmp_b200.mesh import structured_2d
    fine = ngs.Mesh(structured_2d([12, 12], scale=(2.0, 2.0), offset=(1.0, 1.0), cell='quad'))
    coarse = ngs.Mesh(structured_2d([4, 4], scale=(2.0, 2.0), offset=(1.0, 1.0), cell='quad'))
    x, y = ngs.x, ngs.y
    f = x * x * y - 0.5 * y * y + x
    gfine = ngs.GridFunction(ngs.H1(fine, order=2))
    gfine.Set(f)
    phi = ngs.GridFunction(ngs.H1(coarse, order=2))
    gphi = ngs.GridFunction(ngs.VectorH1(coarse, order=2))
    mag = ngs.GridFunction(ngs.H1(coarse, order=2))
    phi.Set(gfine)
    gphi.Set(ngs.Grad(gfine))
    mag.Set(ngs.Norm(ngs.Grad(gfine)))
    ref = ngs.GridFunction(ngs.H1(coarse, order=2))
    ref.Set(f)
    assert np.abs(phi.vec.NumPy() - ref.vec.NumPy()).max() < 1e-12
    rg = ngs.GridFunction(ngs.VectorH1(coarse, order=2))
    rg.Set(ngs.CoefficientFunction((2.0 * x * y + 1.0, x * x - y)))
    assert np.abs(gphi.vec.NumPy() - rg.vec.NumPy()).max() < 1e-11
    rm = ngs.GridFunction(ngs.H1(coarse, order=2))
    rm.Set(ngs.sqrt((2.0 * x * y + 1.0) ** 2 + (x * x - y) ** 2))
    assert np.abs(mag.vec.NumPy() - rm.vec.NumPy()).max() < 1e-10


def test_netgen_builder_shim_structured_meshes(ngs):
    """The call sequence of mesh_helpers.get_Netgen_nonconformal (:494-690), 2-D quads and 3-D hexes."""
    from opencmp_b200 import netgen_shim as ngm
    N = (3, 2)
    msh = ngm.Mesh()
    msh.dim = 2
    pts = []
    for i in range(N[1] + 1):
        for j in range(N[0] + 1):
            pts.append(msh.Add(ngm.MeshPoint(ngm.Pnt(j / N[0], i / N[1], 0.0))))
    dom = msh.AddRegion('dom', dim=2)
    b, r, t, l = (msh.AddRegion(n, dim=1) for n in ('bottom', 'right', 'top', 'left'))
    for i in range(N[1]):
        for j in range(N[0]):
            p1 = i * (N[0] + 1) + j
            msh.Add(ngm.Element2D(dom, [pts[p1], pts[p1 + 1], pts[p1 + 2 + N[0]], pts[p1 + 1 + N[0]]]))
    for i in range(N[1]):
        msh.Add(ngm.Element1D([pts[N[0] + i * (N[0] + 1)], pts[N[0] + (i + 1) * (N[0] + 1)]], index=r))
        msh.Add(ngm.Element1D([pts[(i + 1) * (N[0] + 1)], pts[i * (N[0] + 1)]], index=l))
    for i in range(N[0]):
        msh.Add(ngm.Element1D([pts[i], pts[i + 1]], index=b))
        msh.Add(ngm.Element1D([pts[1 + i + N[1] * (N[0] + 1)], pts[i + N[1] * (N[0] + 1)]], index=t))
    msh.Compress()
    m = ngs.Mesh(msh)
    assert m.cell_type == 'quad' and m.ne == 6 and m.nv == 12
    assert m.GetBoundaries() == ('bottom', 'right', 'top', 'left') and m.GetMaterials() == ('dom',)
    m.check_affine()
    assert abs(ngs.Integrate(ngs.CoefficientFunction(1.0), m) - 1.0) < 1e-14
    for name, count in (('bottom', 3), ('top', 3), ('left', 2), ('right', 2)):
        assert (m.bnd_region == m.bnd_names.index(name)).sum() == count


def test_edt_stand_in():
    from opencmp_b200.netgen_shim import edt
    a = np.ones((7, 9), dtype=np.float32)
    a[3, 4] = 0.0
    d = edt(a)
    jj, ii = np.meshgrid(np.arange(7), np.arange(9), indexing='ij')
    assert d.dtype == np.float32 and np.abs(d - np.hypot(jj - 3, ii - 4)).max() < 1e-6
