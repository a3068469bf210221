"""VTKOutput (.vtu export used by the reference's post-processing, opencmp/post_processing/output_conversions.py:253):
files parse as VTK XML, geometry tiles the mesh with positively oriented cells, and point data equal the fields."""
import xml.etree.ElementTree as ET

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


def _read(path):
    root = ET.parse(path).getroot()
    piece = root.find('UnstructuredGrid/Piece')
    arr = lambda el: np.array(el.text.split(), dtype=np.float64)
    pts = arr(piece.find('Points/DataArray')).reshape(-1, 3)
    cells = {d.get('Name'): arr(d).astype(np.int64) for d in piece.findall('Cells/DataArray')}
    data = {d.get('Name'): arr(d).reshape(-1, int(d.get('NumberOfComponents'))) for d in piece.findall('PointData/DataArray')}
    assert int(piece.get('NumberOfPoints')) == len(pts) and int(piece.get('NumberOfCells')) == len(cells['offsets'])
    return pts, cells, data


@pytest.mark.parametrize('subdivision', [0, 2])
def test_vtu_triangles_hdiv_and_scalar(ngs, tmp_path, subdivision):
    from opencmp_b200.mesh import delaunay_rectangle
    m = ngs.Mesh(delaunay_rectangle(5, seed=2))
    fes = ngs.FESpace([ngs.HDiv(m, order=2), ngs.L2(m, order=1)])
    g = ngs.GridFunction(fes)
    x, y = ngs.x, ngs.y
    uex = ngs.CoefficientFunction((x * y - 1.0, 0.5 * x * x + y))
    g.components[0].Set(uex)
    g.components[1].Set(1.0 + 2.0 * x - y)
    out = ngs.VTKOutput(ma=m, coefs=[g.components[0], g.components[1], ngs.sin(x) * y], names=['u', 'p', 'f'],
                        filename=str(tmp_path / 'sol'), subdivision=subdivision)
    path = out.Do()
    pts, cells, data = _read(path)
    n = 2 ** subdivision
    assert len(cells['offsets']) == m.ne * n * n and set(cells['types']) == {5}
    X, Y = pts[:, 0], pts[:, 1]
    assert np.abs(data['u'][:, 0] - (X * Y - 1.0)).max() < 1e-11
    assert np.abs(data['u'][:, 1] - (0.5 * X * X + Y)).max() < 1e-11
    assert np.abs(data['u'][:, 2]).max() == 0.0                         # vectors are padded to three components
    assert np.abs(data['p'][:, 0] - (1.0 + 2.0 * X - Y)).max() < 1e-11
    assert np.abs(data['f'][:, 0] - np.sin(X) * Y).max() < 1e-14
    tri = cells['connectivity'].reshape(-1, 3)
    a, b, c = pts[tri[:, 0]], pts[tri[:, 1]], pts[tri[:, 2]]
    area = 0.5 * ((b[:, 0] - a[:, 0]) * (c[:, 1] - a[:, 1]) - (b[:, 1] - a[:, 1]) * (c[:, 0] - a[:, 0]))
    assert (area > 0).all() and abs(area.sum() - 1.0) < 1e-12


def test_vtu_hexes_taylor_hood(ngs, tmp_path):
    from opencmp_b200.mesh import structured_3d
    m = ngs.Mesh(structured_3d([2, 3, 2], scale=(2.0, 1.0, 1.0)))
    fes = ngs.FESpace([ngs.VectorH1(m, order=2), ngs.H1(m, order=1)])
    g = ngs.GridFunction(fes)
    x, y, z = ngs.x, ngs.y, ngs.z
    g.components[0].Set(ngs.CoefficientFunction((x * z, y * y, 1.0 - x * y)))
    g.components[1].Set(x + y * z)
    path = ngs.VTKOutput(ma=m, coefs=[c for c in g.components], names=['u', 'p'], filename=str(tmp_path / 'h'),
                         subdivision=1).Do()
    pts, cells, data = _read(path)
    assert len(cells['offsets']) == m.ne * 8 and set(cells['types']) == {12}
    X, Y, Z = pts.T
    assert np.abs(data['u'] - np.stack([X * Z, Y * Y, 1.0 - X * Y], axis=1)).max() < 1e-11
    assert np.abs(data['p'][:, 0] - (X + Y * Z)).max() < 1e-11
    hexa = cells['connectivity'].reshape(-1, 8)
    p = pts[hexa]
    vol = np.einsum('ni,ni->n', np.cross(p[:, 1] - p[:, 0], p[:, 3] - p[:, 0]), p[:, 4] - p[:, 0])
    assert (vol > 0).all() and abs(vol.sum() - 2.0) < 1e-12           # VTK_HEXAHEDRON ordering, positive volume


def test_vtu_argument_errors(ngs, tmp_path):
    from opencmp_b200.mesh import structured_2d
    m = ngs.Mesh(structured_2d([2, 2]))
    with pytest.raises(ValueError):
        ngs.VTKOutput(ma=m, coefs=[ngs.x], names=['a', 'b'], filename=str(tmp_path / 'e'))
