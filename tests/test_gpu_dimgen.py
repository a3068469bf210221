"""GPU: diffuse-interface phase-field generation on the device (opencmp_b200/dimgen.py, csrc/ocmp_dim.cu; SURVEY 8(f)
N4) against a restatement of the reference's host pipeline — reference opencmp/diffuse_interface/interface.py:31-57
(get_binary_2d with mesh_helpers.ray_trace_2d, :268-302) and :137-180 (get_phi: scipy erosion, exact EDT, erf).
Masks and squared distances are compared bit-exactly (integer work), the phase field at FP32 round-off (the reference
computes the distance and the erf in FP32)."""
import numpy as np
import pytest

from test_gpu_parity import _with

pytestmark = pytest.mark.gpu


def ray_trace_2d(x, y, polygon):
    """mesh_helpers.py:268-302, restated (the carried ``xints`` on horizontal edges included)."""
    n = len(polygon)
    inside = False
    xints = 0.0
    p1x, p1y = polygon[0]
    for i in range(n + 1):
        p2x, p2y = polygon[i % n]
        if y > min(p1y, p2y):
            if y <= max(p1y, p2y):
                if x <= max(p1x, p2x):
                    if p1y != p2y:
                        xints = (y - p1y) * (p2x - p1x) / (p2y - p1y) + p1x
                    if p1x == p2x or x <= xints:
                        inside = not inside
        p1x, p1y = p2x, p2y
    return inside


def host_binary_2d(boundary, N, scale, offset):
    shape = (int(N[0] + 1), int(N[1] + 1))
    out = np.empty(shape)
    for i in range(shape[0]):
        for j in range(shape[1]):
            out[i, j] = ray_trace_2d(i * scale[0] / N[0] - offset[0], j * scale[1] / N[1] - offset[1], boundary)
    return out


def host_phi(binary, lmbda, N, scale, dim):
    import scipy.ndimage as spimg
    import scipy.special as spec
    kernel = np.ones((3,) * dim)
    erosion = spimg.binary_erosion(binary, kernel, 1).astype(np.float32)
    border = (1.0 - (binary - erosion)).astype(np.float32)
    dt = spimg.distance_transform_edt(border != 0).astype(np.float32)
    dt *= min(scale) / min(N)
    e = spec.erf(dt / lmbda)
    return (e * binary + e * (binary - 1.0) + 1.0) / 2.0, border


def _star(k=7, r0=0.9, r1=0.45):
    ang = np.linspace(0.0, 2.0 * np.pi, 2 * k, endpoint=False)
    rad = np.where(np.arange(2 * k) % 2 == 0, r0, r1)
    return [(float(r * np.cos(a)), float(r * np.sin(a))) for r, a in zip(rad, ang)]


@pytest.mark.parametrize('poly', [_star(), [(-0.5, -0.5), (0.5, -0.5), (0.5, 0.5), (-0.5, 0.5)],
                                  [(-1.0, 0.0), (0.0, -0.25), (1.0, 0.0), (0.0, 0.25)]])
def test_ray_traced_mask_is_identical(poly):
    from opencmp_b200 import dimgen
    N, scale, offset = [48, 40], [4.0, 3.0], [2.0, 1.5]
    ref = host_binary_2d(poly, N, scale, offset)
    got = _with('cuda', lambda: dimgen.get_binary_2d(poly, N, scale, offset))
    assert got.shape == ref.shape and np.array_equal(got, ref)
    assert 0 < ref.sum() < ref.size


@pytest.mark.parametrize('shape', [(65, 49), (33, 29, 25)])
def test_distance_transform_is_exact(shape):
    """Squared distances are integers: the FP32 distances must be the correctly rounded roots of the exact ones."""
    import scipy.ndimage as spimg
    from opencmp_b200 import dimgen
    rng = np.random.default_rng(len(shape))
    a = (rng.uniform(size=shape) > 0.02).astype(np.float32)          # sparse zeros = seeds
    ref = spimg.distance_transform_edt(a != 0)
    got = _with('cuda', lambda: dimgen.edt(a))
    assert got.dtype == np.float32 and got.shape == a.shape
    d2_ref = np.rint(ref ** 2).astype(np.int64)
    assert np.array_equal(got, np.sqrt(d2_ref.astype(np.float32)))
    assert (got[a == 0] == 0).all() and d2_ref.max() > 4


@pytest.mark.parametrize('dim', [2, 3])
def test_phase_field_matches_the_host_pipeline(dim):
    from opencmp_b200 import dimgen
    if dim == 2:
        N, scale, offset = [64, 64], [4.0, 4.0], [2.0, 2.0]
        binary = host_binary_2d(_star(), N, scale, offset)
    else:
        N, scale, offset = [32, 28, 24], [2.0, 2.0, 2.0], [1.0, 1.0, 1.0]
        g = np.meshgrid(*[np.linspace(-1.0, 1.0, n + 1) for n in N], indexing='ij')
        binary = ((g[0] ** 2 + g[1] ** 2 + g[2] ** 2) < 0.55 ** 2).astype(np.float64)
    lmbda = 0.08
    ref, _ = host_phi(binary, lmbda, N, scale, dim)
    got = _with('cuda', lambda: dimgen.get_phi(binary, lmbda, N, scale, offset, dim))
    assert got.shape == ref.shape and got.dtype == np.float64
    assert np.abs(got - ref).max() < 5e-7                       # FP32 distance and erf on both sides
    assert got.min() >= 0.0 and got.max() <= 1.0 and got[binary == 1].min() >= 0.5 - 1e-7


def host_rigid_motion(values, inv_R, N, scale, offset):
    """helpers/ngsolve_.py:236-262 restated for a multilinear node field: per node the pre-image under the rotation,
    a bounds test, and the multilinear interpolant of the node values there (what ``orig_gfu(mesh(x, y))`` evaluates)."""
    dim = values.ndim
    out = np.ones_like(values)
    N = np.asarray(N[:dim], dtype=np.float64)
    sc, off = np.asarray(scale[:dim], dtype=np.float64), np.asarray(offset[:dim], dtype=np.float64)
    for idx in np.ndindex(*values.shape):
        xn = -off + sc * np.asarray(idx) / N
        xo = np.asarray(inv_R).reshape(dim, dim) @ xn
        if not ((xo >= -off).all() and (xo <= sc - off).all()):
            continue
        u = (xo + off) / sc * N
        i0 = np.clip(np.floor(u).astype(int), 0, N.astype(int) - 1)
        f = u - i0
        val = 0.0
        for corner in range(1 << dim):
            bits = [(corner >> a) & 1 for a in range(dim)]
            w = np.prod([f[a] if bits[a] else 1.0 - f[a] for a in range(dim)])
            val += w * values[tuple(i0 + np.asarray(bits))]
        out[idx] = val
    return out


@pytest.mark.parametrize('dim', [2, 3])
def test_rigid_body_motion_matches_the_host_loop(dim):
    from opencmp_b200 import dimgen
    rng = np.random.default_rng(dim)
    if dim == 2:
        N, scale, offset = [20, 16], [4.0, 3.0], [2.0, 1.5]
        th = 0.37
        inv_R = np.array([[np.cos(th), np.sin(th)], [-np.sin(th), np.cos(th)]])
    else:
        N, scale, offset = [10, 8, 6], [2.0, 2.0, 2.0], [1.0, 1.0, 1.0]
        th = 0.5
        inv_R = np.array([[np.cos(th), np.sin(th), 0.0], [-np.sin(th), np.cos(th), 0.0], [0.0, 0.0, 1.0]])
    values = rng.uniform(0.0, 1.0, tuple(n + 1 for n in N))
    ref = host_rigid_motion(values, inv_R, N, scale, offset)
    got = _with('cuda', lambda: dimgen.rigid_body_motion(values, inv_R, N, scale, offset))
    assert np.abs(got - ref).max() < 1e-12
    assert (ref == 1.0).any() and (ref != 1.0).any()            # some pre-images leave the box, most do not
    ident = _with('cuda', lambda: dimgen.rigid_body_motion(values, np.eye(dim), N, scale, offset))
    assert np.abs(ident - values).max() < 1e-13


def test_gridfunction_rigid_body_motion_drop_in():
    """Same call as reference helpers/ngsolve_.py:212 on a quadrilateral mesh whose vertices are the N-grid nodes."""
    from opencmp_b200 import dimgen
    from opencmp_b200.mesh import structured_2d
    N, scale, offset = [12, 10], [4.0, 3.0], [2.0, 1.5]
    values = np.random.default_rng(5).uniform(0.0, 1.0, (N[0] + 1, N[1] + 1))
    th = 0.3
    inv_R = lambda t: np.array([[np.cos(th * t), np.sin(th * t)], [-np.sin(th * t), np.cos(th * t)]])

    def run():
        import opencmp_b200.ngs as ngs
        m = ngs.Mesh(structured_2d(N, scale=tuple(scale), offset=tuple(offset), cell='quad'))
        fes = ngs.H1(m, order=2)
        lo, hi = (-offset[0], -offset[1]), (scale[0] - offset[0], scale[1] - offset[1])
        orig, gfu = ngs.GridFunction(fes), ngs.GridFunction(fes)
        orig.Set(ngs.VoxelCoefficient(lo, hi, values.transpose(), linear=True))
        t = ngs.Parameter(1.5)
        out = dimgen.gridfunction_rigid_body_motion(t, orig, gfu, inv_R, m, N, scale, offset)
        assert out is gfu
        pts = m.mesh.points if hasattr(m, 'mesh') else m.points
        return ngs.get_backend().to_numpy(gfu.vec.a).copy(), np.asarray(pts).copy(), m.nv
    got, pts, nv = _with('cuda', run)
    ref = host_rigid_motion(values, inv_R(1.5), N, scale, offset)
    i = np.rint((pts[:, 0] + offset[0]) / scale[0] * N[0]).astype(int)
    j = np.rint((pts[:, 1] + offset[1]) / scale[1] * N[1]).astype(int)
    assert np.abs(got[:nv] - ref[i, j]).max() < 1e-12
    assert np.abs(got[nv:]).max() == 0.0                       # multilinear data: no high-order part
