"""``netgen.meshing`` builder subset + ``edt`` stand-in, so that the reference's diffuse-interface pre-processing
(``opencmp/diffuse_interface/{dim,interface,mesh_helpers}.py``) runs unmodified (host side).

* ``Mesh()`` with ``.dim``, ``Add(MeshPoint(Pnt(x,y,z)))``, ``AddRegion(name, dim=)``, ``Add(Element1D/2D/3D)``,
  ``Add(FaceDescriptor(...))``, ``SetBCName``, ``Compress``, ``Load``, ``Save`` — exactly what
  ``mesh_helpers.get_Netgen_nonconformal`` (:494-690) calls to build its structured quad / triangle / hex meshes.
  ``ngs.Mesh(ngmesh)`` converts the collected arrays into an opencmp_b200 mesh.
* ``edt.edt(array)``: Euclidean distance of every non-zero voxel to the nearest zero voxel (the third-party ``edt``
  package the reference calls at ``interface.py:167``), here through ``scipy.ndimage.distance_transform_edt``.
"""
from __future__ import annotations

from typing import List

import numpy as np


class Pnt:
    def __init__(self, x=0.0, y=0.0, z=0.0):
        self.p = (float(x), float(y), float(z))


class MeshPoint:
    def __init__(self, pnt: Pnt):
        self.p = pnt.p


class PointId(int):
    """1-based like Netgen's PointId; usable as a list element in Element*D(...)."""
    @property
    def nr(self):
        return int(self)


class Element1D:
    def __init__(self, vertices, index=1, **_):
        self.vertices, self.index = [int(v) for v in vertices], int(index)


class Element2D:
    def __init__(self, index=1, vertices=(), **_):
        self.vertices, self.index = [int(v) for v in vertices], int(index)


class Element3D:
    def __init__(self, index=1, vertices=(), **_):
        self.vertices, self.index = [int(v) for v in vertices], int(index)


class FaceDescriptor:
    def __init__(self, surfnr=0, domin=1, domout=0, bc=1, **_):
        self.surfnr, self.domin, self.domout, self.bc = surfnr, domin, domout, bc


class Mesh:
    def __init__(self, dim: int = 3):
        self.dim = dim
        self._points: List[tuple] = []
        self._e1: List[Element1D] = []
        self._e2: List[Element2D] = []
        self._e3: List[Element3D] = []
        self._fd: List[FaceDescriptor] = []
        self._regions = {1: [], 2: [], 3: []}       # names per co-dimension-independent "dim" of AddRegion
        self._bcnames = {}
        self._built = None

    # ---- builder -----------------------------------------------------------------------------------------------
    def Add(self, obj):
        self._built = None
        if isinstance(obj, MeshPoint):
            self._points.append(obj.p)
            return PointId(len(self._points))
        if isinstance(obj, Element1D):
            self._e1.append(obj)
            return len(self._e1)
        if isinstance(obj, Element2D):
            self._e2.append(obj)
            return len(self._e2)
        if isinstance(obj, Element3D):
            self._e3.append(obj)
            return len(self._e3)
        if isinstance(obj, FaceDescriptor):
            self._fd.append(obj)
            return len(self._fd)
        raise TypeError('netgen.meshing.Mesh.Add: unsupported object {}'.format(type(obj).__name__))

    def AddRegion(self, name: str, dim: int) -> int:
        self._regions[int(dim)].append(str(name))
        return len(self._regions[int(dim)])

    def SetBCName(self, index: int, name: str) -> None:
        self._bcnames[int(index)] = str(name)

    def Compress(self) -> None:
        return None

    def Load(self, filename: str) -> None:
        from .mesh import load_mesh
        self._built = load_mesh(filename)
        self.dim = self._built.dim

    def Save(self, filename: str) -> None:
        m = self._mesh
        np.savez_compressed(filename if str(filename).endswith('.npz') else str(filename) + '.npz', points=m.points,
                            cells=m.cells, bnd=m.facets[m.bnd_facets], bnd_region=m.bnd_region,
                            bnd_names=np.array(m.bnd_names))

    # ---- conversion ----------------------------------------------------------------------------------------------
    @property
    def _mesh(self):
        if self._built is None:
            self._built = self._convert()
        return self._built

    def _convert(self):
        from .mesh import Mesh as _Mesh, _orient_quads
        P = np.array(self._points, dtype=np.float64).reshape(-1, 3)
        if self.dim == 2:
            nv = {len(e.vertices) for e in self._e2}
            if nv not in ({3}, {4}):
                raise ValueError('netgen.meshing shim: 2-D meshes of triangles or of quadrilaterals only')
            cells = np.array([e.vertices for e in self._e2], dtype=np.int64) - 1
            mat = np.array([e.index for e in self._e2], dtype=np.int32) - 1
            bnd = np.array([e.vertices for e in self._e1], dtype=np.int64).reshape(-1, 2) - 1
            bidx = np.array([e.index for e in self._e1], dtype=np.int32) - 1
            names = list(self._regions[1]) or ['default']
            mats = list(self._regions[2]) or ['default']
            if nv == {4}:
                cells = cells[:, [0, 1, 3, 2]]                      # counter-clockwise -> lattice order
                m = _Mesh(2, 'quad', P, cells, bnd, bidx, names, mat, mats)
                _orient_quads(m)
                return m
            return _Mesh(2, 'tri', P, cells, bnd, bidx, names, mat, mats)
        nv = {len(e.vertices) for e in self._e3}
        if nv != {8}:
            raise NotImplementedError('netgen.meshing shim: 3-D meshes of hexahedra only (the reference\'s non-quad 3-D '
                                      'branch builds 5-vertex pyramids, which no finite-element space in scope supports)')
        c = np.array([e.vertices for e in self._e3], dtype=np.int64) - 1
        # re-order every cell so that local vertex l has bits (x, y, z) = (l & 1, l >> 1 & 1, l >> 2): the structured
        # generator only builds axis-aligned boxes, so the position inside the cell's bounding box decides the slot
        Pc = P[c]
        lo = Pc.min(axis=1, keepdims=True)
        bits = (Pc > lo + 1e-12 * (1.0 + np.abs(lo))).astype(np.int64)
        slot = bits[:, :, 0] + 2 * bits[:, :, 1] + 4 * bits[:, :, 2]
        cells = np.empty_like(c)
        if not (np.sort(slot, axis=1) == np.arange(8)[None, :]).all():
            raise ValueError('netgen.meshing shim: hexahedra must be axis-aligned boxes')
        np.put_along_axis(cells, slot, c, axis=1)
        quads = np.array([e.vertices for e in self._e2], dtype=np.int64).reshape(-1, 4) - 1
        qidx = np.array([e.index for e in self._e2], dtype=np.int32) - 1
        nb = max(self._bcnames) + 1 if self._bcnames else (int(qidx.max()) + 1 if len(qidx) else 0)
        names = [self._bcnames.get(i, 'default') for i in range(nb)]
        return _Mesh(3, 'hex', P, cells, quads, qidx, names)


def ReadGmsh(filename: str):
    from .mesh import read_msh
    out = Mesh()
    out._built = read_msh(filename if str(filename).endswith('.msh') else str(filename) + '.msh')
    out.dim = out._built.dim
    return out


# ---- edt ------------------------------------------------------------------------------------------------------------
def edt(data, anisotropy=None, black_border=False, **_):
    """Exact Euclidean distance transform: distance of each non-zero voxel to the nearest zero voxel. With the CUDA
    backend active the transform runs on the device (dimgen.edt, csrc/ocmp_dim.cu); CPU runs (oracle backend) use
    SciPy."""
    from . import ngs as _ngs
    if getattr(_ngs._backend, 'name', '') == 'cuda' and anisotropy is None and not black_border:
        from . import dimgen
        return dimgen.edt(data)
    import scipy.ndimage as ndi
    a = np.asarray(data) != 0
    if black_border:
        a = np.pad(a, 1, constant_values=False)
    out = ndi.distance_transform_edt(a, sampling=anisotropy)
    if black_border:
        out = out[tuple(slice(1, -1) for _ in range(out.ndim))]
    return out.astype(np.float32 if np.asarray(data).dtype == np.float32 else np.float64)
