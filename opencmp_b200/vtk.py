"""``VTKOutput`` — .vtu export of solution fields (host side, not on the hot path).

Stands in for ``ngsolve.VTKOutput(ma=, coefs=, names=, filename=, subdivision=).Do()`` as called by the reference's
post-processing (``opencmp/post_processing/output_conversions.py:253-254``). Like NGSolve's writer, every cell is
written with its own points (fields may be discontinuous across cells) and ``subdivision = s`` cuts every cell into
2^s pieces per direction; values are evaluated cell-wise from the reference tabulations (vectorised over all cells).
Output: ASCII VTK XML UnstructuredGrid with one PointData array per name (vectors padded to three components).
"""
from __future__ import annotations

import itertools
from typing import List, Sequence

import numpy as np

from .ir import coef_leaves

_NP_UNARY = {'neg': np.negative, 'abs': np.abs, 'sqrt': np.sqrt, 'sin': np.sin, 'cos': np.cos, 'tan': np.tan,
             'exp': np.exp, 'log': np.log, 'tanh': np.tanh, 'floor': np.floor, 'ceil': np.ceil, 'round': np.round,
             'atan': np.arctan}
_NP_BINARY = {'add': np.add, 'sub': np.subtract, 'mul': np.multiply, 'div': np.divide, 'pow': np.power,
              'min': np.minimum, 'max': np.maximum}
_VTK_TYPE = {'tri': 5, 'quad': 9, 'tet': 10, 'hex': 12}
# VTK vertex order of a sub-cell given our lattice order (bits = reference axes)
_VTK_PERM = {'quad': [0, 1, 3, 2], 'hex': [0, 1, 3, 2, 4, 5, 7, 6]}


def _erf(x):
    from math import erf
    return np.vectorize(erf, otypes=[np.float64])(x)


def reference_subcells(cell_type: str, subdivision: int):
    """(points (npts, dim) in the reference cell, sub-cell connectivity into those points) for 2^s pieces per edge."""
    n = 2 ** int(subdivision)
    dim = 2 if cell_type in ('tri', 'quad') else 3
    if cell_type in ('quad', 'hex'):
        idx = list(itertools.product(range(n + 1), repeat=dim))          # (i_{d-1}, ..., i_0), i_0 fastest
        pts = np.array([t[::-1] for t in idx], dtype=np.float64) / n
        num = {t[::-1]: k for k, t in enumerate(idx)}
        cells = []
        for c in itertools.product(range(n), repeat=dim):
            c = c[::-1]
            verts = [num[tuple(c[a] + b[a] for a in range(dim))]
                     for b in (bb[::-1] for bb in itertools.product(range(2), repeat=dim))]
            cells.append([verts[k] for k in _VTK_PERM[cell_type]])
        return pts, np.array(cells, dtype=np.int64)
    if cell_type == 'tri':
        num, pts = {}, []
        for j in range(n + 1):
            for i in range(n + 1 - j):
                num[(i, j)] = len(pts)
                pts.append((i / n, j / n))
        cells = []
        for j in range(n):
            for i in range(n - j):
                cells.append([num[(i, j)], num[(i + 1, j)], num[(i, j + 1)]])
                if i + j < n - 1:
                    cells.append([num[(i + 1, j)], num[(i + 1, j + 1)], num[(i, j + 1)]])
        return np.array(pts, dtype=np.float64), np.array(cells, dtype=np.int64)
    if subdivision:
        raise NotImplementedError('VTKOutput: subdivision of tetrahedra')
    return np.array([[0., 0., 0.], [1., 0., 0.], [0., 1., 0.], [0., 0., 1.]]), np.array([[0, 1, 2, 3]], dtype=np.int64)


def evaluate_on_cells(cf, mesh, ref_pts: np.ndarray) -> np.ndarray:
    """Values of a coefficient function at the images of ``ref_pts`` in every cell: (ncomp, ne, npts)."""
    J = mesh.jacobians()                                                   # (ne, d, d)
    x = mesh.origins()[:, None, :] + np.einsum('eia,ka->eki', J, ref_pts)   # (ne, npts, d)
    Ji = np.linalg.inv(J)
    det = np.linalg.det(J)
    d = mesh.dim
    cache, fcache = {}, {}

    def field(gf, blk):
        key = (id(gf), blk)
        if key not in fcache:
            fes = gf.space
            b = fes.blocks[blk]
            tab = b.basis.tabulate(ref_pts)                                # (npts, nrows, nloc)
            lo = fes.loc_offsets[blk]
            coef = np.asarray(gf.vec_numpy())[fes.cell_dofs[:, lo:lo + b.nloc]]      # (ne, nloc)
            ref = np.einsum('krl,el->ekr', tab, coef)                      # (ne, npts, nrows)
            if b.kind == 'scalar':
                grad = np.einsum('eab,eka->ekb', Ji, ref[:, :, 1:])        # J^-T grad_ref
                phys = np.concatenate([ref[:, :, :1], grad], axis=2)
            else:
                val = np.einsum('eia,eka->eki', J, ref[:, :, :d]) / det[:, None, None]
                g = ref[:, :, d:].reshape(ref.shape[0], ref.shape[1], d, d)
                g = np.einsum('eia,ekab,ebj->ekij', J, g, Ji) / det[:, None, None, None]
                phys = np.concatenate([val, g.reshape(ref.shape[0], ref.shape[1], d * d)], axis=2)
            fcache[key] = phys
        return fcache[key]

    def ev(c):
        if id(c) in cache:
            return cache[id(c)]
        if c.op == 'const':
            v = np.full(x.shape[:2], float(c.val))
        elif c.op == 'param':
            v = np.full(x.shape[:2], float(c.val.Get()))
        elif c.op == 'coord':
            v = x[:, :, c.val] if c.val < d else np.zeros(x.shape[:2])
        elif c.op == 'field':
            gf, blk, row, _side = c.val
            v = field(gf, blk)[:, :, row]
        elif c.op == 'ifpos':
            v = np.where(ev(c.args[0]) > 0, ev(c.args[1]), ev(c.args[2]))
        elif c.op == 'erf':
            v = _erf(ev(c.args[0]))
        elif c.op in _NP_UNARY:
            v = _NP_UNARY[c.op](ev(c.args[0]))
        elif c.op in _NP_BINARY:
            v = _NP_BINARY[c.op](ev(c.args[0]), ev(c.args[1]))
        elif c.op == 'piecewise':
            kind, lst = c.val
            if kind == 'mat':
                v = np.zeros(x.shape[:2])
                for rid, sub in enumerate(lst):
                    sel = mesh.cell_mat == rid
                    if sel.any():
                        v[sel] = ev(sub)[sel]
            else:
                v = ev(lst[0])
        else:
            raise NotImplementedError('VTKOutput: evaluation of {} inside cells'.format(c.op))
        cache[id(c)] = v
        return v

    return np.stack([ev(s.as_coef()) for s in cf.arr.reshape(-1)], axis=0)


class VTKOutput:
    def __init__(self, ma=None, coefs: Sequence = (), names: Sequence[str] = (), filename: str = 'output',
                 subdivision: int = 0, legacy: bool = False, **_ignored):
        if legacy:
            raise NotImplementedError('VTKOutput: legacy .vtk format')
        if len(coefs) != len(names):
            raise ValueError('VTKOutput: {} coefficient functions but {} names'.format(len(coefs), len(names)))
        self.mesh, self.coefs, self.names = ma, list(coefs), [str(n) for n in names]
        self.filename, self.subdivision = str(filename), int(subdivision)
        self.count = 0

    def Do(self, time: float = None, **_ignored) -> str:
        from .symbolic import CoefficientFunction
        m = self.mesh
        ref_pts, sub = reference_subcells(m.cell_type, self.subdivision)
        npts = len(ref_pts)
        J = m.jacobians()
        x = m.origins()[:, None, :] + np.einsum('eia,ka->eki', J, ref_pts)
        pts = np.zeros((m.ne * npts, 3))
        pts[:, :m.dim] = x.reshape(-1, m.dim)
        conn = (sub[None, :, :] + (np.arange(m.ne) * npts)[:, None, None]).reshape(-1, sub.shape[1])
        if m.cell_type in ('tri', 'tet'):
            # simplices are stored with ascending vertex numbers and may be negatively oriented
            neg = np.repeat(np.linalg.det(J) < 0, len(sub))
            conn[neg, 0], conn[neg, 1] = conn[neg, 1].copy(), conn[neg, 0].copy()
        arrays = []
        for cf, name in zip(self.coefs, self.names):
            cf = CoefficientFunction._lift(cf)
            leaves = coef_leaves([s.as_coef() for s in cf.arr.reshape(-1)], 'field')
            if any(lf.val[0].space.mesh is not m and getattr(lf.val[0].space.mesh, 'points', None) is not m.points
                   for lf in leaves):
                raise ValueError('VTKOutput: a GridFunction lives on a different mesh')
            vals = evaluate_on_cells(cf, m, ref_pts).reshape(cf.arr.size, -1)          # (ncomp, ne*npts)
            if vals.shape[0] > 1:
                out = np.zeros((3, vals.shape[1]))
                out[:min(3, vals.shape[0])] = vals[:3]
                vals = out
            arrays.append((name, vals))
        path = self.filename + ('' if self.count == 0 else '_step{:05d}'.format(self.count)) + '.vtu'
        self.count += 1
        with open(path, 'w') as fh:
            fh.write('<?xml version="1.0"?>\n<VTKFile type="UnstructuredGrid" version="0.1" byte_order="LittleEndian">\n')
            fh.write('<UnstructuredGrid>\n<Piece NumberOfPoints="{}" NumberOfCells="{}">\n'.format(len(pts), len(conn)))
            fh.write('<Points>\n<DataArray type="Float64" NumberOfComponents="3" format="ascii">\n')
            np.savetxt(fh, pts, fmt='%.17g')
            fh.write('</DataArray>\n</Points>\n<Cells>\n<DataArray type="Int64" Name="connectivity" format="ascii">\n')
            np.savetxt(fh, conn, fmt='%d')
            fh.write('</DataArray>\n<DataArray type="Int64" Name="offsets" format="ascii">\n')
            np.savetxt(fh, (np.arange(len(conn)) + 1) * conn.shape[1], fmt='%d')
            fh.write('</DataArray>\n<DataArray type="UInt8" Name="types" format="ascii">\n')
            np.savetxt(fh, np.full(len(conn), _VTK_TYPE[m.cell_type]), fmt='%d')
            fh.write('</DataArray>\n</Cells>\n<PointData>\n')
            for name, vals in arrays:
                fh.write('<DataArray type="Float64" Name="{}" NumberOfComponents="{}" format="ascii">\n'
                         .format(name, vals.shape[0]))
                np.savetxt(fh, vals.T, fmt='%.17g')
                fh.write('</DataArray>\n')
            fh.write('</PointData>\n</Piece>\n</UnstructuredGrid>\n</VTKFile>\n')
        return path
