"""Finite-element spaces: global DOF numbering, cell DOF lists, Dirichlet masks, CSR pattern and scatter maps.

This is the "export once" half of the drop-in boundary: everything ``Model.__init__`` fixes for the whole run
(reference ``opencmp/models/base_model.py:191-205``; FES construction at ``models/poisson.py:60-65``,
``models/ins.py:95-128``) is turned into flat int32/float64 arrays here and uploaded by backend.py.

A space is a list of *blocks*. A block is one scalar (H1/L2) or one vector (HDiv) reference basis with a global DOF
offset: ``VectorH1`` contributes ``dim`` scalar blocks, a compound ``FESpace([...])`` concatenates the blocks of its
components. Global numbering inside a block is lowest-order entity DOFs first: H1 vertices, edges,
faces, cells; HDiv one normal DOF per facet, then higher-order facet DOFs, then cell DOFs; L2 cell by cell.

Dirichlet DOFs stay in the system and are masked by ``FreeDofs()`` (reference ``base_model.py:906``).
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np

from .basis import Basis, make_basis
from .mesh import Mesh, Region, local_topology


class Block:
    """One reference basis + its global numbering."""

    def __init__(self, mesh: Mesh, family: str, order: int, dirichlet: Optional[str], RT: bool = False):
        self.mesh = mesh
        self.family = family
        self.order = order
        self.basis: Basis = make_basis(family, mesh.cell_type, order, RT)
        self.kind = self.basis.kind
        self.nloc = self.basis.ndof
        self.dirichlet = dirichlet if dirichlet else None
        self._number()

    def _number(self) -> None:
        m = self.mesh
        ne = m.ne
        nloc = self.nloc
        cd = np.zeros((ne, nloc), dtype=np.int64)
        cells = np.arange(ne, dtype=np.int64)
        # count per entity class to lay out offsets
        ent = self.basis.entity_dofs
        per = {'v': 0, 'e': 0, 'f': 0, 'lo': 0, 'hi': 0, 'c': 0}
        for (_, _, cnt, tag) in ent:
            per[tag] = cnt if tag != 'c' else cnt
        nface = m.nf if m.dim == 3 else 0
        off = {}
        if self.kind == 'scalar' and self.family == 'H1':
            off['v'] = 0
            off['e'] = m.nv
            off['f'] = off['e'] + m.nedge * per['e']
            off['c'] = off['f'] + nface * per['f']
            self.ndof = off['c'] + ne * per['c']
        elif self.kind == 'hdiv':
            off['lo'] = 0
            off['hi'] = m.nf
            off['c'] = off['hi'] + m.nf * per['hi']
            self.ndof = off['c'] + ne * per['c']
        else:   # L2
            off['c'] = 0
            self.ndof = ne * per['c']
        self._off, self._per = off, per
        pos = 0
        for (edim, le, cnt, tag) in ent:
            k = np.arange(cnt, dtype=np.int64)[None, :]
            if tag == 'v':
                g = m.cells[:, le].astype(np.int64)[:, None] + k
            elif tag == 'e':
                g = off['e'] + m.cell_edges[:, le].astype(np.int64)[:, None] * cnt + k
            elif tag == 'f':
                g = off['f'] + m.cell_facets[:, le].astype(np.int64)[:, None] * cnt + k
            elif tag == 'lo':
                g = off['lo'] + m.cell_facets[:, le].astype(np.int64)[:, None] + k
            elif tag == 'hi':
                g = off['hi'] + m.cell_facets[:, le].astype(np.int64)[:, None] * cnt + k
            else:
                g = off['c'] + cells[:, None] * cnt + k
            cd[:, pos:pos + cnt] = g
            pos += cnt
        assert pos == nloc
        self.cell_dofs = cd
        # Dirichlet mask
        mask = np.zeros(self.ndof, dtype=bool)
        if self.dirichlet is not None and self.family != 'L2':
            reg = m.Boundaries(self.dirichlet).Mask()
            sel = reg[m.bnd_region]
            bf = m.bnd_facets[sel]
            mask[self.facet_trace_dofs(bf).ravel()] = True
        self.dirichlet_mask = mask

    # trace dofs of a set of boundary facets: (n, ntrace) global dofs, plus matching local dof indices
    def facet_trace_local(self) -> np.ndarray:
        """(nfc, ntrace) local dof indices whose functions do not vanish on local facet lf (H1) / carry its
        normal moments (HDiv)."""
        m = self.mesh
        loc = local_topology(m.cell_type)
        out = []
        for lf, fv in enumerate(loc['facets']):
            fv = set(fv)
            idx = []
            pos = 0
            for (edim, le, cnt, tag) in self.basis.entity_dofs:
                take = False
                if tag == 'v':
                    take = le in fv
                elif tag == 'e':
                    take = set(loc['edges'][le]) <= fv
                elif tag == 'f':
                    take = (le == lf)
                elif tag in ('lo', 'hi'):
                    take = (le == lf)
                if take:
                    idx += list(range(pos, pos + cnt))
                pos += cnt
            out.append(idx)
        return np.array(out, dtype=np.int32)

    def facet_trace_dofs(self, facets: np.ndarray) -> np.ndarray:
        m = self.mesh
        tl = self.facet_trace_local()
        c = m.facet_cells[facets, 0]
        lf = m.facet_local[facets, 0]
        return self.cell_dofs[c[:, None], tl[lf]]


class FESpace:
    """NGSolve-like space object. ``FESpace([a, b, ...])`` builds a compound space."""

    def __init__(self, spaces_or_mesh, order: int = 1, dirichlet: str = '', dgjumps: bool = False,
                 family: Optional[str] = None, RT: bool = False, **_ignored):
        if isinstance(spaces_or_mesh, (list, tuple)):
            comps: List[FESpace] = list(spaces_or_mesh)
            self.mesh = comps[0].mesh
            self.components = comps
            self.name = 'Compound'
            self.blocks: List[Block] = []
            self.comp_blocks: List[range] = []
            for c in comps:
                self.comp_blocks.append(range(len(self.blocks), len(self.blocks) + len(c.blocks)))
                self.blocks += c.blocks
            self.dgjumps = bool(dgjumps) or any(c.dgjumps for c in comps)
            self.order = max(c.order for c in comps)
            self.vector = [c.vector for c in comps]
        else:
            self.mesh = spaces_or_mesh
            self.name = family
            self.RT = bool(RT)
            self.order = int(order)
            self.dgjumps = bool(dgjumps)
            self.components = []
            if family == 'VectorH1':
                self.blocks = [Block(self.mesh, 'H1', order, dirichlet) for _ in range(self.mesh.dim)]
                self.vector = True
            else:
                self.blocks = [Block(self.mesh, family, order, dirichlet, RT)]
                self.vector = family == 'HDiv'
            self.comp_blocks = [range(0, len(self.blocks))]
        self._finalize()

    def _finalize(self) -> None:
        off = 0
        self.block_offsets = []
        for b in self.blocks:
            self.block_offsets.append(off)
            off += b.ndof
        self.ndof = off
        self.nloc = sum(b.nloc for b in self.blocks)
        self.loc_offsets = np.cumsum([0] + [b.nloc for b in self.blocks])[:-1].tolist()
        self._cell_dofs = None
        self._pattern = None

    # component offsets for compound spaces (GridFunction.components slices)
    def component_range(self, i: int) -> range:
        blks = self.comp_blocks[i]
        lo = self.block_offsets[blks[0]]
        hi = self.block_offsets[blks[-1]] + self.blocks[blks[-1]].ndof
        return range(lo, hi)

    @property
    def cell_dofs(self) -> np.ndarray:
        """(ne, nloc) int32 global dofs of every cell, blocks concatenated."""
        if self._cell_dofs is None:
            self._cell_dofs = np.ascontiguousarray(np.concatenate(
                [b.cell_dofs + o for b, o in zip(self.blocks, self.block_offsets)], axis=1), dtype=np.int32)
        return self._cell_dofs

    def FreeDofs(self) -> np.ndarray:
        return ~np.concatenate([b.dirichlet_mask for b in self.blocks])

    def Update(self) -> None:
        """Re-number after ``mesh.Refine()`` (reference post_processing/error_analysis.py:83-86)."""
        seen = set()
        for b in self.blocks:
            if id(b) not in seen:
                b._number()
                seen.add(id(b))
        for c in self.components:
            c._finalize()
        self._finalize()

    # ---- rows: physical operator rows per block ----------------------------------------------------------------
    @property
    def row_offsets(self) -> List[int]:
        out, r = [], 0
        for b in self.blocks:
            out.append(r)
            r += b.basis.nrows
        return out

    @property
    def nrows(self) -> int:
        return sum(b.basis.nrows for b in self.blocks)

    # ---- sparsity ---------------------------------------------------------------------------------------------
    def pattern(self) -> 'Pattern':
        if self._pattern is None:
            self._pattern = Pattern(self)
        return self._pattern


class Pattern:
    """CSR sparsity with every cell DOF pair (structural zeros kept, like NGSolve — SURVEY App. A) plus, for
    ``dgjumps`` spaces, every DOF pair of facet-neighbouring cells; and the element -> nnz scatter maps.

    rowptr (n+1) int32/int64, colidx (nnz) int32 sorted per row,
    cell2nnz (ne, nloc*nloc) int32: CSR position of local entry (i, j), row-major,
    facet2nnz (nif, 2, nloc*nloc) int32: positions of the off-diagonal blocks (rows of side s, columns of side 1-s)
    of every interior facet, in mesh.interior_facets order.
    """

    def __init__(self, fes: FESpace):
        m = fes.mesh
        n = np.int64(fes.ndof)
        cd = fes.cell_dofs.astype(np.int64)
        ne, nloc = cd.shape
        keys = [(cd[:, :, None] * n + cd[:, None, :]).reshape(-1)]
        if fes.dgjumps and len(m.interior_facets):
            fc = m.facet_cells[m.interior_facets]
            d0, d1 = cd[fc[:, 0]], cd[fc[:, 1]]
            keys.append((d0[:, :, None] * n + d1[:, None, :]).reshape(-1))
            keys.append((d1[:, :, None] * n + d0[:, None, :]).reshape(-1))
        ar = np.arange(int(n), dtype=np.int64)
        ukey, maps = _unique_and_locate(keys, keys + [ar * n + ar])
        self.nnz = int(ukey.shape[0])
        self.n = int(n)
        if self.nnz >= 2 ** 31:
            raise ValueError('pattern with more than 2^31 entries: partition the mesh across GPUs')
        rows = ukey // n
        self.colidx = np.ascontiguousarray(ukey - rows * n, dtype=np.int32)
        rowptr = np.zeros(self.n + 1, dtype=np.int64)
        rowptr[1:] = np.bincount(rows, minlength=self.n)
        self.rowptr = np.ascontiguousarray(np.cumsum(rowptr), dtype=np.int32)
        self.cell2nnz = np.ascontiguousarray(maps[0].reshape(ne, nloc * nloc), dtype=np.int32)
        if len(keys) > 1:
            nif = len(m.interior_facets)
            self.facet2nnz = np.ascontiguousarray(np.stack([maps[1].reshape(nif, nloc * nloc),
                                                            maps[2].reshape(nif, nloc * nloc)], axis=1), dtype=np.int32)
        else:
            self.facet2nnz = np.zeros((0, 2, nloc * nloc), dtype=np.int32)
        # position of the diagonal entries (Jacobi, Dirichlet handling)
        self.diag = np.ascontiguousarray(maps[-1], dtype=np.int32)


# above this many keys the sort / search runs in slices of the key range: keeps every single torch call well below
# 2^31 elements (the 3-D N = 64 pattern has 2.08e9 keys) and bounds the temporary device memory. Everything measured
# in round 1 (up to the 0.88e9 keys of the 3-D N = 48 pattern) stays on the one-shot path.
_SLICE_KEYS = 1 << 30


def _unique_and_locate(key_lists, queries, device=None, slice_keys=None):
    """Sorted unique int64 keys of ``key_lists`` and the position of every query key in them. One-off set-up work:
    done with torch on the GPU when one is visible (sort / searchsorted of 10^8 keys), with NumPy otherwise."""
    total = sum(k.size for k in key_lists)
    slice_keys = _SLICE_KEYS if slice_keys is None else slice_keys
    if total > (1 << 22) or device is not None:
        try:
            import torch
            if device is not None or torch.cuda.is_available():
                dev = device if device is not None else torch.device('cuda', torch.cuda.current_device())
                if total > slice_keys:
                    return _unique_and_locate_sliced(torch, dev, key_lists, queries, -(-total // slice_keys))
                ukey = torch.unique(torch.cat([torch.from_numpy(k).to(dev) for k in key_lists]))
                maps = [torch.searchsorted(ukey, torch.from_numpy(np.ascontiguousarray(q)).to(dev)).to(torch.int32)
                        .cpu().numpy() for q in queries]
                return ukey.cpu().numpy(), maps
        except ImportError:
            pass
    ukey = np.unique(np.concatenate(key_lists))
    return ukey, [np.searchsorted(ukey, q) for q in queries]


def _unique_and_locate_sliced(torch, dev, key_lists, queries, nslices):
    """The same in ``nslices`` slices of the key range [lo, hi): the unique keys of a slice are a contiguous run of the
    global result, so slices are sorted independently and concatenated; a query is located inside its own slice and
    offset by the number of unique keys before it. Keys stay on the host between slices."""
    lo = min(int(k.min()) for k in key_lists if k.size)
    hi = max(int(k.max()) for k in key_lists if k.size) + 1
    edges = [lo + (hi - lo) * i // nslices for i in range(nslices + 1)]
    edges[-1] = hi
    qs = [np.ascontiguousarray(q) for q in queries]
    maps = [np.empty(q.shape, dtype=np.int32) for q in qs]
    parts, base = [], 0
    for a, b in zip(edges[:-1], edges[1:]):
        sel = [k[(k >= a) & (k < b)] for k in key_lists]
        sel = [x for x in sel if x.size]
        if not sel:
            continue
        u = torch.unique(torch.cat([torch.from_numpy(np.ascontiguousarray(x)).to(dev) for x in sel]))
        for q, m in zip(qs, maps):
            idx = np.nonzero((q >= a) & (q < b))[0]
            if idx.size:
                loc = torch.searchsorted(u, torch.from_numpy(q[idx]).to(dev)) + base
                m[idx] = loc.to(torch.int32).cpu().numpy()
        parts.append(u.cpu().numpy())
        base += int(u.numel())
        del u
    return np.concatenate(parts), maps


# NGSolve-style constructors -----------------------------------------------------------------------------------
def H1(mesh: Mesh, order: int = 1, dirichlet: str = '', dgjumps: bool = False, **kw) -> FESpace:
    return FESpace(mesh, order=order, dirichlet=dirichlet, dgjumps=dgjumps, family='H1')


def VectorH1(mesh: Mesh, order: int = 1, dirichlet: str = '', dgjumps: bool = False, **kw) -> FESpace:
    return FESpace(mesh, order=order, dirichlet=dirichlet, dgjumps=dgjumps, family='VectorH1')


def L2(mesh: Mesh, order: int = 0, dirichlet: str = '', dgjumps: bool = False, **kw) -> FESpace:
    return FESpace(mesh, order=order, dirichlet='', dgjumps=dgjumps, family='L2')


def HDiv(mesh: Mesh, order: int = 1, dirichlet: str = '', dgjumps: bool = False, RT: bool = False, **kw) -> FESpace:
    return FESpace(mesh, order=order, dirichlet=dirichlet, dgjumps=dgjumps, family='HDiv', RT=RT)
