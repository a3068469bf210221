"""Mesh export layer (host side, NumPy).

Builds the arrays the CUDA path consumes once per run: vertex coordinates, cell connectivity, facet
connectivity with the two neighbouring cells, boundary-facet region ids, and (for 3-D H1 spaces) edges.

Replaces what OpenCMP gets from ``ngs.Mesh(filename)`` (reference ``opencmp/helpers/io.py:69-110``) and from the
structured generator ``get_Netgen_nonconformal`` (reference ``opencmp/diffuse_interface/mesh_helpers.py:494-690``).

Conventions (ours, chosen so one reference tabulation serves every cell):

* simplices store their vertices sorted by ascending global vertex number, so every local edge/face runs from the
  lower to the higher global vertex on *both* neighbouring cells; det(J) may be negative.
* tensor-product cells (quad/hex) are only produced by the structured generators, whose node numbering makes every
  local axis point along +x/+y/+z; the reference cell is [0,1]^d and local vertex ``l`` has bits (a,b,c) = (xi,eta,zeta).
"""
from __future__ import annotations

import re
from typing import Dict, List, Optional, Sequence

import numpy as np

# local topology tables ------------------------------------------------------------------------------------------
_LOCAL = {
    'tri': dict(dim=2, nv=3, facets=[(0, 1), (0, 2), (1, 2)], edges=[(0, 1), (0, 2), (1, 2)],
                ref=np.array([[0., 0.], [1., 0.], [0., 1.]])),
    'quad': dict(dim=2, nv=4, facets=[(0, 1), (2, 3), (0, 2), (1, 3)], edges=[(0, 1), (2, 3), (0, 2), (1, 3)],
                 ref=np.array([[0., 0.], [1., 0.], [0., 1.], [1., 1.]])),
    'tet': dict(dim=3, nv=4, facets=[(0, 1, 2), (0, 1, 3), (0, 2, 3), (1, 2, 3)],
                edges=[(0, 1), (0, 2), (0, 3), (1, 2), (1, 3), (2, 3)],
                ref=np.array([[0., 0., 0.], [1., 0., 0.], [0., 1., 0.], [0., 0., 1.]])),
    'hex': dict(dim=3, nv=8,
                # faces listed as the 4 local vertices in (low bits first) order; face f: axis = f//2, side = f%2
                facets=[(0, 2, 4, 6), (1, 3, 5, 7), (0, 1, 4, 5), (2, 3, 6, 7), (0, 1, 2, 3), (4, 5, 6, 7)],
                edges=[(0, 1), (2, 3), (4, 5), (6, 7), (0, 2), (1, 3), (4, 6), (5, 7), (0, 4), (1, 5), (2, 6), (3, 7)],
                ref=np.array([[a, b, c] for c in (0., 1.) for b in (0., 1.) for a in (0., 1.)])),
}


def local_topology(cell_type: str) -> dict:
    return _LOCAL[cell_type]


class Region:
    """A set of boundary (or material) indices, the analogue of ``mesh.Boundaries('a|b')``."""

    def __init__(self, mesh: 'Mesh', ids: Sequence[int], kind: str = 'bnd'):
        self.mesh = mesh
        self.ids = tuple(sorted(set(int(i) for i in ids)))
        self.kind = kind

    def Mask(self) -> np.ndarray:
        n = len(self.mesh.bnd_names) if self.kind == 'bnd' else len(self.mesh.mat_names)
        m = np.zeros(n, dtype=bool)
        m[list(self.ids)] = True
        return m

    def __add__(self, other: 'Region') -> 'Region':
        return Region(self.mesh, self.ids + other.ids, self.kind)

    def __repr__(self) -> str:
        return 'Region({}, {})'.format(self.kind, self.ids)


class Mesh:
    """Unstructured conforming mesh of one cell type with facet topology.

    Attributes
    ----------
    dim, cell_type
    points        (np, dim) float64
    cells         (ne, nv) int32 — simplices sorted ascending per row
    cell_mat      (ne,) int32 material index (0-based into mat_names)
    facets        (nf, nfv) int32 unique facets (vertex tuples sorted for simplices)
    facet_cells   (nf, 2) int32 neighbouring cells, -1 if none; facet_cells[:,0] is the "this" side
    facet_local   (nf, 2) int32 local facet number inside each neighbouring cell
    cell_facets   (ne, nfc) int32
    bnd_facets    (nb,) int32 facet ids on the boundary, bnd_region (nb,) int32 region index into bnd_names
    edges, cell_edges   (3-D only; in 2-D edges are the facets)
    """

    def __init__(self, dim: int, cell_type: str, points: np.ndarray, cells: np.ndarray,
                 bnd_elems: np.ndarray, bnd_index: np.ndarray, bnd_names: List[str],
                 cell_mat: Optional[np.ndarray] = None, mat_names: Optional[List[str]] = None,
                 structured: Optional[dict] = None):
        self.dim = dim
        self.cell_type = cell_type
        self.points = np.ascontiguousarray(points[:, :dim], dtype=np.float64)
        cells = np.asarray(cells, dtype=np.int64)
        if cell_type in ('tri', 'tet'):
            cells = np.sort(cells, axis=1)
        self.cells = np.ascontiguousarray(cells, dtype=np.int32)
        self.ne = self.cells.shape[0]
        self.nv = self.points.shape[0]
        self.cell_mat = np.zeros(self.ne, np.int32) if cell_mat is None else np.asarray(cell_mat, np.int32)
        self.mat_names = list(mat_names) if mat_names is not None else ['default']
        self.bnd_names = list(bnd_names)
        self.structured = structured
        self._build_topology(np.asarray(bnd_elems, dtype=np.int64), np.asarray(bnd_index, dtype=np.int32))

    # ---------------------------------------------------------------------------------------------------------
    def _build_topology(self, bnd_elems: np.ndarray, bnd_index: np.ndarray) -> None:
        loc = _LOCAL[self.cell_type]
        lf = np.array(loc['facets'], dtype=np.int64)            # (nfc, nfv)
        nfc, nfv = lf.shape
        ne = self.ne
        cf = self.cells.astype(np.int64)[:, lf]                  # (ne, nfc, nfv)
        key_src = np.sort(cf, axis=2).reshape(ne * nfc, nfv)
        base = np.int64(self.nv)
        # a quadrilateral face is identified by its three lowest vertices (keeps the key inside int64 for large meshes)
        nkey = min(nfv, 3)
        key = np.zeros(ne * nfc, dtype=np.int64)
        for k in range(nkey):
            key = key * base + key_src[:, k]
        # unique facets ordered by key (deterministic), remember first occurrence for vertex order
        ukey, first, inv = np.unique(key, return_index=True, return_inverse=True)
        nf = ukey.shape[0]
        self.nf = nf
        self.facets = np.ascontiguousarray(cf.reshape(ne * nfc, nfv)[first], dtype=np.int32)
        self.cell_facets = np.ascontiguousarray(inv.reshape(ne, nfc), dtype=np.int32)
        # neighbouring cells: lowest cell number first ("this" side)
        order = np.argsort(inv, kind='stable')
        sorted_f = inv[order]
        cell_of = (order // nfc).astype(np.int32)
        loc_of = (order % nfc).astype(np.int32)
        start = np.searchsorted(sorted_f, np.arange(nf))
        count = np.diff(np.append(start, sorted_f.shape[0]))
        if count.max() > 2:
            raise ValueError('non-manifold mesh: a facet has more than two neighbouring cells')
        self.facet_cells = -np.ones((nf, 2), dtype=np.int32)
        self.facet_local = -np.ones((nf, 2), dtype=np.int32)
        self.facet_cells[:, 0] = cell_of[start]
        self.facet_local[:, 0] = loc_of[start]
        two = count == 2
        self.facet_cells[two, 1] = cell_of[start[two] + 1]
        self.facet_local[two, 1] = loc_of[start[two] + 1]
        self.interior_facets = np.nonzero(two)[0].astype(np.int32)
        bfac = np.nonzero(~two)[0].astype(np.int32)
        # boundary regions from the boundary elements of the file
        region = -np.ones(nf, dtype=np.int32)
        if bnd_elems.size:
            bk = np.zeros(bnd_elems.shape[0], dtype=np.int64)
            bs = np.sort(bnd_elems, axis=1)
            for k in range(nkey):
                bk = bk * base + bs[:, k]
            pos = np.searchsorted(ukey, bk)
            ok = (pos < nf) & (ukey[np.minimum(pos, nf - 1)] == bk)
            region[pos[ok]] = bnd_index[ok]
        if np.any(region[bfac] < 0):
            # unnamed boundary facets get region 'default'
            if 'default' not in self.bnd_names:
                self.bnd_names.append('default')
            region[bfac[region[bfac] < 0]] = self.bnd_names.index('default')
        self.bnd_facets = bfac
        self.bnd_region = np.ascontiguousarray(region[bfac], dtype=np.int32)
        # edges
        if self.dim == 2:
            self.edges = self.facets
            self.cell_edges = self.cell_facets
            self.nedge = nf
        else:
            le = np.array(loc['edges'], dtype=np.int64)
            ce = np.sort(self.cells.astype(np.int64)[:, le], axis=2)
            ek = ce[:, :, 0].reshape(-1) * base + ce[:, :, 1].reshape(-1)
            uek, efirst, einv = np.unique(ek, return_index=True, return_inverse=True)
            self.edges = np.ascontiguousarray(ce.reshape(-1, 2)[efirst], dtype=np.int32)
            self.cell_edges = np.ascontiguousarray(einv.reshape(ne, le.shape[0]), dtype=np.int32)
            self.nedge = uek.shape[0]

    # ---- NGSolve-like accessors (reference call sites: base_model.py:191-198, boundary_conditions.py:129) -----
    def GetBoundaries(self) -> tuple:
        return tuple(self.bnd_names)

    def GetMaterials(self) -> tuple:
        return tuple(self.mat_names)

    def Boundaries(self, pattern: str) -> Region:
        ids = [i for i, name in enumerate(self.bnd_names) if _match(pattern, name)]
        return Region(self, ids, 'bnd')

    def Materials(self, pattern: str) -> Region:
        ids = [i for i, name in enumerate(self.mat_names) if _match(pattern, name)]
        return Region(self, ids, 'mat')

    # ---- geometry -------------------------------------------------------------------------------------------
    def jacobians(self) -> np.ndarray:
        """(ne, dim, dim) affine Jacobians J[e, i, a] = d x_i / d xi_a."""
        P = self.points
        c = self.cells
        if self.cell_type in ('tri', 'tet'):
            return np.stack([P[c[:, a + 1]] - P[c[:, 0]] for a in range(self.dim)], axis=2)
        if self.cell_type == 'quad':
            return np.stack([P[c[:, 1]] - P[c[:, 0]], P[c[:, 2]] - P[c[:, 0]]], axis=2)
        return np.stack([P[c[:, 1]] - P[c[:, 0]], P[c[:, 2]] - P[c[:, 0]], P[c[:, 4]] - P[c[:, 0]]], axis=2)

    def origins(self) -> np.ndarray:
        return self.points[self.cells[:, 0]]

    def check_affine(self, tol: float = 1e-12) -> None:
        if self.cell_type in ('tri', 'tet'):
            return
        P, c = self.points, self.cells
        J = self.jacobians()
        ref = _LOCAL[self.cell_type]['ref']
        pred = P[c[:, 0]][:, None, :] + np.einsum('eia,la->eli', J, ref)
        err = np.abs(pred - P[c]).max()
        if err > tol * max(1.0, np.abs(P).max()):
            raise ValueError('quad/hex cells must be affine (parallelograms / parallelepipeds)')

    # ---- uniform refinement (reference: post_processing/error_analysis.py:83-86 calls mesh.Refine()) ---------
    def _refine_tensor(self) -> 'Mesh':
        """Uniform refinement of a quad / hex mesh: every cell is cut into 2^d children on its 3^d lattice. New vertices
        are keyed by the set of parent vertices they average (edge midpoints, face centres, cell centres), so shared
        ones are created once. Children of cell e are the fine cells 2^d e .. 2^d e + 2^d - 1."""
        import itertools
        d = self.dim
        nvc = 2 ** d
        S = [(0,), (0, 1), (1,)]

        def lattice_sets(dd):
            out = []
            for idx in itertools.product(range(3), repeat=dd):          # idx = (i_{dd-1}, ..., i_0), i_0 fastest
                axes = idx[::-1]
                out.append([sum(b << a for a, b in enumerate(bits))
                            for bits in itertools.product(*[S[i] for i in axes])])
            return out

        def lattice_ids(elems, dd, table):
            """(n, 3^dd) global vertex ids of the lattice points of each element (corner ids for corners, new ids from
            ``table`` — a dict filled on the fly — otherwise)."""
            n = elems.shape[0]
            sets = lattice_sets(dd)
            keys = -np.ones((n, len(sets), nvc), dtype=np.int64)
            for l, sub in enumerate(sets):
                keys[:, l, :len(sub)] = elems[:, sub]
            keys = np.sort(keys, axis=2)
            return keys

        c = self.cells.astype(np.int64)
        ck = lattice_ids(c, d, None)                                      # (ne, 3^d, nvc)
        bf = self.facets[self.bnd_facets].astype(np.int64)
        bk = lattice_ids(bf, d - 1, None)                                 # (nb, 3^(d-1), nvc)
        allk = np.concatenate([ck.reshape(-1, nvc), bk.reshape(-1, nvc)], axis=0)
        corner = allk[:, -2] < 0                                          # exactly one parent vertex
        uk, inv = _unique_rows(allk[~corner])
        ids = np.empty(allk.shape[0], dtype=np.int64)
        ids[corner] = allk[corner, -1]
        ids[~corner] = self.nv + inv.reshape(-1)
        cnt = (uk >= 0).sum(axis=1)
        newP = np.where(uk[:, :, None] >= 0, self.points[np.maximum(uk, 0)], 0.0).sum(axis=1) / cnt[:, None]
        P = np.vstack([self.points, newP])
        cl = ids[:ck.shape[0] * ck.shape[1]].reshape(self.ne, -1)        # (ne, 3^d)
        bl = ids[ck.shape[0] * ck.shape[1]:].reshape(len(bf), -1)        # (nb, 3^(d-1))

        def children(lat, dd):
            out = []
            for off in itertools.product(range(2), repeat=dd):            # child offset (o_{dd-1}, ..., o_0)
                o = off[::-1]
                verts = []
                for bits in itertools.product(range(2), repeat=dd):
                    bt = bits[::-1]
                    verts.append(sum((o[a] + bt[a]) * 3 ** a for a in range(dd)))
                out.append(lat[:, verts])
            return np.stack(out, axis=1)                                   # (n, 2^dd, 2^dd)

        cells = children(cl, d).reshape(-1, nvc)
        bnd = children(bl, d - 1).reshape(-1, 2 ** (d - 1))
        bidx = np.repeat(self.bnd_region, 2 ** (d - 1))
        return Mesh(d, self.cell_type, P, cells, bnd, bidx, self.bnd_names, np.repeat(self.cell_mat, nvc),
                    self.mat_names)

    def Refine(self) -> 'Mesh':
        if self.cell_type in ('quad', 'hex'):
            fine = self._refine_tensor()
            coarse = Mesh.__new__(Mesh)
            coarse.__dict__.update({k: v for k, v in self.__dict__.items() if k != '_b200_scalar_space'})
            self.__dict__.update(fine.__dict__)
            self.__dict__.pop('_b200_scalar_space', None)
            self.coarse = coarse
            return self
        if self.cell_type != 'tri':
            raise NotImplementedError('Refine is implemented for triangle, quadrilateral and hexahedral meshes')
        P = self.points
        mid = 0.5 * (P[self.facets[:, 0]] + P[self.facets[:, 1]])
        newP = np.vstack([P, mid])
        m = self.cell_facets.astype(np.int64) + self.nv          # local edges (0,1),(0,2),(1,2)
        c = self.cells.astype(np.int64)
        v0, v1, v2 = c[:, 0], c[:, 1], c[:, 2]
        m01, m02, m12 = m[:, 0], m[:, 1], m[:, 2]
        cells = np.concatenate([np.stack([v0, m01, m02], 1), np.stack([v1, m01, m12], 1),
                                np.stack([v2, m02, m12], 1), np.stack([m01, m02, m12], 1)], axis=0)
        # interleave so children of a cell stay together
        cells = cells.reshape(4, self.ne, 3).transpose(1, 0, 2).reshape(-1, 3)
        bf = self.facets[self.bnd_facets].astype(np.int64)
        bm = self.bnd_facets.astype(np.int64) + self.nv
        bnd = np.concatenate([np.stack([bf[:, 0], bm], 1), np.stack([bm, bf[:, 1]], 1)], axis=0)
        bidx = np.concatenate([self.bnd_region, self.bnd_region])
        fine = Mesh(2, 'tri', newP, cells, bnd, bidx, self.bnd_names,
                    np.repeat(self.cell_mat, 4), self.mat_names)
        # keep the coarse level for geometric multigrid: children of coarse cell e are fine cells 4e .. 4e+3
        coarse = Mesh.__new__(Mesh)
        coarse.__dict__.update({k: v for k, v in self.__dict__.items() if k != '_b200_scalar_space'})
        self.__dict__.update(fine.__dict__)
        self.__dict__.pop('_b200_scalar_space', None)
        self.coarse = coarse
        return self

    def Curve(self, order: int) -> None:
        """The shipped .vol files carry no geometry section, so curving is a no-op (straight-sided cells)."""
        return None


def _unique_rows(rows: np.ndarray):
    """np.unique(rows, axis=0, return_inverse=True) for int64 rows, through a 64-bit multiplicative hash of each row
    (a 1-D sort instead of a lexicographic one: ~8x faster on the 10^7 lattice keys of a refined hex mesh). The result
    is verified; a hash collision falls back to the lexicographic version."""
    rng = np.random.default_rng(0x5eed)
    mult = (rng.integers(1, 2 ** 62, size=rows.shape[1], dtype=np.int64) * 2 + 1).astype(np.uint64)
    with np.errstate(over='ignore'):
        h = (rows.astype(np.uint64) * mult[None, :]).sum(axis=1, dtype=np.uint64)
    _, first, inv = np.unique(h, return_index=True, return_inverse=True)
    uk = rows[first]
    if not np.array_equal(uk[inv], rows):
        return np.unique(rows, axis=0, return_inverse=True)
    # keep np.unique's lexicographic order of the unique rows so numbering does not depend on the hash
    order = np.lexsort(uk.T[::-1])
    rank = np.empty(len(order), dtype=np.int64)
    rank[order] = np.arange(len(order))
    return uk[order], rank[inv.reshape(-1)]


def _match(pattern: str, name: str) -> bool:
    if pattern is None:
        return False
    for alt in str(pattern).split('|'):
        if alt == name:
            return True
        if any(ch in alt for ch in '.*[]?+') and re.fullmatch(alt, name):
            return True
    return False


# ---------------------------------------------------------------------------------------------------------------
# Readers
# ---------------------------------------------------------------------------------------------------------------
def read_vol(path: str) -> Mesh:
    """Netgen ASCII ``.vol`` reader (format example: reference ``examples/Poisson/unit_square_coarse.vol``)."""
    with open(path, 'r') as fh:
        lines = [ln.strip() for ln in fh]
    sec: Dict[str, List[List[str]]] = {}
    scalars: Dict[str, str] = {}
    i = 0
    counted = {'surfaceelements', 'surfaceelementsgi', 'surfaceelementsuv', 'volumeelements', 'edgesegmentsgi2',
               'edgesegments', 'points', 'bcnames', 'materials', 'pointelements', 'face_colours', 'cd2names',
               'identifications', 'identificationtypes', 'singular_points', 'singular_edge_left',
               'singular_edge_right', 'singular_face_inside', 'singular_face_outside', 'cd3names'}
    while i < len(lines):
        ln = lines[i]
        if not ln or ln.startswith('#'):
            i += 1
            continue
        if ln in ('dimension', 'geomtype'):
            scalars[ln] = lines[i + 1]
            i += 2
            continue
        if ln in counted:
            n = int(lines[i + 1].split()[0])
            rows = []
            j = i + 2
            while len(rows) < n:
                if lines[j] and not lines[j].startswith('#'):
                    rows.append(lines[j].split())
                j += 1
            sec[ln] = rows
            i = j
            continue
        i += 1
    dim = int(scalars.get('dimension', '3'))
    pts = np.array([[float(x) for x in r[:3]] for r in sec['points']], dtype=np.float64)
    if dim == 2:
        surf = sec.get('surfaceelements') or sec.get('surfaceelementsgi') or sec.get('surfaceelementsuv')
        nps = {int(r[4]) for r in surf}
        if nps == {3}:
            ctype = 'tri'
        elif nps == {4}:
            ctype = 'quad'
        else:
            raise ValueError('mixed or unsupported 2-D cell types: {}'.format(nps))
        cells = np.array([[int(x) - 1 for x in r[5:5 + int(r[4])]] for r in surf], dtype=np.int64)
        if ctype == 'quad':
            cells = cells[:, [0, 1, 3, 2]]       # ccw -> bit ordering
        segs = sec.get('edgesegmentsgi2') or sec.get('edgesegments') or []
        bnd = np.array([[int(r[2]) - 1, int(r[3]) - 1] for r in segs], dtype=np.int64).reshape(-1, 2)
        bidx = np.array([int(r[0]) - 1 for r in segs], dtype=np.int32)
        names = _names(sec.get('bcnames'), bidx)
        mat = np.zeros(len(cells), np.int32)
        mesh = Mesh(2, ctype, pts, cells, bnd, bidx, names, mat, _names(sec.get('materials'), mat))
        if ctype == 'quad':
            _orient_quads(mesh)
        return mesh
    vol = sec['volumeelements']
    nps = {int(r[1]) for r in vol}
    if nps != {4}:
        raise ValueError('only tetrahedral 3-D .vol meshes are supported by the reader')
    cells = np.array([[int(x) - 1 for x in r[2:6]] for r in vol], dtype=np.int64)
    mat = np.array([int(r[0]) - 1 for r in vol], dtype=np.int32)
    surf = sec.get('surfaceelements') or sec.get('surfaceelementsgi') or []
    bnd = np.array([[int(x) - 1 for x in r[5:8]] for r in surf], dtype=np.int64).reshape(-1, 3)
    bidx = np.array([int(r[1]) - 1 for r in surf], dtype=np.int32)
    return Mesh(3, 'tet', pts, cells, bnd, bidx, _names(sec.get('bcnames'), bidx), mat,
                _names(sec.get('materials'), mat))


def _names(rows, idx) -> List[str]:
    n = int(idx.max()) + 1 if len(idx) else 0
    names = ['default'] * n
    if rows:
        for r in rows:
            k = int(r[0]) - 1
            if k >= len(names):
                names.extend(['default'] * (k + 1 - len(names)))
            names[k] = r[1] if len(r) > 1 else 'default'
    return names


def _orient_quads(mesh: Mesh) -> None:
    """Re-number the local vertices of axis-consistent quads so that local axes follow ascending vertex numbers."""
    c = mesh.cells.astype(np.int64)
    # choose the local origin = smallest vertex; then xi-neighbour < eta-neighbour
    out = c.copy()
    nbr = {0: (1, 2, 3), 1: (0, 3, 2), 2: (0, 3, 1), 3: (1, 2, 0)}
    for e in range(c.shape[0]):
        o = int(np.argmin(c[e]))
        a, b, d = nbr[o]
        if c[e, a] > c[e, b]:
            a, b = b, a
        out[e] = [c[e, o], c[e, a], c[e, b], c[e, d]]
    fresh = Mesh(2, 'quad', mesh.points, out, mesh.facets[mesh.bnd_facets], mesh.bnd_region, mesh.bnd_names,
                 mesh.cell_mat, mesh.mat_names)
    mesh.__dict__.update(fresh.__dict__)


def read_msh(path: str) -> Mesh:
    """Gmsh 2.2 ASCII reader (triangles in 2-D, tetrahedra + boundary triangles in 3-D)."""
    with open(path, 'r') as fh:
        txt = fh.read().split('\n')
    i = 0
    phys: Dict[tuple, str] = {}
    nodes = None
    node_ids = None
    elems: List[List[int]] = []
    while i < len(txt):
        ln = txt[i].strip()
        if ln == '$PhysicalNames':
            n = int(txt[i + 1])
            for r in txt[i + 2:i + 2 + n]:
                d, tag, name = r.split(None, 2)
                phys[(int(d), int(tag))] = name.strip().strip('"')
            i += n + 2
        elif ln == '$Nodes':
            n = int(txt[i + 1])
            arr = np.array([r.split() for r in txt[i + 2:i + 2 + n]], dtype=np.float64)
            node_ids = arr[:, 0].astype(np.int64)
            nodes = arr[:, 1:4]
            i += n + 2
        elif ln == '$Elements':
            n = int(txt[i + 1])
            elems = [[int(x) for x in r.split()] for r in txt[i + 2:i + 2 + n]]
            i += n + 2
        else:
            i += 1
    remap = np.zeros(int(node_ids.max()) + 1, dtype=np.int64)
    remap[node_ids] = np.arange(len(node_ids))
    by_type: Dict[int, list] = {}
    for e in elems:
        et, ntags = e[1], e[2]
        by_type.setdefault(et, []).append((e[3] if ntags else 0, [remap[v] for v in e[3 + ntags:]]))
    if 4 in by_type:
        dim, ctype, ct, bt = 3, 'tet', 4, 2
    else:
        dim, ctype, ct, bt = 2, 'tri', 2, 1
    cells = np.array([v for _, v in by_type[ct]], dtype=np.int64)
    ctag = [t for t, _ in by_type[ct]]
    btag = [t for t, _ in by_type.get(bt, [])]
    bnd = np.array([v for _, v in by_type.get(bt, [])], dtype=np.int64).reshape(-1, dim)
    utag = sorted(set(btag))
    names = [phys.get((dim - 1, t), 'default') for t in utag]
    bidx = np.array([utag.index(t) for t in btag], dtype=np.int32)
    umat = sorted(set(ctag))
    mat = np.array([umat.index(t) for t in ctag], dtype=np.int32)
    used = np.unique(cells)
    comp = -np.ones(len(nodes), dtype=np.int64)
    comp[used] = np.arange(len(used))
    return Mesh(dim, ctype, nodes[used], comp[cells], comp[bnd], bidx, names, mat,
                [phys.get((dim, t), 'default') for t in umat])


def load_mesh(path: str) -> Mesh:
    if path.endswith('.vol'):
        return read_vol(path)
    if path.endswith('.msh'):
        return read_msh(path)
    raise TypeError('Only .vol (Netgen) and .msh (Gmsh) meshes can be used.')


# ---------------------------------------------------------------------------------------------------------------
# Structured generators (numbering spec: reference diffuse_interface/mesh_helpers.py:512-683)
# ---------------------------------------------------------------------------------------------------------------
def structured_2d(N: Sequence[int], scale: Sequence[float] = (1.0, 1.0), offset: Sequence[float] = (0.0, 0.0),
                  cell: str = 'tri', pattern: str = 'diag') -> Mesh:
    """Nx x Ny structured mesh of [-off, -off+scale]^2.

    cell='quad' : one quad per square (reference ``quad=True`` branch, mesh_helpers.py:540-549)
    cell='tri', pattern='diag'  : 2 triangles per square (the synthetic throughput meshes of SURVEY 8(d))
    cell='tri', pattern='cross' : 4 triangles around an added centre node (reference ``quad=False`` branch; the caller
                                   halves N like mesh_helpers.py:518-519 if it wants the reference's element count)
    Boundary names bottom,right,top,left (mesh_helpers.py:533-537).
    """
    nx, ny = int(N[0]), int(N[1])
    jj, ii = np.meshgrid(np.arange(nx + 1), np.arange(ny + 1))
    x = -offset[0] + scale[0] * jj / nx
    y = -offset[1] + scale[1] * ii / ny
    pts = np.stack([x.ravel(), y.ravel()], axis=1)
    j, i = np.meshgrid(np.arange(nx), np.arange(ny))
    p1 = (i * (nx + 1) + j).ravel()
    p2 = p1 + 1
    p4 = p1 + nx + 1
    p3 = p4 + 1
    if cell == 'quad':
        cells = np.stack([p1, p2, p4, p3], axis=1)
        ctype = 'quad'
    elif pattern == 'diag':
        cells = np.stack([np.stack([p1, p2, p3], 1), np.stack([p1, p3, p4], 1)], axis=1).reshape(-1, 3)
        ctype = 'tri'
    else:
        cx = 0.5 * (pts[p1, 0] + pts[p2, 0])
        cy = 0.5 * (pts[p1, 1] + pts[p3, 1])
        c = pts.shape[0] + np.arange(p1.shape[0])
        pts = np.vstack([pts, np.stack([cx, cy], 1)])
        cells = np.stack([np.stack([p1, p2, c], 1), np.stack([p2, p3, c], 1), np.stack([p3, p4, c], 1),
                          np.stack([p4, p1, c], 1)], axis=1).reshape(-1, 3)
        ctype = 'tri'
    jb = np.arange(nx)
    ib = np.arange(ny)
    bottom = np.stack([jb, jb + 1], 1)
    top = np.stack([ny * (nx + 1) + jb, ny * (nx + 1) + jb + 1], 1)
    right = np.stack([nx + ib * (nx + 1), nx + (ib + 1) * (nx + 1)], 1)
    left = np.stack([ib * (nx + 1), (ib + 1) * (nx + 1)], 1)
    bnd = np.vstack([bottom, right, top, left])
    bidx = np.concatenate([np.full(nx, 0), np.full(ny, 1), np.full(nx, 2), np.full(ny, 3)]).astype(np.int32)
    info = dict(N=(nx, ny), scale=tuple(scale), offset=tuple(offset), cell=cell, pattern=pattern)
    return Mesh(2, ctype, pts, cells, bnd, bidx, ['bottom', 'right', 'top', 'left'], structured=info)


def structured_3d(N: Sequence[int], scale: Sequence[float] = (1.0, 1.0, 1.0),
                  offset: Sequence[float] = (0.0, 0.0, 0.0), cell: str = 'hex') -> Mesh:
    """Nx x Ny x Nz structured box; node numbering i*(Ny+1)*(Nz+1) + j*(Nz+1) + k (mesh_helpers.py:584-591).

    cell='hex': one hexahedron per cube (mesh_helpers.py:597-613). cell='tet': 6 tetrahedra per cube (Kuhn split;
    the reference's non-quad branch builds 5-vertex pyramids, which no finite-element space in scope supports).
    Boundary names back,left,front,right,bottom,top (mesh_helpers.py:681-683).
    """
    nx, ny, nz = (int(n) for n in N)
    sy, sx = (nz + 1), (ny + 1) * (nz + 1)
    I, J, K = np.meshgrid(np.arange(nx + 1), np.arange(ny + 1), np.arange(nz + 1), indexing='ij')
    pts = np.stack([-offset[0] + scale[0] * I.ravel() / nx, -offset[1] + scale[1] * J.ravel() / ny,
                    -offset[2] + scale[2] * K.ravel() / nz], axis=1)
    i, j, k = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing='ij')
    base = (i * sx + j * sy + k).ravel()
    corner = lambda a, b, c: base + a * sx + b * sy + c
    if cell == 'hex':
        cells = np.stack([corner(a, b, c) for c in (0, 1) for b in (0, 1) for a in (0, 1)], axis=1)
        ctype = 'hex'
    else:
        import itertools
        tets = []
        for perm in itertools.permutations(range(3)):
            v = [np.zeros(3, int)]
            for ax in perm:
                w = v[-1].copy()
                w[ax] = 1
                v.append(w)
            tets.append(np.stack([corner(*w) for w in v], axis=1))
        cells = np.stack(tets, axis=1).reshape(-1, 4)
        ctype = 'tet'

    def face(fix_axis, side):
        ax = [0, 1, 2]
        ax.remove(fix_axis)
        n = [nx, ny, nz]
        s = [sx, sy, 1]
        a, b = np.meshgrid(np.arange(n[ax[0]]), np.arange(n[ax[1]]), indexing='ij')
        b0 = (side * n[fix_axis] * s[fix_axis] + a * s[ax[0]] + b * s[ax[1]]).ravel()
        q = np.stack([b0, b0 + s[ax[0]], b0 + s[ax[1]], b0 + s[ax[0]] + s[ax[1]]], axis=1)
        if ctype == 'hex':
            return q
        # the Kuhn split cuts every cube face along the diagonal through its lowest and highest node
        return np.concatenate([q[:, [0, 1, 3]], q[:, [0, 2, 3]]], axis=0)

    parts = [face(0, 0), face(1, 0), face(0, 1), face(1, 1), face(2, 0), face(2, 1)]
    bnd = np.vstack(parts)
    bidx = np.concatenate([np.full(len(p), r) for r, p in enumerate(parts)]).astype(np.int32)
    info = dict(N=(nx, ny, nz), scale=tuple(scale), offset=tuple(offset), cell=cell)
    return Mesh(3, ctype, pts, cells, bnd, bidx, ['back', 'left', 'front', 'right', 'bottom', 'top'],
                structured=info)


def delaunay_rectangle(n: int, seed: int = 0, size: Sequence[float] = (1.0, 1.0),
                       names: Sequence[str] = ('bottom', 'right', 'top', 'left')) -> Mesh:
    """Unstructured triangle mesh of [0,Lx]x[0,Ly]: Delaunay triangulation of jittered interior points plus evenly
    spaced boundary points. Used by the tests and benches in place of the reference's Netgen-generated .vol files
    (which do not travel to the GPU box)."""
    from scipy.spatial import Delaunay
    rng = np.random.default_rng(seed)
    lx, ly = float(size[0]), float(size[1])
    nx = max(2, int(round(n * lx / max(lx, ly))))
    ny = max(2, int(round(n * ly / max(lx, ly))))
    gx, gy = np.meshgrid(np.arange(1, nx) / nx, np.arange(1, ny) / ny)
    inner = np.stack([gx.ravel(), gy.ravel()], 1)
    inner += rng.uniform(-0.3, 0.3, inner.shape) / np.array([nx, ny])
    tx, ty = np.arange(nx + 1) / nx, np.arange(1, ny) / ny
    bnd = np.concatenate([np.stack([tx, 0 * tx], 1), np.stack([tx, 0 * tx + 1], 1),
                          np.stack([0 * ty, ty], 1), np.stack([0 * ty + 1, ty], 1)])
    pts = np.concatenate([bnd, inner]) * np.array([lx, ly])
    tri = Delaunay(pts).simplices
    P = pts[tri]
    e1, e2 = P[:, 1] - P[:, 0], P[:, 2] - P[:, 0]
    area = 0.5 * np.abs(e1[:, 0] * e2[:, 1] - e1[:, 1] * e2[:, 0])
    tri = tri[area > 1e-12 * lx * ly]
    tmp = Mesh(2, 'tri', pts, tri, np.zeros((0, 2), np.int64), np.zeros(0, np.int32), list(names))
    bf = tmp.facets[tmp.bnd_facets].astype(np.int64)
    mid = 0.5 * (pts[bf[:, 0]] + pts[bf[:, 1]])
    eps = 1e-9
    idx = np.where(mid[:, 1] < eps, 0, np.where(mid[:, 0] > lx - eps, 1, np.where(mid[:, 1] > ly - eps, 2, 3)))
    return Mesh(2, 'tri', pts, tri, bf, idx.astype(np.int32), list(names))
