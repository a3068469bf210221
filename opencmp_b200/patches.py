"""Vertex patches of the additive-Schwarz smoother (host side, built once per space).

A patch belongs to a mesh vertex v and holds a subset of the DOFs living on the cells around v (its star). Per block
of the (compound) space the subset is either

* ``closed`` — every DOF of every star cell (the patch the 2-D HDiv-DG multigrid uses), or
* ``open``   — only the DOFs interior to the star: a DOF belongs to the open star of v iff *every* cell containing it
  contains v (the vertex DOF itself, the DOFs of the edges / faces / cells touching v).

Kinds: ``vertex`` = all blocks closed, ``star`` = all blocks open, ``vanka`` = the last block (the pressure of a
Taylor-Hood pair) closed and the others open — the Vanka-type patch used for the 3-D Q2/Q1 Taylor-Hood systems, where a
closed velocity star (375 DOFs) would not fit a register-tiled patch inversion.

Stands in for the block structure NGSolve builds for ``Preconditioner(a, 'local', block=...)`` / its multigrid smoother
blocks (reference opencmp/models/base_model.py:365-383 only forwards the type string).
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

KINDS = {'vertex': 'closed', 'star': 'open', 'vanka': 'vanka'}


def vertex_patch_dofs(fes, kind: str = 'vertex', vmask=None, drop_constrained: bool = True) -> np.ndarray:
    """(npatch, bs) int32 DOF lists, ascending, padded with -1; one row per mesh vertex (``vmask``: keep only those)."""
    if kind not in KINDS:
        raise ValueError('unknown patch kind {}'.format(kind))
    m = fes.mesh
    ne, nvc = m.cells.shape
    cd = fes.cell_dofs.astype(np.int64)
    nloc = cd.shape[1]
    ndof = fes.ndof
    cell_idx = np.repeat(np.arange(ne, dtype=np.int64), nloc)
    D = sp.csr_matrix((np.ones(ne * nloc), (cell_idx, cd.ravel())), shape=(ne, ndof))
    D.sum_duplicates()
    D.data[:] = 1.0
    V = sp.csr_matrix((np.ones(ne * nvc), (np.repeat(np.arange(ne, dtype=np.int64), nvc),
                                           m.cells.astype(np.int64).ravel())), shape=(ne, m.nv))
    C = (V.T @ D).tocoo()                                   # C[v, dof] = number of star cells of v containing dof
    mult = np.asarray(D.sum(axis=0)).ravel()                # number of cells containing each dof
    # which blocks are closed
    nblk = len(fes.blocks)
    closed_blk = {'closed': [True] * nblk, 'open': [False] * nblk,
                  'vanka': [False] * (nblk - 1) + [True]}[KINDS[kind]]
    closed = np.zeros(ndof, dtype=bool)
    off = 0
    for blk, cl in zip(fes.blocks, closed_blk):
        closed[off:off + blk.ndof] = cl
        off += blk.ndof
    keep = closed[C.col] | (C.data >= mult[C.col] - 0.5)
    if drop_constrained:
        # constrained DOFs only carry identity rows in a patch matrix and their correction is masked afterwards
        keep &= np.asarray(fes.FreeDofs(), dtype=bool)[C.col]
    v, d = C.row[keep], C.col[keep]
    if vmask is not None:
        sel = np.asarray(vmask, dtype=bool)[v]
        v, d = v[sel], d[sel]
    order = np.lexsort((d, v))
    v, d = v[order], d[order]
    verts = np.arange(m.nv) if vmask is None else np.nonzero(np.asarray(vmask, dtype=bool))[0]
    start = np.searchsorted(v, verts)
    stop = np.searchsorted(v, verts, side='right')
    cnt = stop - start
    bs = int(cnt.max()) if len(cnt) else 0
    out = -np.ones((len(verts), bs), dtype=np.int32)
    pos = np.arange(len(v)) - np.repeat(start, cnt)
    rows = np.repeat(np.arange(len(verts)), cnt)
    out[rows, pos] = d
    return out
