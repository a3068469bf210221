"""Geometric multigrid preconditioner on the uniform-refinement hierarchy kept by ``Mesh.Refine()``.

Stands in for ``ngs.Preconditioner(a, 'multigrid')`` (reference opencmp/models/base_model.py:365-383 lists the types
OpenCMP forwards to NGSolve: local, direct, multigrid, h1amg, bddc). The spaces of a uniformly refined affine mesh are
nested, so the prolongation of every block (H1 / L2: plain pull-back, HDiv: contravariant Piola pull-back) is an exact
embedding; it is tabulated once per *child class* (the affine map fine reference cell -> coarse reference cell takes
only a handful of values under red refinement) and scattered into one CSR matrix on the host.

Levels below the finest re-discretise the field-independent part of the bilinear form (convection by the Oseen wind
and other DOF-vector weighted terms only act on the finest level, through the smoother and the Krylov method around
the cycle). The cycle itself runs inside the C ABI (ocmp_krylov with pre_kind = 3): vertex-patch additive Schwarz
smoothing, CSR restriction / prolongation, explicit inverse on the coarsest level.
"""
from __future__ import annotations

from typing import List

import numpy as np
import scipy.sparse as sp

from .quadrature import cell_rule
from .space import FESpace


def mesh_levels(mesh) -> list:
    """[coarsest, ..., finest] as kept by Mesh.Refine()."""
    out = [mesh]
    while getattr(out[-1], 'coarse', None) is not None:
        out.append(out[-1].coarse)
    return out[::-1]


def clone_space(fes: FESpace, mesh) -> FESpace:
    """Same space family / order / Dirichlet names on another mesh."""
    cls = type(fes)
    if fes.components:
        return cls([clone_space(c, mesh) for c in fes.components], dgjumps=fes.dgjumps)
    b = fes.blocks[0]
    return cls(mesh, order=fes.order, dirichlet=b.dirichlet or '', dgjumps=fes.dgjumps, family=fes.name,
               RT=getattr(fes, 'RT', False))


def prolongation(fes_c: FESpace, fes_f: FESpace, parent=None) -> sp.csr_matrix:
    """(ndof_f x ndof_c) embedding of the coarse space into the fine space. ``parent[k]`` is the coarse cell containing
    fine cell k (default: uniform refinement kept by Mesh.Refine(), parent k // 4; the element-partitioned layer passes
    the parent map of its local sub-meshes)."""
    mc, mf = fes_c.mesh, fes_f.mesh
    if parent is None:
        nch = 2 ** mf.dim                      # children per cell: red refinement of triangles, 2^d for quads / hexes
        if mf.ne != nch * mc.ne:
            raise ValueError('prolongation needs a uniformly refined mesh')
        parent = np.arange(mf.ne) // nch
    parent = np.asarray(parent, dtype=np.int64)
    dim = mf.dim
    Jc, Jf = mc.jacobians(), mf.jacobians()
    Jci = np.linalg.inv(Jc)[parent]
    A = np.einsum('eab,ebc->eac', Jci, Jf)                              # xi_c = A xi_f + b
    b = np.einsum('eab,eb->ea', Jci, mf.origins() - mc.origins()[parent])
    key = np.round(np.concatenate([A.reshape(mf.ne, -1), b], axis=1), 9)
    classes, cls_of = np.unique(key, axis=0, return_inverse=True)
    cls_of = cls_of.reshape(-1)
    rows, cols, vals = [], [], []
    cdf, cdc = fes_f.cell_dofs.astype(np.int64), fes_c.cell_dofs.astype(np.int64)
    for blk_f, blk_c, lo_f, lo_c in zip(fes_f.blocks, fes_c.blocks, fes_f.loc_offsets, fes_c.loc_offsets):
        bf, bc = blk_f.basis, blk_c.basis
        deg = 2 * bf.order + 1
        pts, w = cell_rule(mf.cell_type, deg)
        tf = bf.tabulate(pts)                                           # (nq, nrows, nloc_f)
        nv = 1 if bf.kind == 'scalar' else dim
        phi_f = tf[:, :nv, :]
        M = np.einsum('q,qci,qcj->ij', w, phi_f, phi_f)
        Minv = np.linalg.inv(M)
        for k, row in enumerate(classes):
            Ak = row[:dim * dim].reshape(dim, dim)
            bk = row[dim * dim:]
            tc = bc.tabulate(pts @ Ak.T + bk)[:, :nv, :]                # coarse functions at the mapped points
            if bf.kind != 'scalar':
                tc = np.linalg.det(Ak) * np.einsum('ab,qbj->qaj', np.linalg.inv(Ak), tc)
            Ploc = Minv @ np.einsum('q,qci,qcj->ij', w, phi_f, tc)      # (nloc_f, nloc_c)
            cells = np.nonzero(cls_of == k)[0]
            r = cdf[cells][:, lo_f:lo_f + blk_f.nloc]
            c = cdc[parent[cells]][:, lo_c:lo_c + blk_c.nloc]
            nz = np.abs(Ploc) > 1e-13
            ii, jj = np.nonzero(nz)
            rows.append(r[:, ii].ravel())
            cols.append(c[:, jj].ravel())
            vals.append(np.broadcast_to(Ploc[ii, jj], (len(cells), len(ii))).ravel())
    rows, cols, vals = np.concatenate(rows), np.concatenate(cols), np.concatenate(vals)
    key = rows * np.int64(fes_c.ndof) + cols
    _, first = np.unique(key, return_index=True)
    P = sp.csr_matrix((vals[first], (rows[first], cols[first])), shape=(fes_f.ndof, fes_c.ndof))
    P.sort_indices()
    return P


def coefficient_fields(bf) -> list:
    """GridFunctions the bilinear form is weighted with that do NOT live on (a component of) its own space: coefficient
    fields such as the DIM phase field / masks (reference diffuse_interface/dim.py:390-445), which every level of the
    hierarchy has to see. Fields on the form's own spaces (Oseen wind, previous iterates) are linearisation data and
    only act on the finest level. ``gf.mg_static`` overrides the rule."""
    from .ir import coef_leaves
    fes = bf.space
    own = [fes] + list(fes.components)
    out, seen = [], set()
    for cf, _ in bf.integrals.items:
        s = cf.arr.reshape(())[()]
        for lf in coef_leaves(list(s.t.values()), 'field'):
            gf = lf.val[0]
            if id(gf) in seen:
                continue
            seen.add(id(gf))
            static = getattr(gf, 'mg_static', None)
            if static is None:
                static = not any(gf.space is o for o in own)
            if static:
                out.append(gf)
    return out


def restrict_field(gf, mesh_c, parent=None):
    """Coarse-level stand-in of a coefficient field: re-evaluated from the expression it was ``Set`` from when that is
    known, otherwise the least-squares restriction min ||P c - f|| through the exact prolongation of its space. Clipped
    to the range of the fine DOF vector (mirrors the reference's clamp of phi, dim.py:434-435)."""
    from . import ngs
    import scipy.sparse.linalg as spla
    be = ngs.get_backend()
    fes_c = clone_space(gf.space, mesh_c)
    out = ngs.GridFunction(fes_c)
    fine = be.to_numpy(gf.vec.a)
    src = getattr(gf, '_set_source', None)
    if src is not None:
        out.Set(src)
        vals = be.to_numpy(out.vec.a)
    else:
        P = prolongation(fes_c, gf.space, parent)
        vals, _ = spla.cg(P.T @ P, P.T @ fine, rtol=1e-12, maxiter=2000)
    if fine.size:
        vals = np.clip(vals, fine.min(), fine.max())
    out.vec.data = ngs.BaseVector(be.from_numpy(vals))
    out._set_source = src
    return out


def coarse_state_key(programs) -> tuple:
    """Everything the re-discretised coarse-level operators depend on at run time: the values of the Parameters their
    coefficient programs read (dt, t, ...). Their coefficient *fields* are the static stand-ins built once
    (``restrict_field``); linearisation data (Oseen wind, previous iterates) only act on the finest level. While the key
    is unchanged the coarse matrices, their patch inverses and the coarsest-level inverse are still valid — NGSolve's
    own multigrid likewise keeps the coarse-level matrices it assembled earlier. ``OCMP_MG_REUSE_COARSE=0`` forces the
    full set-up on every ``Update()``."""
    return tuple(tuple(prog.param_values(integ).tolist()) for prog in programs for integ in prog.integrals)


class SmootherLag:
    """Opt-in (``OCMP_MG_LAG=1``) lagging of the FINEST level's patch inverses: they are a preconditioner, and between
    Picard iterations / time steps the operator only changes through the Oseen wind, so the inverses of an earlier
    assembly usually smooth just as well (2-D INS workload on the CPU restatement: 18 / 15 iterations per step with
    inverses never refreshed over four steps, exactly as with fresh ones). The GMRES iteration counts steer it: a
    solve that needs more than ``1.25 x + 2`` iterations of the first solve after the last fresh set-up triggers a
    fresh set-up at the next ``Update()``. Off by default: ``Preconditioner.Update()`` then re-inverts every time, like
    the library it stands in for."""

    def __init__(self):
        import os
        self.enabled = os.environ.get('OCMP_MG_LAG', '0') == '1'
        self.base = None            # iterations of the first solve after the last fresh set-up
        self.last = None            # iterations of the most recent solve
        self._fresh = False
        self.fresh_setups = 0

    def need_refresh(self, have_inverses: bool, forced: bool = False) -> bool:
        stale = self.base is not None and self.last is not None and self.last > 1.25 * self.base + 2
        fresh = forced or not self.enabled or not have_inverses or stale
        if fresh:
            self._fresh = True
            self.base = self.last = None
            self.fresh_setups += 1
        return fresh

    def note_solve(self, iterations: int) -> None:
        if self._fresh:
            self.base = int(iterations)
            self._fresh = False
        self.last = int(iterations)


def inherit_cell_penalty() -> bool:
    """Coarse-level operators evaluate ``specialcf.mesh_size`` inside cell integrals with the FINE mesh size (default;
    ``OCMP_MG_COARSE_H=own`` restores the coarse level's own size). The reference's diffuse-interface forms carry the
    volume penalty alpha = ipc k^2 / h (models/ins_dim.py:60-104 through base_model.py:150): outside the fluid the
    momentum rows reduce to alpha (1 - phi) u + grad p, i.e. the pressure there obeys a Darcy problem with
    permeability 1 / alpha ~ h. A coarse operator re-discretised with its own h doubles that permeability per level,
    so the coarse-grid correction of the smooth pressure modes of the solid region is too small by 2, 4, 8, ... —
    every additional level adds modes the cycle barely reduces, and the GMRES iteration count grows with the mesh
    (26 / 46 iterations at 8^3 / 16^3 hexes, 63 at 48^3, 143-325 at 96^3). With the fine level's penalty on every level
    the count is 21-24 / 26-27 at 8^3 / 16^3 (profiles/r2_mg_penalty_study.md). Interior-penalty terms on facets
    (DG forms) are left alone: there the level's own h is the right scale."""
    import os
    return os.environ.get('OCMP_MG_COARSE_H', 'fine') != 'own'


def coarse_mesh_size_scales(nlevels: int) -> list:
    """Factor on the cell mesh size of the coarse levels 0 .. nlevels - 2 (uniform refinement halves h per level)."""
    if not inherit_cell_penalty():
        return [1.0] * (nlevels - 1)
    return [0.5 ** (nlevels - 1 - l) for l in range(nlevels - 1)]


def reuse_coarse_enabled() -> bool:
    import os
    return os.environ.get('OCMP_MG_REUSE_COARSE', '1') != '0'


# ---- device side: level hierarchy handed to the C ABI (pre_kind = 3) ---------------------------------------------
class MultigridState:
    """Built by ``CudaBackend.precond_setup(..., 'multigrid')``; refreshed on every ``Preconditioner.Update()``."""
    kind = 3

    def __init__(self, be, bf, nu: int = 2, omega: float = 1.0):
        import ctypes as C
        from .backend import MGLevel
        from .symbolic import lower_form
        from . import ngs
        self.be, self.bf = be, bf
        fes = bf.space
        meshes = mesh_levels(fes.mesh)
        self.nlevels = len(meshes)
        if self.nlevels < 2:
            raise ValueError("Preconditioner type 'multigrid' needs a mesh refined with Mesh.Refine()")
        self.spaces = [clone_space(fes, m) for m in meshes[:-1]] + [fes]
        # coefficient fields (DIM phase field, masks) get a stand-in on every coarse level, finest to coarsest
        self.field_maps = [dict() for _ in meshes[:-1]]
        self._coarse_fields = []
        for gf in coefficient_fields(bf):
            cur = gf
            for l in range(self.nlevels - 2, -1, -1):
                cur = restrict_field(cur, self.spaces[l].mesh)
                self.field_maps[l][id(gf)] = cur
                self._coarse_fields.append(cur)
        self.programs = [lower_form(s, bf.integrals, 2, drop_fields=True, field_map=fm, cell_mesh_size_scale=hs)
                         for s, fm, hs in zip(self.spaces[:-1], self.field_maps,
                                              coarse_mesh_size_scales(self.nlevels))]
        self.mats = [ngs.Matrix(s) for s in self.spaces[:-1]]
        self.transfers = []
        for lo, hi in zip(self.spaces[:-1], self.spaces[1:]):
            P = prolongation(lo, hi)
            R = P.T.tocsr()
            R.sort_indices()
            self.transfers.append(tuple(be._up(a) for a in (P.indptr.astype(np.int32), P.indices.astype(np.int32),
                                                            P.data, R.indptr.astype(np.int32),
                                                            R.indices.astype(np.int32), R.data)))
        self.work = [be.zeros(4 * s.ndof) for s in self.spaces]
        self.masks = [be._up(s.FreeDofs().astype(np.float64)) for s in self.spaces]
        n0 = self.spaces[0].ndof
        pat0 = self.spaces[0].pattern()
        t = be.torch
        self.rows0 = be._up(np.repeat(np.arange(n0), np.diff(pat0.rowptr)).astype(np.int64))
        self.cols0 = be._up(pat0.colidx.astype(np.int64))
        self.inv_rowptr = be._up((np.arange(n0 + 1, dtype=np.int64) * n0).astype(np.int32))
        self.inv_colidx = be._up(np.tile(np.arange(n0, dtype=np.int32), n0))
        self.inv_vals = None
        self.levels = (MGLevel * self.nlevels)()
        self.nu, self.omega = nu, omega
        self.fm = self.masks[-1]
        self.smoothers = [None] * self.nlevels
        # OCMP_SPMV_FP32=1: the cycle applies FP32-stored copies of the level matrices (the Krylov method keeps FP64)
        import os
        self.spmv_fp32 = os.environ.get('OCMP_SPMV_FP32', '0') == '1'
        self.vals32 = [None] * self.nlevels
        self._coarse_key = None
        self.coarse_setups = 0                 # how often the coarse levels were (re)built — diagnostics / tests
        self.updates = 0
        self.lag = SmootherLag()

    def update(self, fine_mat):
        be = self.be
        t = be.torch
        self.updates += 1
        key = coarse_state_key(self.programs)
        reuse = reuse_coarse_enabled() and key == self._coarse_key and self.inv_vals is not None
        if not reuse:
            for l in range(self.nlevels - 1):
                be.assemble_matrix(self.programs[l], self.mats[l])
            # coarsest level: explicit inverse of the free-free block (identity on constrained dofs)
            n0 = self.spaces[0].ndof
            dense = t.zeros((n0, n0), dtype=t.float64, device=be.device)
            dense[self.rows0, self.cols0] = self.mats[0].values
            m = self.masks[0]
            dense = dense * m[:, None] * m[None, :] + t.diag(1.0 - m)
            self.inv_vals = t.linalg.inv(dense).contiguous().view(-1)
            self._coarse_key = key
            self.coarse_setups += 1
        for l in range(self.nlevels):
            mat = fine_mat if l == self.nlevels - 1 else self.mats[l]
            lv = self.levels[l]
            if l == 0:
                sys_ = be._system(mat, self.masks[0], None)
                sys_.pre_kind = 4
                sys_.inv_rowptr, sys_.inv_colidx = self.inv_rowptr.data_ptr(), self.inv_colidx.data_ptr()
                sys_.inv_vals = self.inv_vals.data_ptr()
            else:
                if l < self.nlevels - 1:
                    fresh = not (reuse and self.smoothers[l] is not None)
                else:                       # finest level: always, unless the opt-in lag policy says otherwise
                    fresh = self.lag.need_refresh(self.smoothers[l] is not None, forced=not reuse)
                if fresh:
                    self.smoothers[l] = be.precond_setup(mat, 'asm', self.spaces[l].FreeDofs(), mask=self.masks[l],
                                                          state=self.smoothers[l])
                sm = self.smoothers[l]
                sys_ = be._system(mat, self.masks[l], sm)
                if self.spmv_fp32:
                    # the finest-level matrix changed with this assembly even when its smoother is lagged
                    if fresh or l == self.nlevels - 1 or self.vals32[l] is None:
                        self.vals32[l] = be.fp32_copy(mat.values, self.vals32[l])
                    sys_.vals32 = self.vals32[l].data_ptr()
                P = self.transfers[l - 1]
                lv.ncoarse = self.spaces[l - 1].ndof
                lv.p_rowptr, lv.p_colidx, lv.p_vals = (a.data_ptr() for a in P[:3])
                lv.r_rowptr, lv.r_colidx, lv.r_vals = (a.data_ptr() for a in P[3:])
            lv.sys = sys_
            lv.work = self.work[l].data_ptr()
            lv.nu, lv.omega = self.nu, self.omega
        return self

    def note_solve(self, iterations: int) -> None:
        """Called by ``CudaBackend.krylov`` after every solve preconditioned with this state."""
        self.lag.note_solve(iterations)
