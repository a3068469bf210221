"""Element-partitioned geometric multigrid + GMRES on top of dist.py.

Every level of the ``Mesh.Refine()`` hierarchy is partitioned with the same (nested) contiguous cell blocks; a rank
holds, per level, a local mesh (owned cells + two vertex layers), the local operator assembled by the ordinary
single-GPU kernels, the vertex-star patches of the vertices it owns, and the local prolongation. The coarsest level is
replicated (every cell local) and solved redundantly with an explicit inverse. All vectors are local (owned + ghost)
and kept *consistent*: ghost entries equal the owner's value.

Communication per V-cycle and level: one halo exchange after each SpMV, ``reverse_add`` + exchange after each smoother
application (patch corrections reach ghost DOFs) and after the restriction. Per GMRES iteration: the Gram-Schmidt
coefficients are all-reduced as one small vector.

On the GPU the whole GMRES + V-cycle runs inside the C ABI (ocmp_krylov with the element-partitioned fields of
ocmp_system / ocmp_mg_level): exchanges and all-reduces are issued from the C driver on the kernels' stream. The Python
cycle below is the same algorithm written with array expressions; it is what the gloo tests run on CPU and the
reference the native driver is checked against (OCMP_DIST_NATIVE=0 selects it on the GPU).
"""
from __future__ import annotations

import os
from typing import List, Optional

import numpy as np

from .dist import DofMap, Partition
from .multigrid import (SmootherLag, clone_space, coarse_mesh_size_scales, coarse_state_key, coefficient_fields,
                        mesh_levels, prolongation, restrict_field, reuse_coarse_enabled)


def _allreduce(vals, like):
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return np.asarray(vals, dtype=np.float64)
    dev = like.device if isinstance(like, torch.Tensor) else 'cpu'
    t = torch.as_tensor(np.asarray(vals, dtype=np.float64), device=dev)
    dist.all_reduce(t)
    return t.cpu().numpy()


def _broadcast(x):
    import torch
    import torch.distributed as dist
    if dist.is_initialized() and dist.get_world_size() > 1:
        t = x if isinstance(x, torch.Tensor) else torch.from_numpy(x)
        dist.broadcast(t, src=0)


class _Level:
    pass


def owned_node_groups(owned: np.ndarray, shift: int, nc: int) -> int:
    """For the grouped SpMV over a rank's own rows (ocmp_system.n_spmv_groups): the ascending list of owned rows must
    hold the owned rows below ``shift``, then their nc - 1 copies shifted by ``shift`` (the other components of the same
    nodes), then the remaining rows. Returns the number of such nodes, or 0 when the components of some node are not
    owned together."""
    owned = np.asarray(owned, dtype=bool)
    first = owned[:shift]
    for c in range(1, nc):
        if not np.array_equal(owned[c * shift:(c + 1) * shift], first):
            return 0
    return int(first.sum())


class DistributedMultigrid:
    """Levels [0 .. L]; level L is the space of the bilinear form ``bf`` (built on the rank's local finest mesh)."""

    def __init__(self, be, bf, global_fine_mesh, part_fine: Partition, nu: int = 1, omega: float = 0.7,
                 replicate_below: int = 100000):
        from . import ngs
        from .symbolic import lower_form
        self.be, self.bf, self.nu, self.omega = be, bf, nu, omega
        world, rank = part_fine.nranks, part_fine.rank
        gmeshes = mesh_levels(global_fine_mesh)
        L = len(gmeshes) - 1
        self.levels: List[_Level] = []
        cfields = coefficient_fields(bf)
        for l, gm in enumerate(gmeshes):
            lv = _Level()
            gfes = clone_space(bf.space, ngs.Mesh(gm))                       # global numbering only
            # small levels are replicated: every cell local, all work redundant, no communication inside the level
            lv.replicated = l < L and (l == 0 or gfes.ndof < replicate_below)
            if l == L:
                lv.part = part_fine
                lv.mesh = bf.space.mesh
                lv.fes = bf.space
            else:
                lv.part = Partition(gm, world, rank, layers=None if lv.replicated else part_fine._layers,
                                    rank_of_cells=part_fine.rank_of_cells)
                lv.mesh = ngs.Mesh(lv.part.local_mesh())
                lv.fes = clone_space(bf.space, lv.mesh)
            lv.map = DofMap(lv.part, gfes, lv.fes)
            lv.free = be.from_numpy(lv.fes.FreeDofs().astype(np.float64))
            lv.owned = be.from_numpy(np.ones(lv.fes.ndof) if lv.replicated else lv.map.owned.astype(np.float64))
            lv.n = lv.fes.ndof
            if l < L:
                lv.mat = ngs.Matrix(lv.fes)
            if l > 0:
                vown = None if lv.replicated else lv.part.vertex_owner[lv.mesh.global_vertices] == rank
                lv.patches = be.patch_state(lv.fes, vown)
                cnt = be.patch_count(lv.patches, lv.n)
                if not lv.replicated:
                    lv.map.exchange_sum(cnt)
                lv.wgt = lv.free / (cnt + (cnt == 0))
                lv.pw = 1.0 / (cnt + (cnt == 0))
                prev = self.levels[l - 1]
                parent_global = lv.part.local_cells // (2 ** lv.mesh.dim)
                parent = np.searchsorted(prev.part.local_cells, parent_global)
                assert (prev.part.local_cells[parent] == parent_global).all(), 'parent of a local cell is not local'
                P = prolongation(prev.fes, lv.fes, parent)
                lv.P = be.csr_handle(P)
                # restriction = transpose of the prolongation restricted to the fine entries this rank OWNS: every
                # fine entry is then counted by exactly one rank (the partial sums are added up by restrict_sum), and
                # the product never reads the ghost entries of the residual — which the owned-row products leave stale
                R = P.T.tocsr()
                if not lv.replicated:
                    import scipy.sparse as sp
                    R = (R @ sp.diags(lv.map.owned.astype(np.float64))).tocsr()
                    R.eliminate_zeros()
                lv.R = be.csr_handle(R)
            self.levels.append(lv)
        # coefficient fields (DIM phase field, masks) get a stand-in on every coarse level, finest to coarsest
        cur = {id(gf): gf for gf in cfields}
        hscale = coarse_mesh_size_scales(L + 1)
        for l in range(L - 1, -1, -1):
            lv, up = self.levels[l], self.levels[l + 1]
            parent_global = up.part.local_cells // (2 ** up.mesh.dim)
            parent = np.searchsorted(lv.part.local_cells, parent_global)
            lv.field_map = {k: restrict_field(g, lv.mesh, parent) for k, g in cur.items()}
            cur = dict(lv.field_map)
            fmap = {id(gf): lv.field_map[id(gf)] for gf in cfields}
            lv.program = lower_form(lv.fes, bf.integrals, 2, drop_fields=True, field_map=fmap,
                                    cell_mesh_size_scale=hscale[l])
        self.inv0 = None
        self._native = None
        self._work = None
        self._coarse_key = None
        self._fresh_coarse = True
        self.coarse_setups = 0
        self.lag = SmootherLag()

    # ---- set-up after every assembly ---------------------------------------------------------------------------
    def update(self):
        be = self.be
        # coarse levels only depend on the Parameters their programs read (multigrid.coarse_state_key): while those are
        # unchanged, only the finest level (the operator that actually changed) is set up again
        key = coarse_state_key([lv.program for lv in self.levels[:-1]])
        reuse = reuse_coarse_enabled() and key == self._coarse_key and self.inv0 is not None
        for l, lv in enumerate(self.levels):
            if l < len(self.levels) - 1:
                if reuse:
                    continue
                be.assemble_matrix(lv.program, lv.mat)
                if lv.replicated:
                    # redundant work must be bitwise identical on every rank (the regularised pressure mode amplifies
                    # round-off differences of the inverses). The two-phase assembly is bit-reproducible for identical
                    # inputs, but the coarse stand-ins of the coefficient fields are clipped to the range of each
                    # rank's LOCAL fine vector (multigrid.restrict_field), so the inputs differ in the last bits:
                    # all ranks take rank 0's coarse matrix values (set-up only, a few MB)
                    _broadcast(lv.mat.values)
            else:
                lv.mat = self.bf.mat
            if l == 0:
                self.inv0 = be.dense_inverse(lv.mat, lv.free)
            elif l < len(self.levels) - 1 or \
                    self.lag.need_refresh(lv.patches.get('inv') is not None, forced=not reuse):
                be.patch_setup(lv.mat, lv.patches, lv.free)
        if not reuse:
            self._coarse_key = key
            self.coarse_setups += 1
        self._fresh_coarse = not reuse
        self._native = None
        if getattr(be, 'name', '') == 'cuda' and os.environ.get('OCMP_DIST_NATIVE', '1') != '0':
            self._native = self._native_levels()
        return self

    def _native_levels(self):
        """Level array for the C ABI driver (ocmp_krylov, pre_kind 3) with the element-partitioned fields filled in:
        the whole GMRES + V-cycle then runs inside one C call per solve — halo exchanges (ocmp_halo_run) and the
        all-reduces of the Gram-Schmidt coefficients are issued from there, on the same stream as the kernels."""
        import ctypes as C
        from .backend import MGLevel, System, STORAGE_ID
        be = self.be
        dist_on = be.comm_init()
        nl = len(self.levels)
        arr = (MGLevel * nl)()
        for l, lv in enumerate(self.levels):
            s = arr[l].sys
            pd = be.pattern_data(lv.fes)
            s.nrows = lv.n
            s.rowptr, s.colidx, s.vals = pd['rowptr'].data_ptr(), pd['colidx'].data_ptr(), lv.mat.values.data_ptr()
            s.freemask = lv.free.data_ptr()
            distributed = dist_on and not lv.replicated
            s.owned = lv.owned.data_ptr() if distributed else None
            if distributed:
                if not hasattr(lv, 'own_rows'):
                    lv.own_rows = be._up(np.nonzero(lv.map.owned)[0].astype(np.int32))
                s.spmv_rows, s.n_spmv_rows = lv.own_rows.data_ptr(), int(lv.own_rows.numel())
                if pd.get('runs') is not None and pd['runs'][3]:
                    s.n_spmv_groups = owned_node_groups(lv.map.owned, pd['runs'][1], pd['runs'][2])
            if pd.get('runs') is not None:
                s.run_len, s.run_shift, s.run_nc = pd['runs'][0].data_ptr(), pd['runs'][1], pd['runs'][2]
                s.run_grouped = pd['runs'][1] if pd['runs'][3] else 0
            if not hasattr(lv, 'work'):
                lv.work = be.zeros(4 * lv.n)
            arr[l].work = lv.work.data_ptr()
            arr[l].nu, arr[l].omega = self.nu, self.omega
            if l == 0:
                s.pre_kind = 4
                s.inv_rowptr, s.inv_colidx = self.inv0['rowptr'].data_ptr(), self.inv0['colidx'].data_ptr()
                s.inv_vals = self.inv0['vals'].data_ptr()
                s.halo_fwd = lv.map.plan_handle(be, 'fwd') + 1 if dist_on else 0
                continue
            pt = lv.patches
            s.pre_kind = 2
            s.npatch, s.bs = pt['npatch'], pt['bs']
            s.patch_dofs, s.inv_blocks = pt['dofs'].data_ptr(), pt['inv'].data_ptr()
            s.inv_storage = STORAGE_ID[pt.get('storage', 'fp64')]
            s.patch_inc_ptr, s.patch_inc_idx = pt['inc_ptr'].data_ptr(), pt['inc_idx'].data_ptr()
            s.patch_ybuf = pt['ybuf'].data_ptr()
            if os.environ.get('OCMP_SPMV_FP32', '0') == '1':
                if self._fresh_coarse or l == nl - 1 or getattr(lv, 'vals32', None) is None:
                    lv.vals32 = be.fp32_copy(lv.mat.values, getattr(lv, 'vals32', None))
                s.vals32 = lv.vals32.data_ptr()
            s.patch_weight = lv.pw.data_ptr()
            prev = self.levels[l - 1]
            arr[l].ncoarse = prev.n
            arr[l].p_rowptr, arr[l].p_colidx, arr[l].p_vals = (lv.P[k].data_ptr() for k in ('rowptr', 'colidx', 'vals'))
            arr[l].r_rowptr, arr[l].r_colidx, arr[l].r_vals = (lv.R[k].data_ptr() for k in ('rowptr', 'colidx', 'vals'))
            if distributed:
                s.halo_fwd = lv.map.plan_handle(be, 'fwd') + 1
                s.halo_sum = lv.map.plan_handle(be, 'sum') + 1
                arr[l].restrict_sum = prev.map.plan_handle(be, 'sum') + 1
                if prev.replicated:
                    arr[l].handover = prev.map.plan_handle(be, 'fwd') + 1
        top = System.from_buffer_copy(arr[nl - 1].sys)
        top.pre_kind = 3
        top.nlevels = nl
        top.levels = C.addressof(arr)
        return top, arr

    def _gmres_native(self, b, x, tol, maxit, restart):
        import ctypes as C
        be = self.be
        top, _ = self._native
        wl = be.lib.ocmp_krylov_work_len(top.nrows, 1, restart)
        if self._work is None or self._work.numel() < wl:
            self._work = be.torch.empty(wl, dtype=be.torch.float64, device=be.device)
        it, res = C.c_int(0), C.c_double(0.0)
        be._ck(be.lib.ocmp_krylov(C.byref(top), 1, b.data_ptr(), x.data_ptr(), float(tol), int(maxit), int(restart),
                                  1.0, self._work.data_ptr(), wl, C.byref(it), C.byref(res), be._stream()))
        return it.value, res.value

    # ---- level operators (all vectors consistent) ----------------------------------------------------------------
    def mult(self, l, x, owned_only: bool = False):
        """y = A x with consistent ghosts; ``owned_only`` skips the halo exchange when only owned entries are used
        afterwards (the residual that is restricted)."""
        lv = self.levels[l]
        y = self.be.zeros(lv.n)
        self.be.spmv(lv.mat, x, y)
        if not lv.replicated and not owned_only:
            lv.map.exchange(y)
        return y

    def smooth(self, l, r):
        lv = self.levels[l]
        z = self.be.zeros(lv.n)
        self.be.patch_apply(lv.patches, r, z)
        if not lv.replicated:
            lv.map.exchange_sum(z)
        return z * lv.wgt

    def vcycle(self, l, b):
        lv = self.levels[l]
        be = self.be
        if l == 0:
            x = be.zeros(lv.n)
            be.csr_mult(self.inv0, b, x)
            # every rank solved redundantly with ITS OWN copy of the (atomically assembled, hence round-off-different)
            # coarse matrix; the regularised pressure mode amplifies such differences, so take the owners' values
            lv.map.exchange(x)
            return x * lv.free
        x = self.omega * self.smooth(l, b)
        for _ in range(1, self.nu):
            x = x + self.omega * self.smooth(l, (b - self.mult(l, x)) * lv.free)
        r = (b - self.mult(l, x, owned_only=True)) * lv.free
        prev = self.levels[l - 1]
        bc = be.zeros(prev.n)
        be.csr_mult(lv.R, r * lv.owned, bc)
        if not lv.replicated:
            prev.map.exchange_sum(bc)
        xc = self.vcycle(l - 1, bc * prev.free)
        if prev.replicated and not lv.replicated:
            prev.map.exchange(xc)          # redundant coarse work differs by (amplified) round-off: owners' values win
        t = be.zeros(lv.n)
        be.csr_mult(lv.P, xc, t)
        x = x + t * lv.free
        for _ in range(self.nu):
            x = x + self.omega * self.smooth(l, (b - self.mult(l, x)) * lv.free)
        return x

    def precond(self, r):
        return self.vcycle(len(self.levels) - 1, r)

    # ---- Krylov ------------------------------------------------------------------------------------------------
    def dots(self, V, k, w):
        """<V_j, w> over owned entries, j < k, one all-reduce."""
        lv = self.levels[-1]
        local = V[:k] @ (w * lv.owned)
        loc = local.cpu().numpy() if hasattr(local, 'cpu') else np.asarray(local)
        return _allreduce(loc, w)

    def gmres(self, b, x, tol=1e-10, maxit=300, restart=50):
        """Left-preconditioned restarted GMRES with CGS2 on the free DOFs; x (consistent) holds the initial guess and
        the Dirichlet values. Same recurrence as ocmp_krylov kind 1."""
        if self._native is not None:
            it, res = self._gmres_native(b, x, tol, maxit, restart)
        else:
            it, res = self._gmres_python(b, x, tol, maxit, restart)
        self.lag.note_solve(it)
        return it, res

    def _gmres_python(self, b, x, tol, maxit, restart):
        be = self.be
        top = len(self.levels) - 1
        lv = self.levels[top]
        n = lv.n
        V = be.zeros((restart + 1) * n).reshape(restart + 1, n)
        self.history = []                  # relative preconditioned residual per iteration (diagnostics)
        it, beta0, res = 0, None, 0.0
        done = False
        while not done and it < maxit:
            r = (b - self.mult(top, x)) * lv.free
            z = self.precond(r)
            beta = float(np.sqrt(self.dots(z.reshape(1, n), 1, z)[0]))
            if beta0 is None:
                beta0 = beta
            res = beta
            if beta == 0.0 or beta < tol * beta0:
                break
            V[0] = z / beta
            H = np.zeros((restart + 1, restart))
            cs, sn, g = np.zeros(restart), np.zeros(restart), np.zeros(restart + 1)
            g[0] = beta
            k = 0
            while k < restart and it < maxit:
                w = self.precond(self.mult(top, V[k]) * lv.free)
                h = self.dots(V, k + 1, w)
                w = w - self._lincomb(V, k + 1, h)
                h2 = self.dots(V, k + 1, w)
                w = w - self._lincomb(V, k + 1, h2)
                h = h + h2
                hn = float(np.sqrt(self.dots(w.reshape(1, n), 1, w)[0]))
                H[:k + 1, k] = h
                H[k + 1, k] = hn
                for i in range(k):
                    a_, b_ = H[i, k], H[i + 1, k]
                    H[i, k], H[i + 1, k] = cs[i] * a_ + sn[i] * b_, -sn[i] * a_ + cs[i] * b_
                den = np.hypot(H[k, k], H[k + 1, k])
                cs[k], sn[k] = (1.0, 0.0) if den == 0 else (H[k, k] / den, H[k + 1, k] / den)
                H[k, k] = cs[k] * H[k, k] + sn[k] * H[k + 1, k]
                H[k + 1, k] = 0.0
                g[k + 1] = -sn[k] * g[k]
                g[k] = cs[k] * g[k]
                it += 1
                res = abs(g[k + 1])
                self.history.append(res / beta0)
                if hn > 0:
                    V[k + 1] = w / hn
                k += 1
                if res < tol * beta0 or hn == 0:
                    done = True
                    break
            y = np.linalg.solve(np.triu(H[:k, :k]), g[:k]) if k else np.zeros(0)
            x += self._lincomb(V, k, y)
        return it, res

    def _lincomb(self, V, k, coef):
        c = self.be.from_numpy(np.asarray(coef, dtype=np.float64))
        return c @ V[:k]
