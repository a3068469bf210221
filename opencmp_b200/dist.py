"""Element-partitioned multi-GPU layer: one process per GPU, torch.distributed for the plumbing.

The reference has no distributed code at all ("MPI = future work"); this is the B200-side design:

* the mesh is partitioned by cells into contiguous blocks of the generator's cell order (``Partition``); every rank
  builds a LOCAL mesh = its owned cells plus ``layers`` vertex-layers of ghost cells, and runs the ordinary single-GPU
  machinery (spaces, pattern, assembly kernels) on it — rows of owned DOFs are complete with one ghost layer (two for
  vertex-patch smoothers), ghost rows are never used;
* a DOF is owned by the rank owning the lowest-numbered global cell that contains it (``DofMap``); the halo plan
  (who sends which owned values to whom) is computed redundantly on every rank from the replicated global numbering,
  so no set-up communication is needed;
* per operator application ONE halo exchange of ghost values (``exchange``: batched isend/irecv over NCCL/NVLink, or
  gloo in the CPU tests), per Krylov iteration one all-reduce per dot product over the owned entries (``dot``);
  contributions computed for ghost entries (restriction, patch corrections) go back with ``reverse_add``.

On this layer run: distributed SpMV, dots and Jacobi-preconditioned CG (``DistributedOperator``), and the distributed
geometric multigrid + GMRES of dist_mg.py that bench.py --gpus N uses for the 2-D INS and 3-D INS-DIM time steps (on
the GPU driven by the C ABI, which issues the halo exchanges and all-reduces itself). tests/test_dist_gloo.py covers
all of it with world_size 2 on CPU.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import numpy as np

from .mesh import Mesh


class Partition:
    """Cell partition of a global mesh and the local (owned + ghost) mesh of one rank."""

    def __init__(self, mesh: Mesh, nranks: int, rank: int, layers: Optional[int] = 1, rank_of_cells=None):
        """``rank_of_cells(mesh) -> (ne,) ranks`` overrides the default partition into contiguous blocks of the cell
        order (e.g. ``block_ranks``: compact bricks by cell centroid); it is kept and applied to every level of a
        refinement hierarchy, so it must give a cell and its children the same rank."""
        self.mesh, self.nranks, self.rank = mesh, nranks, rank
        self._layers = layers
        self.rank_of_cells = rank_of_cells
        ne = mesh.ne
        if rank_of_cells is None:
            self.cell_rank = (np.arange(ne, dtype=np.int64) * nranks // ne).astype(np.int32)
        else:
            self.cell_rank = np.asarray(rank_of_cells(mesh), dtype=np.int32)
            assert self.cell_rank.shape == (ne,) and 0 <= self.cell_rank.min() and self.cell_rank.max() < nranks
        owned = self.cell_rank == rank
        local = owned.copy()
        if layers is None:                            # replicated level: every cell is local
            local[:] = True
        for _ in range(layers or 0):                  # grow by vertex adjacency
            vmask = np.zeros(mesh.nv, dtype=bool)
            vmask[mesh.cells[local].ravel()] = True
            local = vmask[mesh.cells].any(axis=1)
        self.local_cells = np.nonzero(local)[0]
        self.owned_local = owned[self.local_cells]    # mask over local cells
        # vertex ownership (patch smoothers): rank of the lowest global cell containing the vertex
        vowner = np.full(mesh.nv, nranks, dtype=np.int32)
        order = np.arange(ne - 1, -1, -1)
        vowner[mesh.cells[order].ravel()] = np.repeat(self.cell_rank[order], mesh.cells.shape[1])
        self.vertex_owner = vowner

    def local_mesh(self) -> Mesh:
        """Sub-mesh in ascending global cell / vertex order (keeps the sorted-vertex convention); boundary facets of
        the global mesh keep their region, cut facets get the extra region '_cut' (never a boundary condition).
        Material 'owned' / 'ghost' lets ``Integrate(..., definedon=mesh.Materials('owned'))`` count every cell once."""
        g = self.mesh
        cells = g.cells[self.local_cells].astype(np.int64)
        verts = np.unique(cells)
        v_l = -np.ones(g.nv, dtype=np.int64)
        v_l[verts] = np.arange(len(verts))
        bf = g.facets[g.bnd_facets].astype(np.int64)
        keep = (v_l[bf] >= 0).all(axis=1)
        names = list(g.bnd_names) + ['_cut']
        loc = Mesh(g.dim, g.cell_type, g.points[verts], v_l[cells], v_l[bf[keep]], g.bnd_region[keep], names,
                   np.where(self.owned_local, 0, 1).astype(np.int32), ['owned', 'ghost'])
        # facets on the local boundary that are not global boundary facets: region '_cut'
        cut = loc.bnd_names.index('_cut')
        if 'default' in loc.bnd_names:
            d = loc.bnd_names.index('default')
            loc.bnd_region[loc.bnd_region == d] = cut
        # a global boundary facet whose vertices are all local but whose cell is not local must not count: the Mesh
        # constructor only tags facets that exist locally, so nothing else to do
        loc.global_cells = self.local_cells
        loc.global_vertices = verts
        return loc


def block_ranks(blocks, lo, hi):
    """Partition of the box [lo, hi] into blocks[0] x blocks[1] (x blocks[2]) equal bricks: the rank of a cell is the
    brick its centroid lies in. Nested under uniform refinement as long as the brick faces are mesh faces of the
    coarsest partitioned level."""
    blocks = [int(b) for b in blocks]

    def rank_of_cells(mesh):
        c = mesh.points[mesh.cells].mean(axis=1)
        r = np.zeros(mesh.ne, dtype=np.int64)
        for a in range(mesh.dim):
            i = np.floor((c[:, a] - lo[a]) / (hi[a] - lo[a]) * blocks[a]).astype(np.int64)
            r = r * blocks[a] + np.clip(i, 0, blocks[a] - 1)
        return r
    return rank_of_cells


def brick_grid(world: int):
    """(bx, by, bz) with bx * by * bz = world, as cubic as powers of two allow: 1 -> (1,1,1), 2 -> (2,1,1),
    4 -> (2,2,1), 8 -> (2,2,2), 16 -> (4,2,2) ..."""
    b = [1, 1, 1]
    a = 0
    w = world
    while w > 1:
        if w % 2:
            raise ValueError('the brick layout needs a power-of-two number of ranks')
        b[a % 3] *= 2
        w //= 2
        a += 1
    return tuple(b)


class DofMap:
    """Local <-> global DOF correspondence, ownership and halo plan of one space on one rank."""

    def __init__(self, part: Partition, fes_global, fes_local):
        self.part = part
        R, me = part.nranks, part.rank
        gcd = fes_global.cell_dofs.astype(np.int64)
        lcd = fes_local.cell_dofs.astype(np.int64)
        self.nlocal, self.nglobal = fes_local.ndof, fes_global.ndof
        l2g = -np.ones(self.nlocal, dtype=np.int64)
        l2g[lcd.ravel()] = gcd[part.local_cells].ravel()
        assert (l2g >= 0).all()
        self.l2g = l2g
        # owner of every global dof = rank of the lowest global cell containing it
        owner = np.full(self.nglobal, R, dtype=np.int32)
        order = np.arange(fes_global.mesh.ne - 1, -1, -1)
        owner[gcd[order].ravel()] = np.repeat(part.cell_rank[order], gcd.shape[1])
        self.owner_local = owner[l2g]
        self.owned = self.owner_local == me
        g2l = -np.ones(self.nglobal, dtype=np.int64)
        g2l[l2g] = np.arange(self.nlocal)
        # halo plan, identical on both sides because it is a pure function of the replicated global numbering
        self.recv: Dict[int, np.ndarray] = {}
        self.send: Dict[int, np.ndarray] = {}
        self.shared: Dict[int, np.ndarray] = {}
        ghosts = np.nonzero(~self.owned)[0]
        for s in np.unique(self.owner_local[ghosts]):
            idx = ghosts[self.owner_local[ghosts] == s]
            self.recv[int(s)] = idx[np.argsort(l2g[idx])]
        for s in _candidate_peers(part):
            other = Partition(part.mesh, R, s, layers=_layers_of(part), rank_of_cells=part.rank_of_cells)
            theirs = np.unique(gcd[other.local_cells].ravel())
            mine = theirs[owner[theirs] == me]          # ascending global id = their recv order
            if mine.size:
                self.send[s] = g2l[mine]
            both = theirs[g2l[theirs] >= 0]             # dofs local to both ranks, ascending global id on both sides
            if both.size:
                self.shared[s] = g2l[both]
        self._bufs = None

    # ---- communication -------------------------------------------------------------------------------------
    def exchange(self, x, reverse_add: bool = False) -> None:
        """Refresh the ghost entries of x from their owners (or, with ``reverse_add``, add the ghost entries of x
        into the owners' entries). x is a backend array (torch tensor on the GPU, ndarray under gloo tests)."""
        import torch
        import torch.distributed as dist
        if not dist.is_initialized() or dist.get_world_size() == 1:
            return
        t = x if isinstance(x, torch.Tensor) else torch.from_numpy(x)
        if t.is_cuda and self._native('rev' if reverse_add else 'fwd', t, add=reverse_add):
            return
        out_plan, in_plan = (self.recv, self.send) if reverse_add else (self.send, self.recv)
        ops, staged = [], []
        for s, idx in sorted(out_plan.items()):
            it = self._index(idx, t.device)
            buf = t.index_select(0, it).contiguous()
            ops.append(dist.P2POp(dist.isend, buf, s))
        for s, idx in sorted(in_plan.items()):
            buf = torch.empty(len(idx), dtype=t.dtype, device=t.device)
            staged.append((idx, buf))
            ops.append(dist.P2POp(dist.irecv, buf, s))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        for idx, buf in staged:
            it = self._index(idx, t.device)
            if reverse_add:
                t.index_add_(0, it, buf)
            else:
                t.index_copy_(0, it, buf)

    def exchange_sum(self, x) -> None:
        """x holds per-rank PARTIAL contributions (patch corrections, restricted residuals): afterwards every rank holds
        the total on all its local entries — one round instead of reverse_add + exchange."""
        import torch
        import torch.distributed as dist
        if not dist.is_initialized() or dist.get_world_size() == 1:
            return
        t = x if isinstance(x, torch.Tensor) else torch.from_numpy(x)
        if t.is_cuda and self._native('sum', t, add=True):
            return
        ops, staged = [], []
        for s, idx in sorted(self.shared.items()):
            it = self._index(idx, t.device)
            ops.append(dist.P2POp(dist.isend, t.index_select(0, it).contiguous(), s))
            buf = torch.empty(len(idx), dtype=t.dtype, device=t.device)
            staged.append((it, buf))
            ops.append(dist.P2POp(dist.irecv, buf, s))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        for it, buf in staged:
            t.index_add_(0, it, buf)

    def _native(self, kind: str, t, add: bool) -> bool:
        """Halo exchange through the C ABI (ocmp_halo_run: pack + grouped NCCL send/recv + unpack in one call)."""
        from . import ngs
        be = ngs.get_backend()
        if getattr(be, 'name', '') != 'cuda' or not be.comm_init():
            return False
        be.halo_run(self.plan_handle(be, kind), t, add)
        return True

    def plan_handle(self, be, kind: str) -> int:
        """C-side halo plan ('fwd' ghost refresh, 'rev' reverse add, 'sum' partial sums) of this map; built once."""
        plans = self.__dict__.setdefault('_plans', {})
        if kind not in plans:
            out_plan, in_plan = {'fwd': (self.send, self.recv), 'rev': (self.recv, self.send),
                                 'sum': (self.shared, self.shared)}[kind]
            peers = sorted(set(out_plan) | set(in_plan))
            empty = np.zeros(0, dtype=np.int64)
            plans[kind] = be.halo_plan(peers, [out_plan.get(s, empty) for s in peers],
                                       [in_plan.get(s, empty) for s in peers])
        return plans[kind][0]

    def _index(self, idx, device):
        import torch
        key = (id(idx), str(device))
        if self._bufs is None:
            self._bufs = {}
        if key not in self._bufs:
            self._bufs[key] = torch.from_numpy(np.ascontiguousarray(idx)).to(device)
        return self._bufs[key]

    def dot(self, be, x, y, mask) -> float:
        """<x, y> over the owned entries, summed over ranks (one all-reduce)."""
        import torch
        import torch.distributed as dist
        local = be.dot(x * mask, y)
        if dist.is_initialized() and dist.get_world_size() > 1:
            t = torch.tensor([local], dtype=torch.float64, device=x.device if isinstance(x, torch.Tensor) else 'cpu')
            dist.all_reduce(t)
            local = float(t[0])
        return local


def _candidate_peers(part: Partition) -> List[int]:
    """Ranks whose local (owned + ghost) cells can intersect this rank's: those owning a cell within twice the ghost
    depth of this rank's owned cells (every rank on a replicated level). Keeps the halo-plan construction from
    rebuilding the partition of all R ranks when only the geometric neighbours matter."""
    R, me = part.nranks, part.rank
    layers = _layers_of(part)
    if layers is None:
        return [s for s in range(R) if s != me]
    mesh = part.mesh
    near = part.cell_rank == me
    for _ in range(2 * layers):
        vmask = np.zeros(mesh.nv, dtype=bool)
        vmask[mesh.cells[near].ravel()] = True
        near = vmask[mesh.cells].any(axis=1)
    return [int(s) for s in np.unique(part.cell_rank[near]) if s != me]


def _layers_of(part: Partition) -> int:
    return part._layers


class DistributedOperator:
    """Assembled local matrix (or the matrix-free action of the local form) + halo exchange = the global operator
    restricted to this rank's owned rows."""

    def __init__(self, be, mat, dofmap: DofMap):
        self.be, self.mat, self.map = be, mat, dofmap
        self.mask = be.from_numpy(dofmap.owned.astype(np.float64))

    def mult(self, x, y) -> None:
        if getattr(self.mat, 'matrix_free', False):     # BilinearForm(nonassemble=True).mat of the local mesh: the
            self.mat.bf.apply_arrays(x, y)              # form's action on owned + ghost cells, owned rows complete
        else:
            self.be.spmv(self.mat, x, y)
        self.map.exchange(y)

    def dot(self, x, y) -> float:
        return self.map.dot(self.be, x, y, self.mask)

    def cg(self, b, x, dinv, free, tol: float = 1e-10, maxit: int = 1000):
        """Jacobi-preconditioned CG on the free DOFs (same recurrence as ocmp_krylov kind 0 / NGSolve's CG)."""
        be = self.be
        r = be.zeros(len(b))
        self.mult(x, r)
        r = (b - r) * free
        z = r * dinv
        p = z.clone() if hasattr(z, 'clone') else z.copy()
        rz = self.dot(r, z)
        err0 = np.sqrt(abs(rz))
        Ap = be.zeros(len(b))
        it = 0
        while it < maxit and rz != 0.0:
            self.mult(p, Ap)
            Ap = Ap * free
            alpha = rz / self.dot(p, Ap)
            x += alpha * p
            r -= alpha * Ap
            z = r * dinv
            rzn = self.dot(r, z)
            p = z + (rzn / rz) * p
            rz = rzn
            it += 1
            if np.sqrt(abs(rz)) < tol * err0:
                break
        return it, float(np.sqrt(abs(rz)))
