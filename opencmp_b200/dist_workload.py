"""Distributed (element-partitioned) Poisson workload for bench.py --gpus N: H1 order-p Poisson + mass on a
structured [n, n*R] triangle mesh, R ranks, one cell block per rank. Exercises the halo-exchange SpMV, the all-reduced
dot products and Jacobi-CG of opencmp_b200/dist.py on real GPUs (NCCL over NVLink)."""
from __future__ import annotations

import time

import numpy as np

from . import ngs
from .dist import Partition, DofMap, DistributedOperator
from .mesh import structured_2d


class DistributedPoisson:
    def __init__(self, n_per_rank: int, order: int, world: int, rank: int):
        be = ngs.get_backend()
        self.be = be
        gmesh = structured_2d([n_per_rank, n_per_rank * world], scale=(1.0, float(world)))
        self.part = Partition(gmesh, world, rank, layers=1)
        gfes = ngs.FESpace([ngs.H1(ngs.Mesh(gmesh), order=order, dirichlet='bottom|top')])     # numbering only
        self.mesh = ngs.Mesh(self.part.local_mesh())
        self.fes = ngs.FESpace([ngs.H1(self.mesh, order=order, dirichlet='bottom|top')])
        self.map = DofMap(self.part, gfes, self.fes)
        self.nglobal, self.ncells_global = gfes.ndof, gmesh.ne
        u, v = self.fes.TrialFunction()[0], self.fes.TestFunction()[0]
        self.a = ngs.BilinearForm(self.fes)
        self.a += (ngs.InnerProduct(ngs.Grad(u), ngs.Grad(v)) + u * v) * ngs.dx
        self.L = ngs.LinearForm(self.fes)
        self.L += (1.0 + ngs.sin(3 * ngs.x) * ngs.y) * v * ngs.dx
        self.a.Assemble()
        self.L.Assemble()
        self.op = DistributedOperator(be, self.a.mat, self.map)
        free = be.from_numpy(self.fes.FreeDofs().astype(np.float64))
        diag = self.a.mat.values[be._up(self.fes.pattern().diag.astype(np.int64))]
        self.dinv = free / diag
        self.map.exchange(self.dinv)
        self.free = free
        self.b = self.L.vec.a.clone()
        self.map.exchange(self.b)
        owned_rows = np.nonzero(self.map.owned)[0]
        pat = self.fes.pattern()
        self.nnz_owned = int((pat.rowptr[owned_rows + 1] - pat.rowptr[owned_rows]).sum())
        self.nrows_owned = int(len(owned_rows))

    def spmv_bytes_owned(self) -> int:
        return self.nnz_owned * 12 + self.nrows_owned * 20

    def time_spmv(self, iters: int = 50):
        import torch
        x = self.be.from_numpy(np.random.default_rng(1).uniform(-1, 1, self.fes.ndof))
        y = self.be.zeros(self.fes.ndof)
        for _ in range(5):
            self.op.mult(x, y)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            self.op.mult(x, y)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters

    def solve(self, tol=1e-8, maxit=200):
        import torch
        x = self.be.zeros(self.fes.ndof)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        its, res = self.op.cg(self.b, x, self.dinv, self.free, tol=tol, maxit=maxit)
        torch.cuda.synchronize()
        return its, res, time.perf_counter() - t0, x


class DistributedINS:
    """The INS Taylor-Green time step of workloads.INSTaylorGreen, element-partitioned: R ranks, rank r owns the
    strip [0,pi] x [r pi, (r+1) pi] (n0 x n0 coarse squares refined k times), geometric multigrid + GMRES distributed
    with dist_mg.DistributedMultigrid. Per-rank work is fixed as R grows (weak scaling)."""

    def __init__(self, N: int, world: int, rank: int, order: int = 3, n0: int = 4, strips: int = None,
                 replicate_below: int = 100000, **kw):
        from .workloads import INSTaylorGreen
        from .dist_mg import DistributedMultigrid
        self.world, self.rank = world, rank
        k, n = 0, N
        while n % 2 == 0 and n > n0:
            n //= 2
            k += 1
        strips = world if strips is None else strips          # domain [0,pi] x [0, strips*pi]
        gmesh = structured_2d([n, n * strips], scale=(np.pi, np.pi * strips))
        for _ in range(k):
            gmesh.Refine()
        self.gmesh = gmesh
        self.part = Partition(gmesh, world, rank, layers=2)
        outer = self

        class _Local(INSTaylorGreen):
            def _integrate(self, cf):
                val = ngs.Integrate(cf, self.mesh, definedon=self.mesh.Materials('owned'))
                import torch
                import torch.distributed as dist
                if dist.is_initialized() and dist.get_world_size() > 1:
                    t = torch.tensor([val], dtype=torch.float64,
                                     device='cuda' if ngs.get_backend().name == 'cuda' else 'cpu')
                    dist.all_reduce(t)
                    val = float(t[0])
                return val

            def apply_dirichlet_bcs(self):
                super().apply_dirichlet_bcs()
                outer.mg.levels[-1].map.exchange(self.gfu.vec.a)

            def assemble(self):
                self.a.Assemble()
                self.L.Assemble()
                outer.mg.levels[-1].map.exchange(self.L.vec.a)
                outer.mg.update()

            def linear_solve(self):
                it, res = outer.mg.gmres(self.L.vec.a, self.gfu.vec.a, tol=self.linear_tolerance,
                                         maxit=self.linear_max_iterations, restart=50)
                self.linear_iterations.append(it)

        self.w = _Local(N, order=order, mesh=self.part.local_mesh(), preconditioner=None, **kw)
        self.mg = DistributedMultigrid(ngs.get_backend(), self.w.a, gmesh, self.part, replicate_below=replicate_below)
        top = self.mg.levels[-1].map
        for gf in (self.w.gfu, self.w.gfu_0):
            top.exchange(gf.vec.a)
        vmap = self._velocity_map()
        vmap.exchange(self.w.W.vec.a)
        self._vmap = vmap
        step0 = self.w.step

        def step():
            out = step0()
            vmap.exchange(self.w.W.vec.a)
            return out
        self.w.step = step

    def _velocity_map(self):
        gV = ngs.HDiv(ngs.Mesh(self.gmesh), order=self.w.order, dirichlet=self.w.dirichlet, dgjumps=True)
        return DofMap(self.part, gV, self.w.V)

    def step(self):
        return self.w.step()

    @property
    def ndof_global(self):
        return self.mg.levels[-1].map.nglobal


def sphere_total_cells(N: int, world: int) -> int:
    """Cells per direction of the one-sphere box on ``world`` ranks: the even number closest to N world^(1/3) whose
    coarsening ends on a small coarsest mesh — N, 4N/3, 5N/3, 2N for 1, 2, 4, 8 ranks (48, 64, 80, 96 at N = 48)."""
    factor = {1: (1, 1), 2: (4, 3), 4: (5, 3), 8: (2, 1)}.get(world)
    if factor is None:
        raise ValueError("layout 'sphere' is defined for 1, 2, 4 or 8 ranks")
    ntot = N * factor[0] // factor[1]
    if (N * factor[0]) % factor[1] or ntot % 2:
        raise ValueError("layout 'sphere' on {} ranks needs N divisible by {} with an even total ({})".format(
            world, 2 * factor[1], ntot))
    return ntot


class DistributedINSDIM3D:
    """The 3-D INS-DIM time step of workloads.INSSphereDIM3D, element-partitioned: R ranks, rank r owns the brick
    [-1 + 2r, 1 + 2r] x [-1,1]^2 (n0^3 coarse hexes refined k times) with its own diffuse-interface sphere; geometric
    multigrid + GMRES distributed with dist_mg.DistributedMultigrid (the phase field gets a stand-in on every local
    coarse level). Per-rank work is fixed as R grows (weak scaling)."""

    def __init__(self, N: int, world: int, rank: int, order: int = 2, n0: int = 2, bricks: int = None,
                 replicate_below: int = 100000, layout: str = None, **kw):
        """layout 'bricks' (default; ``bricks`` = number of [-1,1]^3 bricks in a row along x, one sphere each, N^3
        cells per brick, contiguous cell blocks per rank) or 'sphere': ONE sphere in [-1,1]^3 (BASELINE configs[4] as
        written), the box meshed with Ntot^3 hexes, Ntot = sphere_total_cells(N, world) growing with the rank count
        (N, 4N/3, 5N/3, 2N on 1, 2, 4, 8 ranks), split into dist.brick_grid(world) compact bricks of cells, one per
        rank. Isotropic cells throughout: refining only the split directions (cells of aspect ratio 2) kept the cell
        count per rank exactly constant but GMRES no longer converged (48 x 24 x 24 on one B200: 400 iterations
        without reaching 1e-12 against 45 on 24^3; profiles/r2_bench_results.md), so the per-rank load varies
        instead: N^3, 1.19 N^3, 1.16 N^3, N^3 cells on 1, 2, 4, 8 ranks."""
        from .mesh import structured_3d
        from .workloads import INSSphereDIM3D
        from .dist import block_ranks, brick_grid
        from .dist_mg import DistributedMultigrid
        self.world, self.rank = world, rank
        if layout == 'sphere':
            grid = brick_grid(world)
            ntot = sphere_total_cells(N, world)
            k, n = 0, ntot
            while n % 2 == 0 and n > n0:
                n //= 2
                k += 1
            gmesh = structured_3d([n, n, n], scale=(2.0, 2.0, 2.0), offset=(1.0, 1.0, 1.0))
            rank_of_cells = block_ranks(grid, (-1.0,) * 3, (1.0,) * 3)
            kw.setdefault('periodic', (False, False, False))
            N = ntot
        else:
            k, n = 0, N
            while n % 2 == 0 and n > n0:
                n //= 2
                k += 1
            bricks = world if bricks is None else bricks
            gmesh = structured_3d([n * bricks, n, n], scale=(2.0 * bricks, 2.0, 2.0), offset=(1.0, 1.0, 1.0))
            rank_of_cells = None
        for _ in range(k):
            gmesh.Refine()
        self.gmesh = gmesh
        self.part = Partition(gmesh, world, rank, layers=2, rank_of_cells=rank_of_cells)
        outer = self

        def integrate(cf):
            w = outer.w
            val = ngs.Integrate(cf, w.mesh, definedon=w.mesh.Materials('owned'))
            import torch
            import torch.distributed as dist
            if dist.is_initialized() and dist.get_world_size() > 1:
                t = torch.tensor([val], dtype=torch.float64,
                                 device='cuda' if ngs.get_backend().name == 'cuda' else 'cpu')
                dist.all_reduce(t)
                val = float(t[0])
            return val

        class _Local(INSSphereDIM3D):
            def apply_dirichlet_bcs(self):
                super().apply_dirichlet_bcs()
                if getattr(outer, 'mg', None) is not None:
                    outer.mg.levels[-1].map.exchange(self.gfu.vec.a)

            def assemble(self):
                self.a.Assemble()
                self.L.Assemble()
                outer.mg.levels[-1].map.exchange(self.L.vec.a)
                outer.mg.update()

            def linear_solve(self):
                it, res = outer.mg.gmres(self.L.vec.a, self.gfu.vec.a, tol=self.linear_tolerance,
                                         maxit=self.linear_max_iterations, restart=self.gmres_restart)
                self.linear_iterations.append(it)

        self.mg = None
        self.w = None
        w = _Local.__new__(_Local)
        self.w = w
        _Local.__init__(w, N, order=order, mesh=self.part.local_mesh(), preconditioner=None, integrate=integrate, **kw)
        # coefficient fields are projected on the local mesh: on its outer (cut) faces the local average differs from
        # the global one, so ghost entries take the owner's value like every other consistent vector
        gH = ngs.H1(ngs.Mesh(gmesh), order=order)
        hmap = DofMap(self.part, gH, w.fes_phi)
        for gf in (w.phi, w.mask):
            hmap.exchange(gf.vec.a)
        self.mg = DistributedMultigrid(ngs.get_backend(), w.a, gmesh, self.part, replicate_below=replicate_below)
        top = self.mg.levels[-1].map
        for gf in (w.gfu, w.gfu_0):
            top.exchange(gf.vec.a)
        gV = ngs.VectorH1(ngs.Mesh(gmesh), order=order, dirichlet=w.dirichlet)
        self._vmap = DofMap(self.part, gV, w.V)
        self._vmap.exchange(w.W.vec.a)
        step0 = w.step

        def step():
            out = step0()
            self._vmap.exchange(w.W.vec.a)
            return out
        w.step = step

    def step(self):
        return self.w.step()

    @property
    def ndof_global(self):
        return self.mg.levels[-1].map.nglobal
