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
