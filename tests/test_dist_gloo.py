"""World-size-2 gloo tests (CPU) of the element-partitioned layer: partition, local meshes, DOF ownership, halo
exchange, distributed SpMV / dot products / Jacobi-CG — compared with the single-process result."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from opencmp_b200.mesh import structured_2d


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _build(ngs, mesh, family, order, dg):
    m = ngs.Mesh(mesh)
    fes = ngs.FESpace([getattr(ngs, family)(m, order=order, dirichlet='left|bottom', dgjumps=dg)], dgjumps=dg)
    u, v = fes.TrialFunction()[0], fes.TestFunction()[0]
    a = ngs.BilinearForm(fes)
    a += (ngs.InnerProduct(ngs.Grad(u), ngs.Grad(v)) + u * v) * ngs.dx
    if dg:
        n = ngs.specialcf.normal(2)
        h = ngs.specialcf.mesh_size
        ju, jv = u - u.Other(), v - v.Other()
        gu, gv = 0.5 * (ngs.Grad(u) + ngs.Grad(u.Other())), 0.5 * (ngs.Grad(v) + ngs.Grad(v.Other()))
        a += (-ju * n * gv - gu * n * jv + 40.0 / h * ju * jv) * ngs.dx(skeleton=True)
    L = ngs.LinearForm(fes)
    L += (1.0 + ngs.x * ngs.sin(3 * ngs.y)) * v * ngs.dx
    a.Assemble()
    L.Assemble()
    return m, fes, a, L


def _worker(rank, world, port, family, order, dg, out):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        import opencmp_b200.ngs as ngs
        from oracle.backend import OracleBackend
        from opencmp_b200.dist import Partition, DofMap, DistributedOperator
        be = OracleBackend()
        ngs.set_backend(be)
        gmesh = structured_2d([8, 6])
        gm, gfes, ga, gL = _build(ngs, gmesh, family, order, dg)
        part = Partition(gmesh, world, rank, layers=1)
        lmesh = part.local_mesh()
        lm, lfes, la, lL = _build(ngs, lmesh, family, order, dg)
        dm = DofMap(part, gfes, lfes)
        # every global dof is owned exactly once
        cnt = torch.zeros(gfes.ndof, dtype=torch.float64)
        cnt[torch.from_numpy(dm.l2g[dm.owned])] = 1.0
        dist.all_reduce(cnt)
        assert bool((cnt == 1.0).all())
        op = DistributedOperator(be, la.mat, dm)
        # SpMV: owned rows of the local product equal the global product
        rng = np.random.default_rng(0)
        xg = rng.uniform(-1, 1, gfes.ndof)
        yg = np.zeros(gfes.ndof)
        be.spmv(ga.mat, xg, yg)
        xl = xg[dm.l2g].copy()
        yl = np.zeros(lfes.ndof)
        op.mult(xl, yl)
        assert np.abs(yl - yg[dm.l2g]).max() < 1e-12 * np.abs(yg).max()       # ghosts refreshed too
        # the same product matrix-free: the action of the local form (BilinearForm(nonassemble=True)) + halo exchange
        twin = ngs.BilinearForm(lfes, nonassemble=True)
        twin += la.integrals
        yf = np.zeros(lfes.ndof)
        DistributedOperator(be, twin.mat, dm).mult(xl, yf)
        assert np.abs(yf - yg[dm.l2g]).max() < 1e-12 * np.abs(yg).max()
        # right-hand side: owned entries complete
        assert np.abs(lL.vec.a[dm.owned] - gL.vec.a[dm.l2g[dm.owned]]).max() < 1e-12
        # dot product
        assert abs(op.dot(xl, yl) - float(xg @ yg)) < 1e-10 * abs(float(xg @ yg))
        # reverse_add: ghost contributions are summed into the owner
        z = np.ones(lfes.ndof)
        dm.exchange(z, reverse_add=True)
        mult = torch.zeros(gfes.ndof, dtype=torch.float64)
        mult[torch.from_numpy(dm.l2g)] += 1.0
        dist.all_reduce(mult)
        assert np.abs(z[dm.owned] - mult.numpy()[dm.l2g[dm.owned]]).max() == 0.0
        # Jacobi-CG on the free dofs against the single-process direct solve
        free_l = lfes.FreeDofs().astype(np.float64)
        diag = la.mat.values[lfes.pattern().diag]
        dinv = np.where(free_l > 0, 1.0 / diag, 0.0)
        dm.exchange(dinv)                    # ghost rows are incomplete locally: take the owner's diagonal
        b = lL.vec.a.copy()
        dm.exchange(b)
        x = np.zeros(lfes.ndof)
        its, res = op.cg(b, x, dinv, free_l, tol=1e-13, maxit=2000)
        from oracle import fem
        xref = fem.solve_direct(fem.csr_matrix(gfes, ga.mat.values), gL.vec.a, np.zeros(gfes.ndof), gfes.FreeDofs())
        err = np.abs(x - xref[dm.l2g]).max() / np.abs(xref).max()
        assert err < 1e-9, err
        out[rank] = (its, err)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('family,order,dg', [('H1', 2, False), ('L2', 1, True)])
def test_partitioned_operator_matches_global(family, order, dg):
    world = 2
    port = _free_port()
    mgr = mp.get_context('spawn').Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, family, order, dg, out), nprocs=world, join=True)
    assert len(out) == world


def _mg_worker(rank, world, port, replicate_below, out):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        import opencmp_b200.ngs as ngs
        from oracle.backend import OracleBackend
        ngs.set_backend(OracleBackend())
        from opencmp_b200.dist_workload import DistributedINS
        from opencmp_b200.workloads import INSTaylorGreen
        d = DistributedINS(4, world, rank, order=2, n0=2, replicate_below=replicate_below)
        g = INSTaylorGreen(4, order=2, mesh=d.gmesh, linear_solver='direct', preconditioner=None)
        top = d.mg.levels[-1].map
        d.w.gfu.vec.a[:] = g.gfu.vec.a[top.l2g]
        d.w.gfu_0.vec.a[:] = g.gfu_0.vec.a[top.l2g]
        d.w.W.vec.a[:] = g.W.vec.a[d._vmap.l2g]
        d.step()
        g.step()
        nu = d.w.V.ndof
        err = np.abs(d.w.gfu.vec.a - g.gfu.vec.a[top.l2g])[:nu].max() / np.abs(g.gfu.vec.a).max()
        assert err < 1e-9, err
        assert d.w.picard_iterations == g.picard_iterations
        eu_d, _ = d.w.errors()
        eu_g, _ = g.errors()
        assert abs(eu_d - eu_g) < 1e-9 * eu_g + 1e-14          # all-reduced error norm equals the global one
        out[rank] = (err, list(d.w.linear_iterations))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('replicate_below', [0, 100000])
def test_partitioned_multigrid_ins_step_matches_global_solve(replicate_below):
    """One INS time step (Picard loop, assembly, distributed geometric multigrid + GMRES with halo exchange,
    reverse-add and all-reduces) on 2 ranks equals the single-process step with a sparse direct solve."""
    world = 2
    port = _free_port()
    mgr = mp.get_context('spawn').Manager()
    out = mgr.dict()
    mp.spawn(_mg_worker, args=(world, port, replicate_below, out), nprocs=world, join=True)
    assert len(out) == world
    assert out[0][1] == out[1][1]                                 # same iteration counts on both ranks


def _mg3d_worker(rank, world, port, bricks, out):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        import opencmp_b200.ngs as ngs
        from oracle.backend import OracleBackend
        ngs.set_backend(OracleBackend())
        from opencmp_b200.dist_workload import DistributedINSDIM3D
        from opencmp_b200.workloads import INSSphereDIM3D
        kw = dict(nonlinear_max_iterations=1, linear_tolerance=1e-13, lam=1.0)
        if bricks == 'sphere':          # ONE sphere, [-1,1]^3 meshed with 8^3 hexes on 2 ranks, one compact brick each
            d = DistributedINSDIM3D(6, world, rank, n0=2, replicate_below=0, layout='sphere', **kw)
            assert d.gmesh.ne == 8 ** 3 and np.allclose(d.gmesh.points.max(axis=0), 1.0)
            assert (d.part.cell_rank == rank).sum() == 8 ** 3 // 2
            kw = dict(kw, periodic=(False, False, False))
        else:
            d = DistributedINSDIM3D(4, world, rank, n0=2, replicate_below=0, bricks=bricks, **kw)
        g = INSSphereDIM3D(d.gmesh.ne, mesh=d.gmesh, preconditioner=None, **kw)

        def direct():
            inv = g.a.mat.Inverse(g.fes.FreeDofs())
            r = g.L.vec.CreateVector()
            r.data = g.L.vec - g.a.mat * g.gfu.vec
            g.gfu.vec.data += inv * r
            g.linear_iterations.append(0)
        g.linear_solve = direct
        top = d.mg.levels[-1].map
        assert np.abs(d.w.phi.vec.a - g.phi.vec.a[DofMapOf(d, g)]).max() < 1e-12
        d.step()
        g.step()
        nu = d.w.V.ndof
        ref = g.gfu.vec.a[top.l2g]
        err = np.abs(d.w.gfu.vec.a - ref)[:nu].max() / np.abs(ref[:nu]).max()
        # 1e-6, not 1e-9: the reference's phi >= 1e-10 clamp makes the system's condition number exceed 1e10
        assert err < 1e-6, err
        eu_d, _ = d.w.errors()
        eu_g, _ = g.errors()
        assert abs(eu_d - eu_g) < 1e-6 * eu_g
        out[rank] = (err, list(d.w.linear_iterations))
    finally:
        dist.destroy_process_group()


def DofMapOf(d, g):
    """local -> global DOF map of the phase-field space of a DistributedINSDIM3D."""
    from opencmp_b200.dist import DofMap
    return DofMap(d.part, g.fes_phi, d.w.fes_phi).l2g


@pytest.mark.parametrize('bricks', [None, 1, 'sphere'])
def test_partitioned_multigrid_ins_dim_3d_step_matches_global_solve(bricks):
    """3-D INS-DIM (hex Taylor-Hood): distributed multigrid-GMRES with open-star patches and per-level phase fields on
    2 ranks equals the single-process sparse direct solve — with one brick and its own diffuse sphere per rank
    (``bricks=None``), and with ONE sphere in [-1,1]^3 whose cells are split between the ranks (``bricks=1``, the
    layout BASELINE configs[4] describes; ``'sphere'``: the same with the mesh refined with the rank count and one
    compact brick of cells per rank, bench.py's weak-scaling layout)."""
    world = 2
    port = _free_port()
    mgr = mp.get_context('spawn').Manager()
    out = mgr.dict()
    mp.spawn(_mg3d_worker, args=(world, port, bricks, out), nprocs=world, join=True)
    assert len(out) == world
    assert out[0][1] == out[1][1] and max(out[0][1]) < 80


def test_partitioned_multigrid_with_lagged_smoother(monkeypatch):
    """The same 2-rank INS step with OCMP_MG_LAG=1: the lag decision is taken from the (global) GMRES iteration count,
    so both ranks skip the same set-ups; the step still equals the single-process direct solve."""
    monkeypatch.setenv('OCMP_MG_LAG', '1')
    world = 2
    port = _free_port()
    mgr = mp.get_context('spawn').Manager()
    out = mgr.dict()
    mp.spawn(_mg_worker, args=(world, port, 0, out), nprocs=world, join=True)
    assert len(out) == world
    assert out[0][1] == out[1][1]


def _block_exchange_worker(rank, world, port, out):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        import opencmp_b200.ngs as ngs
        from oracle.backend import OracleBackend
        ngs.set_backend(OracleBackend())
        from opencmp_b200.dist import Partition, DofMap, block_ranks, brick_grid
        from opencmp_b200.mesh import structured_3d
        gmesh = structured_3d([6, 6, 4], scale=(2.0,) * 3, offset=(1.0,) * 3)
        grid = brick_grid(world)
        part = Partition(gmesh, world, rank, layers=2, rank_of_cells=block_ranks(grid, (-1.0,) * 3, (1.0,) * 3))
        assert (part.cell_rank == rank).sum() == gmesh.ne // world        # compact bricks of equal size
        gfes = ngs.FESpace([ngs.VectorH1(ngs.Mesh(gmesh), order=2), ngs.H1(ngs.Mesh(gmesh), order=1)])
        lmesh = ngs.Mesh(part.local_mesh())
        lfes = ngs.FESpace([ngs.VectorH1(lmesh, order=2), ngs.H1(lmesh, order=1)])
        m = DofMap(part, gfes, lfes)
        # ghost refresh: owners hold a function of the global dof number, ghosts start as garbage
        x = np.where(m.owned, np.sin(0.1 * m.l2g) + m.l2g, -1e30)
        m.exchange(x)
        assert np.array_equal(x, np.sin(0.1 * m.l2g) + m.l2g)
        # partial sums: every rank contributes 1 on each of its local entries; the total is the number of ranks that
        # hold the dof, which must agree on all of them
        y = np.ones(m.nlocal)
        m.exchange_sum(y)
        cnt = np.zeros(m.nglobal)
        cnt[m.l2g] = 1.0
        tot = torch.from_numpy(cnt)
        dist.all_reduce(tot)
        assert np.array_equal(y, tot.numpy()[m.l2g])
        # ownership is a partition of the global dofs
        own = np.zeros(m.nglobal)
        own[m.l2g[m.owned]] = 1.0
        t2 = torch.from_numpy(own)
        dist.all_reduce(t2)
        assert np.array_equal(t2.numpy(), np.ones(m.nglobal))
        out[rank] = int(len(m.send))
    finally:
        dist.destroy_process_group()


def test_halo_plans_of_a_brick_partition_on_four_ranks():
    """2 x 2 x 1 compact bricks (bench.py's layout on 4 GPUs): the halo plans are built from the peers' partitions, which
    must use the same cell -> rank rule — the default contiguous blocks only coincide with the bricks on 2 ranks."""
    world = 4
    port = _free_port()
    mgr = mp.get_context('spawn').Manager()
    out = mgr.dict()
    mp.spawn(_block_exchange_worker, args=(world, port, out), nprocs=world, join=True)
    assert len(out) == world and all(v == 3 for v in out.values())       # every brick touches the three others


def _dim_counts_worker(rank, world, port, out):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        import opencmp_b200.ngs as ngs
        from oracle.backend import OracleBackend
        ngs.set_backend(OracleBackend())
        from opencmp_b200.dist_workload import DistributedINSDIM3D
        d = DistributedINSDIM3D(6 if world == 2 else 8, world, rank, layout='sphere', replicate_below=3000,
                                nonlinear_max_iterations=1, nonlinear_tolerance=(0.0, 0.0), wall_period=0.1)
        assert d.gmesh.ne == 8 ** 3
        assert [lv.replicated for lv in d.mg.levels] == [True, True, False]
        d.step()
        out[rank] = list(d.w.linear_iterations)
    finally:
        dist.destroy_process_group()


def test_partitioned_dim_step_with_thin_interface_keeps_the_single_process_iteration_count():
    """3-D INS-DIM with the bench's interface width (lambda = 0.25: phi reaches the 1e-10 clamp, the continuity rows of
    the solid region are scaled down by up to ten orders of magnitude) on 8^3 hexes: two ranks (two bricks, two ghost
    layers, partitioned finest level, replicated coarse levels with the fine level's volume penalty —
    multigrid.inherit_cell_penalty) need exactly the GMRES iterations of the single-process run."""
    counts = {}
    for world in (1, 2):
        port = _free_port()
        mgr = mp.get_context('spawn').Manager()
        out = mgr.dict()
        mp.spawn(_dim_counts_worker, args=(world, port, out), nprocs=world, join=True)
        assert len(out) == world and all(out[r] == out[0] for r in range(world))
        counts[world] = out[0]
    assert counts[2] == counts[1] and max(counts[1]) < 30
