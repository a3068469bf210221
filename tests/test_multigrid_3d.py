"""Host logic of the 3-D path (CPU, oracle backend): hex refinement, exact prolongation on the hex hierarchy,
open-star / Vanka vertex patches, coarse-level stand-ins of coefficient fields, and the geometric-multigrid GMRES on the
3-D INS-DIM workload compared with the sparse direct solve."""
import numpy as np
import pytest

import cases


@pytest.fixture()
def ngs():
    import opencmp_b200.ngs as ngs
    from oracle.backend import OracleBackend
    old = ngs._backend
    ngs.set_backend(OracleBackend())
    yield ngs
    ngs.set_backend(old)


@pytest.mark.parametrize('cell', ['quad', 'hex'])
def test_tensor_refine(cell):
    from opencmp_b200.mesh import structured_2d, structured_3d
    if cell == 'quad':
        m, ref = structured_2d([3, 2], cell='quad'), structured_2d([6, 4], cell='quad')
    else:
        m, ref = structured_3d([2, 3, 2]), structured_3d([4, 6, 4])
    ne0 = m.ne
    pc = m.points[m.cells].mean(axis=1)
    m.Refine()
    nch = 2 ** m.dim
    assert m.ne == nch * ne0 == ref.ne and m.nv == ref.nv and m.nf == ref.nf
    m.check_affine()
    assert np.allclose(np.linalg.det(m.jacobians()), np.linalg.det(ref.jacobians())[0])
    # children 2^d e .. 2^d e + 2^d - 1 tile their parent
    cent = m.points[m.cells].mean(axis=1).reshape(ne0, nch, m.dim).mean(axis=1)
    assert np.abs(cent - pc).max() < 1e-14
    assert m.coarse.ne == ne0
    for r in range(len(m.bnd_names)):
        assert (m.bnd_region == r).sum() == (ref.bnd_region == r).sum()
    # same point set as the directly generated fine mesh
    a = np.unique(np.round(m.points, 12), axis=0)
    b = np.unique(np.round(ref.points, 12), axis=0)
    assert np.abs(a - b).max() < 1e-12


def test_hex_prolongation_is_exact(ngs):
    from opencmp_b200.mesh import structured_3d
    from opencmp_b200.multigrid import clone_space, prolongation
    mesh = structured_3d([2, 2, 2])
    mesh.Refine()
    mf = ngs.Mesh(mesh)
    fine = ngs.FESpace([ngs.VectorH1(mf, order=2), ngs.H1(mf, order=1)])
    coarse = clone_space(fine, mesh.coarse)
    P = prolongation(coarse, fine)
    x, y, z = ngs.x, ngs.y, ngs.z
    u = ngs.CoefficientFunction((x * y + z * z, 1.0 - 2.0 * x * z, y * y * x))     # in Q2^3
    p = 1.0 + x - 2.0 * y + 3.0 * z + x * y * z                                     # in Q1
    gc, gf = ngs.GridFunction(coarse), ngs.GridFunction(fine)
    for g in (gc, gf):
        g.components[0].Set(u)
        g.components[1].Set(p)
    assert np.abs(P @ gc.vec.NumPy() - gf.vec.NumPy()).max() < 1e-10


def test_quad_prolongation_is_exact(ngs):
    """2-D quadrilateral hierarchy (the reference's DIM meshes default to quads, expanded_config_parser.py:79)."""
    from opencmp_b200.mesh import structured_2d
    from opencmp_b200.multigrid import clone_space, prolongation
    mesh = structured_2d([3, 2], cell='quad')
    mesh.Refine()
    mf = ngs.Mesh(mesh)
    fine = ngs.FESpace([ngs.H1(mf, order=3), ngs.L2(mf, order=2)])
    coarse = clone_space(fine, mesh.coarse)
    P = prolongation(coarse, fine)
    x, y = ngs.x, ngs.y
    gc, gf = ngs.GridFunction(coarse), ngs.GridFunction(fine)
    for g in (gc, gf):
        g.components[0].Set(x * x * x * y * y - 2.0 * x * y * y * y + 1.0)
        g.components[1].Set(x * x * y * y - x + 3.0 * y)
    assert np.abs(P @ gc.vec.NumPy() - gf.vec.NumPy()).max() < 1e-10


def test_star_and_vanka_patches(ngs):
    from opencmp_b200.mesh import structured_3d
    from opencmp_b200.patches import vertex_patch_dofs
    from oracle.backend import _vertex_patches
    m = ngs.Mesh(structured_3d([4, 4, 4]))
    V = ngs.VectorH1(m, order=2, dirichlet='back|left|front|right|bottom|top')
    fes = ngs.FESpace([V, ngs.H1(m, order=1)])
    closed = vertex_patch_dofs(fes, 'vertex', drop_constrained=False)
    ref = _vertex_patches(fes, None)
    assert all(np.array_equal(r, g[g >= 0]) for r, g in zip(ref, closed))
    star = vertex_patch_dofs(fes, 'star')
    vanka = vertex_patch_dofs(fes, 'vanka')
    centre = int(np.argmin(np.abs(m.points - 0.5).sum(axis=1)))
    assert (star[centre] >= 0).sum() == 3 * 27 + 1            # interior Q2 nodes of the 2x2x2 star + its own pressure
    assert (vanka[centre] >= 0).sum() == 3 * 27 + 27
    assert star.shape[1] <= 160 and closed.shape[1] == 3 * 125 + 27
    free = fes.FreeDofs()
    for tab in (star, vanka):
        d = tab[tab >= 0]
        assert free[d].all()                                   # constrained DOFs dropped
        assert np.array_equal(np.unique(d), np.nonzero(free)[0])   # every free DOF is covered
    for v in range(m.nv):
        assert set(star[v][star[v] >= 0]) <= set(vanka[v][vanka[v] >= 0]) <= set(closed[v][closed[v] >= 0])


def test_coarse_phase_field_and_multigrid_solve(ngs):
    """3-D INS-DIM, 4^3 hexes refined from 2^3: the coarse level sees a restricted phase field (non-singular coarse
    operator) and multigrid-GMRES reproduces the direct solve. Tolerance: see tests/test_gpu_parity.py
    ::test_ins_dim_3d_multigrid_step (condition number > 1e10 from the reference's phi >= 1e-10 clamp)."""
    from opencmp_b200.dist import Partition
    from opencmp_b200.dist_mg import DistributedMultigrid
    from opencmp_b200.mesh import structured_3d
    from opencmp_b200.multigrid import coefficient_fields
    from opencmp_b200.workloads import INSSphereDIM3D
    import scipy.sparse.linalg as spla
    be = ngs.get_backend()
    gm = structured_3d([2, 2, 2], scale=(2.0,) * 3, offset=(1.0,) * 3)
    gm.Refine()
    part = Partition(gm, 1, 0, layers=2)
    w = INSSphereDIM3D(4, preconditioner=None, mesh=part.local_mesh(), lam=1.0)
    fields = coefficient_fields(w.a)
    assert {id(g) for g in fields} == {id(w.phi), id(w.mask)}       # the Oseen wind is not a coefficient field
    mg = DistributedMultigrid(be, w.a, gm, part, replicate_below=0)
    w.t.Set(w.dt.Get())
    w.a.Assemble()
    w.L.Assemble()
    mg.update()
    phi_c = mg.levels[0].field_map[id(w.phi)]
    pc = phi_c.vec.NumPy()
    assert pc.min() >= 1e-10 and pc.max() <= 1.0 and pc.max() > 0.5
    A0 = be._csr(mg.levels[0].mat)
    assert abs(A0).sum() > 0                                         # all-field-weighted form survived on the coarse level
    A, b = be._csr(w.a.mat), w.L.vec.NumPy().copy()
    idx = np.nonzero(np.asarray(w.fes.FreeDofs(), bool))[0]
    x0 = w.gfu.vec.NumPy().copy()
    it, res = mg.gmres(w.L.vec.a, w.gfu.vec.a, tol=1e-13, maxit=200, restart=100)
    xd = x0.copy()
    xd[idx] += spla.splu(A[idx][:, idx].tocsc()).solve((b - A @ x0)[idx])
    nv = w.V.ndof
    got = w.gfu.vec.NumPy()
    assert it < 60
    assert np.linalg.norm((got - xd)[:nv]) < 1e-6 * np.linalg.norm(xd[:nv])


def test_cuda_backend_patch_tables_host_logic(ngs):
    """CudaBackend._patches is host code (NumPy) feeding the smoother kernels: closed stars in 2-D (identical to the
    oracle's), automatic switch to open stars with an even stride in 3-D, owned-vertex subsets for partitioned runs."""
    from opencmp_b200.backend import CudaBackend
    from opencmp_b200.mesh import structured_2d, structured_3d
    from oracle.backend import _vertex_patches
    be = CudaBackend.__new__(CudaBackend)                    # no device needed for the table construction
    cache = {}
    be.space_data = lambda fes: cache.setdefault(id(fes), {})
    be._up = lambda a, dtype=None: np.asarray(a)
    be.zeros = lambda n: np.zeros(int(n))
    m = ngs.Mesh(structured_2d([6, 6], scale=(np.pi, np.pi)))
    fes = ngs.FESpace([ngs.HDiv(m, order=3, dirichlet='top|bottom|left|right', dgjumps=True),
                       ngs.L2(m, order=2, dgjumps=True)], dgjumps=True)
    vm = np.zeros(m.nv, bool)
    vm[::3] = True
    for mask in (None, vm):
        pt = be._patches(fes, 'vertex', mask)
        ref = _vertex_patches(fes, mask)
        assert pt['bs'] == 132 and pt['npatch'] == len(ref)
        assert all(np.array_equal(r, d[d >= 0]) for r, d in zip(ref, pt['dofs']))
    m3 = ngs.Mesh(structured_3d([4, 4, 4]))
    f3 = ngs.FESpace([ngs.VectorH1(m3, order=2, dirichlet='back|left|front|right|bottom|top'), ngs.H1(m3, order=1)])
    p3 = be._patches(f3, 'vertex')
    assert p3['npatch'] == m3.nv and p3['bs'] == 90 and (p3['dofs'][:, -1] == -1).all()
    assert p3['dofs'].dtype == np.int32 and p3['wgt'].min() > 0
    covered = np.unique(p3['dofs'][p3['dofs'] >= 0])
    assert np.array_equal(covered, np.nonzero(f3.FreeDofs())[0])


def test_cuda_backend_patch_tables_fp32_stride(ngs, monkeypatch):
    """OCMP_PATCH_FP32=1 (patch inverses stored in FP32): the patch stride is padded to a multiple of 4 so that every
    stored column starts 16-byte aligned for the bulk copies of k_patch_apply_stream<float>; the DOF content is unchanged."""
    from opencmp_b200.backend import CudaBackend
    from opencmp_b200.mesh import structured_3d
    be = CudaBackend.__new__(CudaBackend)
    be._up = lambda a, dtype=None: np.asarray(a)
    be.zeros = lambda n: np.zeros(int(n))
    m3 = ngs.Mesh(structured_3d([4, 4, 4]))
    f3 = ngs.FESpace([ngs.VectorH1(m3, order=2, dirichlet='back|left|front|right|bottom|top'), ngs.H1(m3, order=1)])
    cache = {}
    be.space_data = lambda fes: cache.setdefault(id(fes), {})
    p64 = be._patches(f3, 'vertex')
    assert p64['bs'] == 90 and p64['storage'] == 'fp64'
    monkeypatch.setenv('OCMP_PATCH_FP32', '1')
    cache.clear()
    p32 = be._patches(f3, 'vertex')
    assert p32['storage'] == 'fp32' and p32['bs'] == 92 and p32['bs'] % 4 == 0
    assert (p32['dofs'][:, 90:] == -1).all() and np.array_equal(p32['dofs'][:, :90], p64['dofs'])
    monkeypatch.setenv('OCMP_PATCH_STORAGE', 'bf16')
    cache.clear()
    p16 = be._patches(f3, 'vertex')
    assert p16['storage'] == 'bf16' and p16['bs'] == 96 and np.array_equal(p16['dofs'][:, :90], p64['dofs'])


def test_coarse_levels_are_reused_until_a_parameter_changes(ngs, monkeypatch):
    """The coarse-level operators of the multigrid hierarchy depend only on the Parameters their programs read (the
    Oseen wind acts on the finest level only): two time steps with reuse give the same iterates and iteration counts as
    with a full set-up on every update, the coarse levels are built once, and a changed dt rebuilds them."""
    from opencmp_b200.dist_workload import DistributedINS

    def run(reuse, change_dt=False):
        monkeypatch.setenv('OCMP_MG_REUSE_COARSE', '1' if reuse else '0')
        d = DistributedINS(8, 1, 0)
        its = []
        for k in range(2):
            if change_dt and k == 1:
                d.w.dt.Set(2.0 * d.w.dt.Get())
            d.w.linear_iterations = []
            d.step()
            its += d.w.linear_iterations
        return d.w.gfu.vec.NumPy().copy(), its, d.mg.coarse_setups
    u_full, its_full, n_full = run(False)
    u_reuse, its_reuse, n_reuse = run(True)
    assert its_reuse == its_full and n_reuse == 1 and n_full == len(its_full)
    assert np.abs(u_reuse - u_full).max() <= 1e-13 * np.abs(u_full).max()
    _, its_dt, n_dt = run(True, change_dt=True)
    assert n_dt == 2


def test_smoother_lag_policy_unit():
    """SmootherLag: off -> always fresh; on -> fresh once, again only after a solve needed > 1.25 x + 2 iterations of
    the first solve after the last fresh set-up, or when forced (coarse levels rebuilt)."""
    import os
    from opencmp_b200.multigrid import SmootherLag
    os.environ.pop('OCMP_MG_LAG', None)
    off = SmootherLag()
    assert all(off.need_refresh(True) for _ in range(3)) and off.fresh_setups == 3
    os.environ['OCMP_MG_LAG'] = '1'
    try:
        lag = SmootherLag()
        assert lag.need_refresh(False)                # nothing stored yet
        lag.note_solve(18)
        assert not lag.need_refresh(True)
        lag.note_solve(15)
        assert not lag.need_refresh(True)
        lag.note_solve(24)                            # 24 <= 1.25 * 18 + 2 = 24.5
        assert not lag.need_refresh(True)
        lag.note_solve(25)
        assert lag.need_refresh(True) and lag.fresh_setups == 2
        lag.note_solve(19)
        assert not lag.need_refresh(True) and lag.need_refresh(True, forced=True)
    finally:
        os.environ.pop('OCMP_MG_LAG', None)


def test_lagged_fine_level_smoother_keeps_the_steps(ngs, monkeypatch):
    """OCMP_MG_LAG=1 on the 2-D INS workload: the finest level's patch inverses are computed once over two time steps,
    iteration counts and solution stay those of the always-fresh smoother."""
    from opencmp_b200.dist_workload import DistributedINS

    def run(lag):
        monkeypatch.setenv('OCMP_MG_LAG', lag)
        d = DistributedINS(8, 1, 0)
        its = []
        for _ in range(2):
            d.w.linear_iterations = []
            d.step()
            its += d.w.linear_iterations
        return d.w.gfu.vec.NumPy().copy(), its, d.mg.lag.fresh_setups
    u0, its0, n0 = run('0')
    u1, its1, n1 = run('1')
    assert n0 == len(its0) and n1 == 1
    assert max(abs(a - b) for a, b in zip(its1, its0)) <= 1
    assert np.abs(u1 - u0).max() < 1e-6 * np.abs(u0).max()          # two GMRES solves to 1e-10, different smoothers


@pytest.mark.parametrize('storage', ['fp32', 'bf16'])
def test_reduced_precision_storage_of_patch_inverses_keeps_iteration_counts(ngs, monkeypatch, storage):
    """The claim behind OCMP_PATCH_STORAGE: patch inverses computed in FP64 and *stored* in FP32 / bfloat16
    precondition the FP64 GMRES just as well. Emulated on the CPU restatement by rounding the oracle's inverses to the
    storage type: same iteration counts (bf16: within 2) and the same solution on the 2-D INS workload."""
    import torch
    from oracle.backend import OracleBackend
    from opencmp_b200.dist_workload import DistributedINS

    def run():
        d = DistributedINS(8, 1, 0)
        d.w.linear_iterations = []
        d.step()
        return d.w.gfu.vec.NumPy().copy(), list(d.w.linear_iterations)
    u64, its64 = run()
    plain = OracleBackend.patch_setup
    dtype = torch.float32 if storage == 'fp32' else torch.bfloat16

    def rounded(self, mat, pt, fm):
        plain(self, mat, pt, fm)
        pt['inv'] = [torch.from_numpy(a).to(dtype).to(torch.float64).numpy() for a in pt['inv']]
    monkeypatch.setattr(OracleBackend, 'patch_setup', rounded)
    u, its = run()
    assert len(its) == len(its64) and max(abs(a - b) for a, b in zip(its, its64)) <= (0 if storage == 'fp32' else 2)
    assert np.abs(u - u64).max() < 1e-6 * np.abs(u64).max()


def test_cell_mesh_size_scale_touches_cell_integrals_only(ngs):
    """lower_form(cell_mesh_size_scale=s): specialcf.mesh_size becomes s h inside cell integrals; boundary-facet
    integrals keep the level's own h."""
    from opencmp_b200.mesh import structured_2d
    from opencmp_b200.symbolic import lower_form
    m = ngs.Mesh(structured_2d([3, 3]))
    fes = ngs.H1(m, order=2)
    u, v = fes.TnT()
    h = ngs.specialcf.mesh_size
    be = ngs.get_backend()

    def assembled(integrals, scale):
        mat = ngs.Matrix(fes)
        be.assemble_matrix(lower_form(fes, integrals, 2, cell_mesh_size_scale=scale), mat)
        return be._csr(mat).toarray()
    cell, facet = (1.0 / h) * u * v * ngs.dx, (1.0 / h) * u * v * ngs.ds
    assert np.abs(assembled(cell, 0.5) - 2.0 * assembled(cell, 1.0)).max() < 1e-13
    assert np.abs(assembled(facet, 0.5) - assembled(facet, 1.0)).max() == 0.0
    both = assembled(cell + facet, 0.25)
    assert np.abs(both - (4.0 * assembled(cell, 1.0) + assembled(facet, 1.0))).max() < 1e-12


def test_coarse_levels_inherit_the_fine_volume_penalty(ngs, monkeypatch):
    """3-D INS-DIM: the volume penalty alpha = ipc k^2 / h of the diffuse-interface forms keeps the FINE level's value in
    the coarse multigrid operators (multigrid.inherit_cell_penalty). The 8^3 solve needs no more GMRES iterations than
    with re-scaled penalties, and OCMP_MG_COARSE_H=own restores the old operators. Both stop when the PRECONDITIONED
    residual fell by 1e-12 (the rule of ngsolve.solvers.GMRes); measured against a sparse direct solve that leaves a
    velocity error of 2.0e-5 (fine) / 3.5e-5 (own) at 8^3 — condition number > 1e10 from the phi >= 1e-10 clamp — so
    the two iterates agree to 5e-4, not to round-off."""
    from opencmp_b200.dist_workload import DistributedINSDIM3D
    from opencmp_b200.multigrid import coarse_mesh_size_scales

    def solve(mode):
        monkeypatch.setenv('OCMP_MG_COARSE_H', mode)
        d = DistributedINSDIM3D(8, 1, 0, layout='sphere', wall_period=0.1)
        w = d.w
        w.t.Set(w.dt.Get())
        w.assemble()
        w.linear_solve()
        return w.linear_iterations[-1], w.gfu.vec.NumPy().copy(), w.V.ndof
    monkeypatch.setenv('OCMP_MG_COARSE_H', 'fine')
    assert coarse_mesh_size_scales(4) == [0.125, 0.25, 0.5]
    it_fine, x_fine, nv = solve('fine')
    monkeypatch.setenv('OCMP_MG_COARSE_H', 'own')
    assert coarse_mesh_size_scales(4) == [1.0, 1.0, 1.0]
    it_own, x_own, _ = solve('own')
    assert it_fine <= it_own and it_fine < 30
    assert np.linalg.norm((x_fine - x_own)[:nv]) < 5e-4 * np.linalg.norm(x_own[:nv])
