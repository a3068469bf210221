"""Matrix-free operator application (north star item 3: "a matrix-free and CSR SpMV"): ``BilinearForm.Apply(x, y)`` and
``BilinearForm(fes, nonassemble=True).mat * x`` — NGSolve's interface for it — lower the form with the trial function
replaced by the DOF vector (``symbolic.form_action``) and run it through the linear-form kernels. Checked here on the
CPU restatement against the assembled CSR matrix times the same vector, over CG, DG (interior facets with ``.Other()``),
HDiv, Oseen (wind field + ``IfPos`` / ``Norm`` coefficients), phi-weighted DIM and 3-D forms; the GPU cases are in
``tests/test_zzzz_gpu_matrix_free.py`` and run dry in ``tests/test_gpu_paths_dry.py``."""
import numpy as np
import pytest

import cases
import opencmp_b200.ngs as ngs
from oracle.backend import OracleBackend

CASES = {
    'poisson_h1_p3': lambda: cases.poisson(cases.square_mesh(), 3, False),
    'poisson_l2_dg_p2': lambda: cases.poisson(cases.square_mesh(), 2, True, family='L2'),
    'stokes_th_p2_oseen': lambda: cases.stokes(cases.channel_mesh(), 2, False, wind=lambda n: cases.random_wind(n)),
    'ins_hdiv_dg_p3_oseen': lambda: cases.stokes(cases.channel_mesh(8), 3, True, wind=lambda n: cases.random_wind(n),
                                                  dt_val=0.01, mass=True),
    'poisson_dim_h1_p2_dg_tri': lambda: cases.poisson_dim(cases.square_mesh(6), 2, True),
    'species_dg_p2': lambda: cases.species(cases.square_mesh(5), 2, lambda n: cases.random_wind(n, 11)),
    'stokes_3d_hex_q2q1': lambda: cases.stokes_3d('hex', 2),
}


@pytest.fixture
def oracle_backend():
    old = ngs._backend
    ngs.set_backend(OracleBackend())
    yield
    ngs.set_backend(old)


def matrix_free_vs_csr(c, seed=0):
    """(A x) through the stored matrix and through ``Apply`` for a seeded x; also through a ``nonassemble`` twin."""
    a = c['a']
    a.Assemble()
    x = np.random.default_rng(seed).uniform(-1, 1, a.space.ndof)
    xv = ngs.BaseVector(ngs.get_backend().from_numpy(x))
    y_csr = (a.mat * xv).NumPy().copy()
    yv = a.mat.CreateColVector()
    a.Apply(xv, yv)
    y_free = yv.NumPy().copy()
    twin = ngs.BilinearForm(a.space, nonassemble=True)
    twin += a.integrals
    twin.Assemble()                                           # a no-op, as in NGSolve
    y_twin = (twin.mat * xv).NumPy().copy()
    return y_csr, y_free, y_twin


@pytest.mark.parametrize('name', sorted(CASES))
def test_apply_equals_assembled_matrix(oracle_backend, name):
    y_csr, y_free, y_twin = matrix_free_vs_csr(CASES[name]())
    scale = np.abs(y_csr).max()
    assert scale > 0
    assert np.abs(y_free - y_csr).max() <= 1e-12 * scale
    assert np.array_equal(y_twin, y_free)


def test_apply_follows_parameters_and_fields(oracle_backend):
    """Like ``Assemble()``, ``Apply`` reads the *current* Parameter and coefficient-field values (the Oseen wind changes
    between Picard iterations, dt between adaptive steps): no re-lowering, the program is cached per form."""
    c = cases.stokes(cases.channel_mesh(), 2, False, wind=lambda n: cases.random_wind(n), dt_val=0.01, mass=True)
    y0 = matrix_free_vs_csr(c)[1]
    prog = c['a']._action[2]
    c['params'][0].Set(0.02)
    c['W'].vec.data = ngs.BaseVector(ngs.get_backend().from_numpy(cases.random_wind(c['V'].ndof, seed=9)))
    y_csr, y_free, _ = matrix_free_vs_csr(c)
    assert c['a']._action[2] is prog
    assert np.abs(y_free - y0).max() > 1e-3 * np.abs(y0).max()
    assert np.abs(y_free - y_csr).max() <= 1e-12 * np.abs(y_csr).max()


def test_nonassemble_form_stores_nothing(oracle_backend):
    c = cases.poisson(cases.square_mesh(), 2, False)
    twin = ngs.BilinearForm(c['fes'], nonassemble=True)
    twin += c['a'].integrals
    assert twin.Assemble() is twin and not hasattr(twin.mat, 'values')
    with pytest.raises(RuntimeError):
        twin.mat.Inverse(c['fes'].FreeDofs())


def test_action_rejects_integrands_without_trial_function(oracle_backend):
    c = cases.poisson(cases.square_mesh(), 1, False)
    v = c['fes'].TestFunction()[0]
    bad = ngs.BilinearForm(c['fes'], nonassemble=True)
    bad += v * ngs.dx
    x = bad.mat.CreateColVector()
    with pytest.raises(ValueError):
        bad.mat * x


def krylov_matrix_free_vs_csr():
    """CG and GMRES solutions of a Poisson problem with the stored and with the matrix-free operator (Jacobi from the
    assembled form): (cg_csr, cg_free, gmres_csr, gmres_free)."""
    c = cases.poisson(cases.square_mesh(), 2, False)
    a, L, fes = c['a'], c['L'], c['fes']
    a.Assemble()
    L.Assemble()
    twin = ngs.BilinearForm(fes, nonassemble=True)
    twin += a.integrals
    pre = ngs.Preconditioner(a, 'local')
    pre.Update()
    free = fes.FreeDofs()
    be = ngs.get_backend()
    r = np.array(be.to_numpy(L.vec.a), dtype=np.float64)
    r[~np.asarray(free, bool)] = 0.0
    rhs = ngs.BaseVector(be.from_numpy(r))
    out = []
    for op in (a.mat, twin.mat):
        out.append(ngs.solvers.CG(mat=op, rhs=rhs, pre=pre, tol=1e-12, maxsteps=500).NumPy().copy())
    for op in (a.mat, twin.mat):
        out.append(ngs.solvers.GMRes(A=op, b=rhs, pre=pre, freedofs=free, tol=1e-12, maxsteps=300).NumPy().copy())
    return out[0], out[1], out[2], out[3]


def test_krylov_on_the_matrix_free_operator(oracle_backend):
    """``solvers.CG(mat=BilinearForm(nonassemble=True).mat, pre=Preconditioner(assembled form))`` and ``GMRes`` with it:
    the Krylov operator is the form's action (``ocmp_system.apply_fn`` on the GPU), the preconditioner is built from a
    stored matrix as usual. Same iterates as with the assembled operator."""
    u_csr, u_free, g_csr, g_free = krylov_matrix_free_vs_csr()
    assert np.abs(u_csr).max() > 0
    assert np.abs(u_free - u_csr).max() <= 1e-10 * np.abs(u_csr).max()
    assert np.abs(g_free - g_csr).max() <= 1e-10 * np.abs(g_csr).max()
    assert np.abs(g_csr - u_csr).max() <= 1e-8 * np.abs(u_csr).max()
