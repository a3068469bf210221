"""GPU: the Krylov drivers of ocmp_krylov that have no other solver-level case — MINRES (kind 3; reference
opencmp/models/base_model.py:929-932 dispatches ``linear_solver = MinRes`` to ngsolve.solvers.MinRes) and the band LU as
the 'direct' preconditioner (base_model.py:365-383) — against the oracle's restatement of the same algorithms."""
import numpy as np
import pytest

import cases
from test_gpu_parity import _with, _rel

pytestmark = pytest.mark.gpu


def _poisson_minres(n=10, order=2):
    def run():
        c = cases.poisson(cases.square_mesh(n), order, False)
        ngs = c['ngs']
        c['gfu'].components[0].Set(c['exact'], definedon=c['mesh'].Boundaries(c['dnames']))
        c['a'].Assemble()
        c['L'].Assemble()
        pre = ngs.Preconditioner(c['a'], 'local')
        pre.Update()
        r = c['L'].vec.CreateVector()
        r.data = c['L'].vec - c['a'].mat * c['gfu'].vec          # MinRes starts from zero (initialize=True)
        dx = ngs.solvers.MinRes(mat=c['a'].mat, rhs=r, pre=pre, tol=1e-13, maxsteps=2000)
        be = ngs.get_backend()
        if hasattr(be, 'krylov_history'):
            hist = be.krylov_history()
        else:
            import oracle.backend as ob
            hist = list(ob.last_history)
        sol = (c['gfu'].vec.NumPy() + dx.NumPy()).copy()
        return sol, np.array(hist)
    return run


def test_minres_iterates_match_oracle():
    ref, href = _with('oracle', _poisson_minres())
    got, hgot = _with('cuda', _poisson_minres())
    assert _rel(got, ref) < 1e-9
    assert abs(len(hgot) - len(href)) <= 2
    k = min(len(hgot), len(href), 40)
    assert k >= 20
    # the residual ratios of the first iterations agree digit for digit (same recurrence, FP64 both sides)
    assert np.allclose(hgot[:k], href[:k], rtol=1e-6, atol=0)


def test_minres_equals_direct_solution():
    def direct():
        c = cases.poisson(cases.square_mesh(10), 2, False)
        c['gfu'].components[0].Set(c['exact'], definedon=c['mesh'].Boundaries(c['dnames']))
        c['a'].Assemble()
        c['L'].Assemble()
        return cases.direct_solve(c)
    ref = _with('oracle', direct)
    got, _ = _with('cuda', _poisson_minres())
    assert _rel(got, ref) < 1e-9


def test_direct_preconditioner_converges_in_one_gmres_step():
    """ngs.Preconditioner(a, 'direct') + GMRes: the band LU applied inside ocmp_krylov (pre_kind 5)."""
    def run():
        c = cases.stokes(cases.channel_mesh(8), 2, False)
        ngs = c['ngs']
        c['gfu'].components[0].Set(c['uex'], definedon=c['mesh'].Boundaries(c['walls']))
        c['a'].Assemble()
        c['L'].Assemble()
        pre = ngs.Preconditioner(c['a'], 'direct')
        pre.Update()
        ngs.solvers.GMRes(A=c['a'].mat, b=c['L'].vec, pre=pre, freedofs=c['fes'].FreeDofs(), x=c['gfu'].vec,
                          tol=1e-12, maxsteps=20)
        return c['gfu'].vec.NumPy().copy(), getattr(ngs.get_backend(), 'last_iters', 1)
    ref, _ = _with('oracle', run)
    got, its = _with('cuda', run)
    assert its <= 3
    assert _rel(got, ref) < 1e-9


def test_unimplemented_preconditioner_types_raise():
    def run():
        c = cases.poisson(cases.square_mesh(4), 1, False)
        ngs = c['ngs']
        c['a'].Assemble()
        for kind in ('h1amg', 'bddc'):
            try:
                ngs.Preconditioner(c['a'], kind).Update()
            except NotImplementedError as exc:
                assert kind in str(exc)
            else:
                assert False, kind + ' must not be silently replaced by another preconditioner'
    _with('cuda', run)


@pytest.mark.parametrize('name', ['stokes_3d_hex_q2q1', 'stokes_th_p2_oseen', 'ins_dim_3d_hex_q2q1'])
def test_component_run_spmv_equals_plain_csr_product(name):
    """The column-index compression of vector-valued spaces (ocmp_spmv_runs / ocmp_spmv_compressed): every row of a
    Taylor-Hood pattern is recognised, and the product equals the plain CSR product (and the oracle's) to round-off."""
    from test_gpu_parity import CASES

    def run():
        c = CASES[name]()
        ngs = c['ngs']
        c['a'].Assemble()
        be = ngs.get_backend()
        x = np.random.default_rng(4).uniform(-1, 1, c['fes'].ndof)
        xv = ngs.BaseVector(be.from_numpy(x))
        y_plain = (c['a'].mat * xv).NumPy().copy()
        out = [y_plain]
        if be.name == 'cuda':
            pd = be.pattern_data(c['fes'])
            assert pd['runs'] is not None
            runlen, shift, nc, grouped = pd['runs']
            assert int((runlen > 0).sum()) == c['fes'].ndof          # all rows carry the component runs
            assert grouped                                             # ... and the components of a node share columns
            for g in (0, 1):                                           # row-wise and node-grouped kernels
                y = be.zeros(c['fes'].ndof)
                be._ck(be.lib.ocmp_spmv_compressed(c['fes'].ndof, pd['rowptr'].data_ptr(), pd['colidx'].data_ptr(),
                                                   c['a'].mat.values.data_ptr(), runlen.data_ptr(), shift, nc, g,
                                                   xv.a.data_ptr(), y.data_ptr(), be._stream()))
                out.append(be.to_numpy(y))
        return out
    ref = _with('oracle', run)
    got = _with('cuda', run)
    assert _rel(got[0], ref[0]) < 1e-12                 # `mat * x` itself goes through the grouped kernel
    assert _rel(got[1], ref[0]) < 1e-12 and _rel(got[2], ref[0]) < 1e-12
