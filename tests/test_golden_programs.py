"""Replay the weak forms that the reference's OWN model classes built (INS Oseen/IMEX, MultiComponentINS, Poisson DG,
stationary INS with stress boundaries — lowered in the build container by tests/golden/make_golden.py) and compare
the assembled operator and right-hand side with the values the oracle produced there:

  * CPU: serialisation round trip (oracle backend) — the fixture is self-consistent;
  * GPU: the CUDA kernels assemble exactly the reference models' forms: A x for three seeded vectors, the diagonal and
    the right-hand side agree to 1e-12 of the largest entry (north-star tolerance for matrix entries)."""
import glob
import os

import numpy as np
import pytest

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
FIXTURES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(HERE, 'prog_*.npz')))
# fixtures added after the round's last GPU session: their GPU replay lives in test_zz_gpu_late_additions.py
LATE = {'prog_ins_dim_dg_quad', 'prog_poisson_dim_2'}


def _replay(ngs, name):
    from opencmp_b200.serialize import load
    mesh, progs, extra = load(os.path.join(HERE, name + '.npz'), ngs)
    be = ngs.get_backend()
    a, L = progs['a'], progs['L']
    mat = ngs.Matrix(a.fes)
    be.assemble_matrix(a, mat)
    rhs = be.zeros(a.fes.ndof)
    be.assemble_vector(L, rhs)
    Y = []
    for x in extra['X']:
        y = be.zeros(a.fes.ndof)
        be.spmv(mat, be.from_numpy(x), y)
        Y.append(be.to_numpy(y).copy())
    vals = be.to_numpy(mat.values)
    diag = vals[a.fes.pattern().diag]
    return np.array(Y), be.to_numpy(rhs).copy(), diag, extra


def _check(Y, rhs, diag, extra, tol):
    scale = float(extra['absmax'])
    n = Y.shape[1]
    assert np.abs(Y - extra['Y']).max() < tol * scale * np.sqrt(n)
    assert np.abs(diag - extra['diag']).max() < tol * scale
    assert np.abs(rhs - extra['rhs']).max() < tol * max(np.abs(extra['rhs']).max(), 1e-300)


def test_fixture_list_is_complete():
    assert len(FIXTURES) >= 6


@pytest.mark.parametrize('name', FIXTURES)
def test_round_trip_on_oracle(oracle_backend, name):
    _check(*_replay(oracle_backend, name), tol=1e-13)


@pytest.mark.gpu
@pytest.mark.parametrize('name', [f for f in FIXTURES if f not in LATE])
def test_gpu_assembles_reference_model_forms(cuda_backend, name):
    _check(*_replay(cuda_backend, name), tol=1e-12)
