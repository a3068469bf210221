"""GPU cases of the matrix-free operator (tests/test_matrix_free.py), written in the round's last session, after its GPU
budget was spent: never run on a B200. Collected after every other GPU file, so that with ``-x`` a surprise here cannot
hide the results of the established cases or of the FP32 / bf16 kernels."""
import pytest

from test_gpu_parity import _rel, _with

pytestmark = pytest.mark.gpu


# ---- k_coef + k_lin with the trial rows as field slots ----------------------------------------------------------
from test_matrix_free import CASES as MATRIX_FREE_CASES, matrix_free_vs_csr  # noqa: E402


@pytest.mark.parametrize('name', sorted(MATRIX_FREE_CASES))
def test_matrix_free_apply_matches_oracle_and_csr(name):
    """``BilinearForm.Apply`` / ``BilinearForm(nonassemble=True).mat * x`` on the GPU: equal to the oracle's A x (1e-12)
    and to the GPU's own CSR SpMV of the assembled matrix."""
    ref_csr, ref_free, _ = _with('oracle', lambda: matrix_free_vs_csr(MATRIX_FREE_CASES[name]()))
    got_csr, got_free, got_twin = _with('cuda', lambda: matrix_free_vs_csr(MATRIX_FREE_CASES[name]()))
    assert _rel(got_free, ref_free) < 1e-12
    assert _rel(got_free, got_csr) < 1e-12
    assert _rel(got_twin, ref_csr) < 1e-12


def test_krylov_on_the_matrix_free_operator_matches_oracle():
    """CG / GMRES inside ``ocmp_krylov`` with ``ocmp_system.apply_fn`` set (the form's action called back per
    iteration) against the oracle and against the stored-operator solves."""
    from test_matrix_free import krylov_matrix_free_vs_csr
    ref = _with('oracle', krylov_matrix_free_vs_csr)
    got = _with('cuda', krylov_matrix_free_vs_csr)
    for r, g in zip(ref, got):
        assert _rel(g, r) < 1e-9
    assert _rel(got[1], got[0]) < 1e-9 and _rel(got[3], got[2]) < 1e-9
