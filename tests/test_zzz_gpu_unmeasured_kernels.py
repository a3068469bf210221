"""GPU cases of code paths whose KERNELS have never run on a B200 (written after round 1's GPU budget was spent): the
reduced-precision storage of the multigrid data (k_patch_invert<T, float / bf16>, k_patch_apply_stream<float / bf16>,
k_spmv_f32) and the device-resident mixers (ocmp_mdot / ocmp_maxpy from Python). Collected after everything else, so
that with ``-x`` a surprise here cannot hide the other results. Their host paths run against the null device in
tests/test_gpu_paths_dry.py, their index logic is emulated in tests/test_patch_kernel_emulation.py."""
import numpy as np  # noqa: F401
import pytest

import cases
from test_gpu_parity import _rel, _with
from test_mixing import SCHEMES, _replay

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('scheme', SCHEMES)
def test_mixers_reproduce_reference_vectors_on_device(scheme):
    """Replay of the reference module's vectors with the history on the GPU (ocmp_mdot / ocmp_maxpy); 1e-9 relative
    like every solution field."""
    from opencmp_b200.backend import CudaBackend
    assert _replay(scheme, CudaBackend()) < 1e-9


# ---- patch inverses stored in FP32 (OCMP_PATCH_FP32=1; opt-in, written after the round's last GPU session) -----------
def _stokes_direct(DG):
    def run():
        c = cases.stokes(cases.channel_mesh(10), 2, DG)
        c['gfu'].components[0].Set(c['uex'], definedon=c['mesh'].Boundaries(c['walls']))
        c['a'].Assemble()
        c['L'].Assemble()
        return cases.direct_solve(c)
    return run


@pytest.mark.parametrize('DG', [False, True])
def test_fp32_stored_patch_inverses_keep_the_solution(DG, monkeypatch):
    """GMRES + vertex-patch additive Schwarz with the inverses stored in FP32 (k_patch_invert<T, float>,
    k_patch_apply_stream<float>) vs sparse LU in the oracle: the preconditioner's storage precision must not show in the FP64
    solution (1e-9 relative, like every solution field)."""
    ref = _with('oracle', _stokes_direct(DG))
    monkeypatch.setenv('OCMP_PATCH_FP32', '1')
    got = _with('cuda', _stokes_direct(DG))
    assert _rel(got, ref) < 1e-9


@pytest.mark.parametrize('storage', ['fp32', 'bf16'])
def test_fp32_stored_multigrid_data_keep_iteration_counts(monkeypatch, storage):
    """3-D INS-DIM multigrid step (open-star patches, stride 90 -> 92 in FP32 mode) with FP32-stored patch inverses and
    level matrices inside the cycle: same GMRES iteration counts and velocity as with FP64 storage (checked on the CPU
    restatement beforehand: 27/23 and 47/35 iterations at N = 8 / 16 either way)."""
    def run():
        c = cases.ins_dim_3d(4, preconditioner='multigrid', lam=1.0, nonlinear_max_iterations=2,
                              nonlinear_tolerance=(0.0, 0.0))
        w = c['workload']
        w.step()
        return w.gfu.components[0].vec.NumPy().copy(), list(w.linear_iterations)
    u64, its64 = _with('cuda', run)
    monkeypatch.setenv('OCMP_PATCH_STORAGE', storage)  # open-star stride 90 -> 92 (fp32) / 96 (bf16)
    monkeypatch.setenv('OCMP_SPMV_FP32', '1')        # and FP32 copies of the level matrices inside the cycle
    u32, its32 = _with('cuda', run)
    assert max(abs(a - b) for a, b in zip(its32, its64)) <= (1 if storage == 'fp32' else 4)
    # both are 1e-12-tolerance solves (preconditioned residual) of a system with condition number > 1e10: each lies
    # within ~2e-7 of the sparse-LU solution (tests/test_gpu_parity.py::test_ins_dim_3d_multigrid_step compares at
    # 1e-6), so two differently rounded preconditioners agree to that, not to round-off (bf16: 1.8e-7 measured)
    assert _rel(u32, u64) < 1e-6


@pytest.mark.parametrize('storage', ['fp32', 'bf16'])
def test_fp32_stored_multigrid_data_ins_2d(monkeypatch, storage):
    """2-D INS Taylor-Green step (HDiv-DG order 3, closed vertex patches of 132 DOFs: the two-chunk paths of
    k_patch_apply_stream<float / bf16>): reduced-precision STORAGE of the preconditioner data leaves the iteration
    counts (bf16: within 3) and the FP64 solution unchanged."""
    import opencmp_b200.ngs as ngs
    from opencmp_b200.workloads import INSTaylorGreen

    def run():
        w = INSTaylorGreen(16)
        w.step()
        return ngs.get_backend().to_numpy(w.gfu.vec.a).copy(), list(w.linear_iterations), w.errors()
    u64, its64, e64 = _with('cuda', run)
    monkeypatch.setenv('OCMP_PATCH_STORAGE', storage)
    monkeypatch.setenv('OCMP_SPMV_FP32', '1')
    u32, its32, e32 = _with('cuda', run)
    assert len(its32) == len(its64) and max(abs(a - b) for a, b in zip(its32, its64)) <= (1 if storage == 'fp32' else 3)
    # two GMRES solves of the same FP64 system to a relative preconditioned residual of 1e-10 each
    assert _rel(u32, u64) < 5e-6      # see test_lagged_smoother_gives_the_same_steps (1.2e-6 measured)
    assert abs(e32[0] - e64[0]) < 1e-7
