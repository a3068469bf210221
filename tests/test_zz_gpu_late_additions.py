"""GPU cases added after the last GPU session of round 1 (the round's GPU budget was spent): Raviart-Thomas and
quadrilateral HDiv elements, and the device-resident nonlinear mixers. They use only kernels whose other cases are
green, but have not run on a B200 yet — kept in this file, which pytest collects last, so that with ``-x`` a surprise
here cannot hide the established parity results."""
import numpy as np
import pytest

import cases
from test_gpu_parity import MAT_TOL, _assembled, _rel, _with
from test_mixing import SCHEMES, _replay

pytestmark = pytest.mark.gpu

CASES = {
    # HDiv(RT=True), reference models/ins.py:114-117
    'stokes_hdiv_rt_p2_oseen': lambda: cases.stokes(cases.channel_mesh(), 2, True, RT=True, dt_val=0.01, mass=True,
                                                     wind=lambda n: cases.random_wind(n)),
    # HDiv-DG on quadrilaterals (RT_[k]): the element of the reference's DIM Stokes / INS models on its default
    # quadrilateral DIM meshes (pytests/full_system/dim/dim_stokes_1)
    'stokes_hdiv_dg_quad_p2_oseen': lambda: cases.stokes(cases.quad_channel_mesh(), 2, True, dt_val=0.01, mass=True,
                                                          wind=lambda n: cases.random_wind(n)),
    'stokes_hdiv_dg_quad_p3': lambda: cases.stokes(cases.quad_channel_mesh(), 3, True),
}


@pytest.mark.parametrize('name', sorted(CASES))
def test_assembly_matches_oracle(name):
    ref = _with('oracle', _assembled(CASES[name]))
    got = _with('cuda', _assembled(CASES[name]))
    assert got['vals'].shape == ref['vals'].shape
    assert _rel(got['bc'], ref['bc']) < 1e-11 or np.abs(ref['bc']).max() == 0
    assert _rel(got['vals'], ref['vals']) < MAT_TOL
    assert _rel(got['rhs'], ref['rhs']) < MAT_TOL
    assert _rel(got['y'], ref['y']) < 1e-12


@pytest.mark.parametrize('scheme', SCHEMES)
def test_mixers_reproduce_reference_vectors_on_device(scheme):
    """Replay of the reference module's vectors with the history on the GPU (ocmp_mdot / ocmp_maxpy); 1e-9 relative
    like every solution field."""
    from opencmp_b200.backend import CudaBackend
    assert _replay(scheme, CudaBackend()) < 1e-9
