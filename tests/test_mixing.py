"""Device-resident nonlinear mixers (opencmp_b200/mixing.py) against golden vectors produced by the
unmodified reference module opencmp/solvers/nonlinear_mixing.py (tests/golden/make_mixing_golden.py)."""
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'mixing_reference.npz')
SCHEMES = ('LinearMixing', 'DiagBroyden', 'Anderson')


def _replay(scheme, backend):
    import opencmp_b200.ngs as ngs
    from opencmp_b200 import mixing
    old = ngs._backend
    ngs.set_backend(backend)
    try:
        g = np.load(GOLD)
        mixer = mixing.make_mixer(scheme, keep_vectors=3) if scheme == 'Anderson' else mixing.make_mixer(scheme)
        worst = 0.0
        for k, (f, x, dx_ref) in enumerate(zip(g[scheme + '_f'], g[scheme + '_x'], g[scheme + '_dx'])):
            fv = ngs.BaseVector(backend.from_numpy(f))
            xv = ngs.BaseVector(backend.from_numpy(x))
            dx = mixer.step(fv, xv, k + 2)
            # the value must be assignable the way base_solver.py:686 does it
            out = ngs.BaseVector(backend.zeros(len(f)))
            out.data = dx
            got = backend.to_numpy(out.a)
            worst = max(worst, np.abs(got - dx_ref).max() / np.abs(dx_ref).max())
        return worst
    finally:
        ngs.set_backend(old)


@pytest.mark.parametrize('scheme', SCHEMES)
def test_mixers_reproduce_reference_vectors_host_arrays(scheme):
    from oracle.backend import OracleBackend
    assert _replay(scheme, OracleBackend()) < 1e-10


def test_unknown_scheme_raises_like_the_reference():
    from opencmp_b200 import mixing
    with pytest.raises(ValueError):
        mixing.make_mixer('Newton')


def test_reference_solver_picks_up_device_mixing():
    """With install_as_ngsolve(device_mixing=True) the unmodified reference imports our module in place of its own."""
    if not os.path.isdir('/root/reference/opencmp'):
        pytest.skip('reference tree not mounted')
    import subprocess
    import sys
    code = ("import sys; sys.path.insert(0, %r); sys.path.insert(0, '/root/reference');"
            "import opencmp_b200.compat as c; c.install_as_ngsolve(device_mixing=True);"
            "import opencmp.solvers.base_solver as b; import opencmp_b200.mixing as m;"
            "assert b.make_mixer is m.make_mixer; print('ok')") % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and 'ok' in r.stdout, r.stderr[-2000:]
