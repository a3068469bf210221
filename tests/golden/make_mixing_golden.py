"""Golden vectors for the nonlinear mixers, produced by the UNMODIFIED reference module
``/root/reference/opencmp/solvers/nonlinear_mixing.py`` (imported through the NGSolve alias; it only uses BaseVector).

A deterministic fixed-point problem x = G(x) (n = 60) is iterated exactly like the stationary branch of
``Solver._solve`` (base_solver.py:678-690): f = G(x) - x_prev, dx = mixer.step(f, x_prev, it), x += dx. For every
scheme the inputs (f, x_prev) and the returned dx of each iteration are stored.

Run in the build container (where /root/reference is mounted):  python tests/golden/make_mixing_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import opencmp_b200.compat as compat          # noqa: E402
import opencmp_b200.ngs as ngs                # noqa: E402
from oracle.backend import OracleBackend      # noqa: E402

compat.install_as_ngsolve()
ngs.set_backend(OracleBackend())
import importlib.util                         # noqa: E402

spec = importlib.util.spec_from_file_location('ref_nonlinear_mixing',
                                              '/root/reference/opencmp/solvers/nonlinear_mixing.py')
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)


def problem(n=60, seed=7):
    rng = np.random.default_rng(seed)
    B = rng.standard_normal((n, n))
    B *= 0.55 / np.linalg.norm(B, 2)
    c = rng.standard_normal(n)
    return lambda x: B @ np.tanh(x) + 0.3 * np.cos(x) + c


def main():
    G = problem()
    out = {}
    for scheme in ('LinearMixing', 'DiagBroyden', 'Anderson'):
        mixer = ref.make_mixer(scheme, keep_vectors=3) if scheme == 'Anderson' else ref.make_mixer(scheme)
        x = np.zeros(60)
        x_prev = None
        fs, xs, dxs = [], [], []
        for it in range(1, 13):
            g = G(x)                                   # "linearized_solve" result
            if it == 1:
                x_prev = g.copy()
                x = g.copy()
                continue
            f = g - x_prev
            dx = mixer.step(ngs.BaseVector(f.copy()), ngs.BaseVector(x_prev.copy()), it)
            fs.append(f.copy())
            xs.append(x_prev.copy())
            dxs.append(np.array(dx, dtype=np.float64).copy())
            x = x + dx
            x_prev = x.copy()
        out[scheme + '_f'] = np.array(fs)
        out[scheme + '_x'] = np.array(xs)
        out[scheme + '_dx'] = np.array(dxs)
        out[scheme + '_resid'] = np.array([np.linalg.norm(G(x) - x)])
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'mixing_reference.npz'), **out)
    for k in ('LinearMixing', 'DiagBroyden', 'Anderson'):
        print(k, 'final fixed-point residual', out[k + '_resid'][0], '|dx| per iteration',
              np.linalg.norm(out[k + '_dx'], axis=1).round(6))


if __name__ == '__main__':
    main()
