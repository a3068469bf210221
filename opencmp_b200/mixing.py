"""Stationary nonlinear mixing on the device.

Drop-in for reference ``opencmp/solvers/nonlinear_mixing.py:19-141`` (``make_mixer`` and the three schemes the
stationary branch of ``Solver._solve`` calls, ``base_solver.py:636,678-690``): same constructor arguments, same
``step(f_vec, x_prev_vec, num_iterations)`` call, same first-step convention (iteration 2 = linear mixing with the
dominant-eigenvalue estimate of alpha), same arithmetic — but the DOF vectors never leave the device. The reference
pulls ``f_vec.FV().NumPy()`` to the host every nonlinear iteration and keeps its history there; here the history lives
in two device arrays, the Gram matrix of Anderson mixing is extended by one batched multi-dot per iteration
(``ocmp_mdot``), the update is one batched multi-axpy (``ocmp_maxpy``), and only the (<= 6 x 6) least-squares system is
solved on the host. ``step`` returns a backend array that ``dx.data = ...`` accepts.

``compat.install_as_ngsolve(device_mixing=True)`` registers this module as ``opencmp.solvers.nonlinear_mixing`` so the
unmodified reference picks it up.
"""
from __future__ import annotations

import numpy as np
import numpy.linalg as npl

from . import ngs


def _be():
    return ngs.get_backend()


def _first_alpha(f_vec, x_prev_vec) -> float:
    """nonlinear_mixing.py:31 / :50 / :86"""
    return min(1.0, 0.5 * max(1.0, x_prev_vec.Norm()) / f_vec.Norm())


class LinearMixing:
    """dx = alpha * f (nonlinear_mixing.py:19-32)."""

    def __init__(self, alpha: float = 1.0, **_) -> None:
        self._alpha = alpha

    def step(self, f_vec, x_prev_vec, num_iterations: int):
        if num_iterations == 2:
            self._alpha = _first_alpha(f_vec, x_prev_vec)
        return self._alpha * f_vec.a


class DiagBroyden:
    """Diagonal Broyden update (nonlinear_mixing.py:35-61); beta is a device vector."""

    def __init__(self, alpha: float = 1.0, **_) -> None:
        self._alpha = alpha
        self._fprev = None
        self._beta = None
        self._dx_prev = None

    def step(self, f_vec, x_prev_vec, num_iterations: int):
        fcurr = f_vec.a
        if num_iterations == 2:
            self._alpha = _first_alpha(f_vec, x_prev_vec)
            self._fprev = fcurr + 0.0
            self._beta = fcurr * 0.0 + 1.0 / self._alpha
            dx = self._alpha * fcurr
        else:
            dxp = self._dx_prev
            nrm2 = float(_be().dot(dxp, dxp))
            self._beta = self._beta - (fcurr - self._fprev + self._beta * dxp) * dxp / nrm2
            dx = fcurr / self._beta
            self._fprev = fcurr + 0.0
        self._dx_prev = dx + 0.0
        return dx


class Anderson:
    """Anderson mixing, Eyert J. Comp. Phys. 124, 271 (1996) (nonlinear_mixing.py:64-121). History = rows of two
    device arrays; ``_slots`` lists the rows in chronological order like the reference's Python lists."""

    def __init__(self, alpha: float = 1.0, keep_vectors: int = 5, w0: float = 0.01,
                 singular_tolerance: float = 1e-6) -> None:
        self._alpha = alpha
        self._keep = keep_vectors
        self._w0 = w0
        self._tol = singular_tolerance
        self._fprev = None
        self._dx = self._df = None          # (keep + 2, n) device arrays
        self._dx_slots: list = []
        self._df_slots: list = []
        self._gram = {}                     # (df slot i, df slot j) -> <df_i, df_j>, i <= j chronologically

    def _alloc(self, like):
        n = like.shape[0]
        rows = self._keep + 2
        self._dx = _be().zeros(rows * n).reshape(rows, n)
        self._df = _be().zeros(rows * n).reshape(rows, n)

    def _free_row(self, used) -> int:
        return next(r for r in range(self._keep + 2) if r not in used)

    def _push(self, which: str, vec) -> int:
        slots = self._dx_slots if which == 'dx' else self._df_slots
        arr = self._dx if which == 'dx' else self._df
        r = self._free_row(slots)
        _be().copy_into(arr[r], vec)
        slots.append(r)
        return r

    def step(self, f_vec, x_prev_vec, num_iterations: int):
        be = _be()
        fcurr = f_vec.a
        if self._dx is None:
            self._alloc(fcurr)
        if num_iterations == 2:
            self._alpha = _first_alpha(f_vec, x_prev_vec)
            self._fprev = fcurr + 0.0
            dx = self._alpha * fcurr
            self._dx_slots, self._df_slots, self._gram = [], [], {}
            self._push('dx', dx)
            return dx
        # num_iterations > 2: full Anderson mixing
        new = self._push('df', fcurr - self._fprev)
        be.copy_into(self._fprev, fcurr)
        if len(self._dx_slots) > self._keep:
            self._dx_slots.pop(0)
            old = self._df_slots.pop(0)
            self._gram = {k: v for k, v in self._gram.items() if old not in k}
        # one batched launch: the new difference vector against every stored one, and f against every stored one
        dots_new = be.mdot(self._df, self._df[new])
        for r in self._df_slots:
            self._gram[(r, new)] = float(dots_new[r])
        m = len(self._dx_slots)
        A = np.zeros((m, m))
        for i in range(m):
            for j in range(i, m):
                A[i, j] = self._gram[(self._df_slots[i], self._df_slots[j])]
        np.fill_diagonal(A, A.diagonal() * (1 + self._w0 ** 2))
        A += np.triu(A, 1).T.conj()
        if abs(npl.det(A)) < self._tol:
            # reset the Jacobian approximation (nonlinear_mixing.py:106-111)
            self._dx_slots, self._df_slots, self._gram = [], [], {}
            dx = self._alpha * fcurr
            self._push('dx', dx)
            return dx
        dff_all = be.mdot(self._df, fcurr)
        dff = np.array([dff_all[r] for r in self._df_slots])
        gamma = npl.solve(A, dff)
        # dx = alpha f - sum_k gamma_k (dx_k + alpha df_k)
        dx = self._alpha * fcurr
        cdx = np.zeros(self._keep + 2)
        cdf = np.zeros(self._keep + 2)
        for k in range(len(self._df_slots)):
            cdx[self._dx_slots[k]] -= gamma[k]
            cdf[self._df_slots[k]] -= gamma[k] * self._alpha
        be.maxpy(self._dx, cdx, dx)
        be.maxpy(self._df, cdf, dx)
        self._push('dx', dx)
        return dx


_SCHEMES = {'LinearMixing': LinearMixing, 'DiagBroyden': DiagBroyden, 'Anderson': Anderson, 'default': Anderson}


def make_mixer(name: str, alpha: float = 1.0, keep_vectors: int = 5, w0: float = 0.01,
               singular_tolerance: float = 1e-6):
    """nonlinear_mixing.py:131-141"""
    if name not in _SCHEMES:
        raise ValueError(f"Unknown nonlinear_solver '{name}'. Choose from: {list(_SCHEMES)}")
    return _SCHEMES[name](alpha=alpha, keep_vectors=keep_vectors, w0=w0, singular_tolerance=singular_tolerance)
