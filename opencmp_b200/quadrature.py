"""Quadrature rules on the reference cells (host side; tabulated once and uploaded).

Reference cells: segment [0,1]; triangle (0,0),(1,0),(0,1); tetrahedron; quad [0,1]^2; hex [0,1]^3.
Simplex rules are collapsed Gauss-Jacobi (conical) products, exact to the requested degree; tensor cells use
Gauss-Legendre products. The integration order policy mirrors what SURVEY App. A records for NGSolve's symbolic
integrators (2*order for forms, fixed order 5 for ``Integrate``, reference helpers/error.py:66-77).
"""
from __future__ import annotations

from functools import lru_cache
from typing import Tuple

import numpy as np
from scipy.special import roots_jacobi


@lru_cache(maxsize=None)
def gauss_01(n: int) -> Tuple[np.ndarray, np.ndarray]:
    x, w = np.polynomial.legendre.leggauss(n)
    return 0.5 * (x + 1.0), 0.5 * w


@lru_cache(maxsize=None)
def _jacobi_01(n: int, alpha: int) -> Tuple[np.ndarray, np.ndarray]:
    """Nodes/weights on [0,1] for weight (1-x)^alpha."""
    x, w = roots_jacobi(n, float(alpha), 0.0)
    return 0.5 * (x + 1.0), w / 2.0 ** (alpha + 1)


def npts_for_degree(deg: int) -> int:
    return max(1, deg // 2 + 1)


@lru_cache(maxsize=None)
def cell_rule(cell_type: str, deg: int) -> Tuple[np.ndarray, np.ndarray]:
    """Points (nq, dim) and weights (nq,) exact for polynomials of total (simplex) / per-axis (tensor) degree deg."""
    n = npts_for_degree(deg)
    if cell_type == 'seg':
        x, w = gauss_01(n)
        return x[:, None].copy(), w.copy()
    if cell_type == 'tri':
        x0, w0 = gauss_01(n)
        x1, w1 = _jacobi_01(n, 1)
        # xi = u (1-v), eta = v : Jacobian (1-v) absorbed in the Jacobi weight
        U, V = np.meshgrid(x0, x1, indexing='ij')
        W = np.outer(w0, w1)
        pts = np.stack([(U * (1 - V)).ravel(), V.ravel()], axis=1)
        return pts, W.ravel().copy()
    if cell_type == 'tet':
        x0, w0 = gauss_01(n)
        x1, w1 = _jacobi_01(n, 1)
        x2, w2 = _jacobi_01(n, 2)
        U, V, T = np.meshgrid(x0, x1, x2, indexing='ij')
        W = w0[:, None, None] * w1[None, :, None] * w2[None, None, :]
        pts = np.stack([(U * (1 - V) * (1 - T)).ravel(), (V * (1 - T)).ravel(), T.ravel()], axis=1)
        return pts, W.ravel().copy()
    if cell_type == 'quad':
        x, w = gauss_01(n)
        A, B = np.meshgrid(x, x, indexing='ij')
        return np.stack([A.ravel(), B.ravel()], axis=1), np.outer(w, w).ravel().copy()
    if cell_type == 'hex':
        x, w = gauss_01(n)
        A, B, C = np.meshgrid(x, x, x, indexing='ij')
        W = w[:, None, None] * w[None, :, None] * w[None, None, :]
        return np.stack([A.ravel(), B.ravel(), C.ravel()], axis=1), W.ravel().copy()
    raise ValueError(cell_type)


def facet_type(cell_type: str) -> str:
    return {'tri': 'seg', 'quad': 'seg', 'tet': 'tri', 'hex': 'quad'}[cell_type]


def facet_rule_in_cell(cell_type: str, deg: int) -> Tuple[np.ndarray, np.ndarray]:
    """Facet rule mapped into the reference cell for every local facet.

    Returns pts (nfc, nqf, dim) and weights (nqf,). Because local facet vertices are listed in ascending order
    (mesh.py), point q is the same physical point when seen from either neighbouring cell.
    """
    from .mesh import local_topology
    loc = local_topology(cell_type)
    ref = loc['ref']
    fp, fw = cell_rule(facet_type(cell_type), deg)
    out = []
    for fv in loc['facets']:
        v = ref[list(fv)]
        if cell_type in ('tri', 'quad'):
            out.append(v[0][None, :] + fp[:, [0]] * (v[1] - v[0])[None, :])
        elif cell_type == 'tet':
            out.append(v[0][None, :] + fp[:, [0]] * (v[1] - v[0])[None, :] + fp[:, [1]] * (v[2] - v[0])[None, :])
        else:
            out.append(v[0][None, :] + fp[:, [0]] * (v[1] - v[0])[None, :] + fp[:, [1]] * (v[2] - v[0])[None, :])
    return np.stack(out, axis=0), fw


def facet_ref_geometry(cell_type: str):
    """Per local facet: reference tangents (nfc, dim-1, dim) and outward reference normals (nfc, dim)."""
    from .mesh import local_topology
    loc = local_topology(cell_type)
    ref = loc['ref']
    dim = loc['dim']
    centre = ref.mean(axis=0)
    tang, nrm = [], []
    for fv in loc['facets']:
        v = ref[list(fv)]
        t = np.stack([v[k + 1] - v[0] for k in range(dim - 1)], axis=0)
        if dim == 2:
            n = np.array([t[0, 1], -t[0, 0]])
        else:
            n = np.cross(t[0], t[1])
        if np.dot(n, v.mean(axis=0) - centre) < 0:
            n = -n
        tang.append(t)
        nrm.append(n / np.linalg.norm(n))
    return np.stack(tang, axis=0), np.stack(nrm, axis=0)
