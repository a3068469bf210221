"""Reference finite-element bases, tabulated on the host once per (cell type, space, order, rule).

Every cell of a mesh shares one reference orientation (mesh.py sorts simplex vertices; structured tensor cells are
axis-consistent), so a single table serves all cells; the CUDA kernels apply the per-cell affine / Piola map.

Families (the spaces OpenCMP requests through ``getattr(ngs, name)``, reference models/poisson.py:60-65,
models/ins.py:95-128, models/multi_component_ins.py:79-116):

* H1 order p  — hierarchical: vertex hats, edge functions lam_a lam_b P^s_k(lam_b - lam_a, lam_a + lam_b), cell bubbles
  (SURVEY App. A records this as NGSolve's construction; DOF order lowest-order first).
* L2 order p  — orthogonal Dubiner (simplex) / Legendre tensor (quad, hex) modes, constant first.
* HDiv order k (BDM_k = [P_k]^d, triangles) — basis dual to edge normal moments against Legendre polynomials
  (degree 0 first) and interior moments; contravariant Piola map with signed det(J).

A basis reports, for the global numbering in space.py, how its local dofs attach to mesh entities.
"""
from __future__ import annotations

from functools import lru_cache
from typing import List, Tuple

import numpy as np
from scipy.special import eval_jacobi, eval_legendre, jacobi, legendre

from .mesh import local_topology
from .quadrature import cell_rule, gauss_01


# ---------------------------------------------------------------------------------------------------------------
# small dense multivariate polynomial helper (monomial coefficients), used only to build simplex bases
# ---------------------------------------------------------------------------------------------------------------
class Poly:
    __slots__ = ('c',)

    def __init__(self, c):
        self.c = np.asarray(c, dtype=np.float64)

    @staticmethod
    def const(d: int, v: float) -> 'Poly':
        return Poly(np.full((1,) * d, float(v)))

    @staticmethod
    def var(d: int, axis: int) -> 'Poly':
        shape = [1] * d
        shape[axis] = 2
        c = np.zeros(shape)
        idx = [0] * d
        idx[axis] = 1
        c[tuple(idx)] = 1.0
        return Poly(c)

    def _pad(self, shape):
        out = np.zeros(shape)
        out[tuple(slice(0, s) for s in self.c.shape)] = self.c
        return out

    def __add__(self, o):
        if not isinstance(o, Poly):
            o = Poly.const(self.c.ndim, o)
        shape = tuple(max(a, b) for a, b in zip(self.c.shape, o.c.shape))
        return Poly(self._pad(shape) + o._pad(shape))

    __radd__ = __add__

    def __neg__(self):
        return Poly(-self.c)

    def __sub__(self, o):
        return self + (-o if isinstance(o, Poly) else -float(o))

    def __rsub__(self, o):
        return (-self) + o

    def __mul__(self, o):
        if not isinstance(o, Poly):
            return Poly(self.c * float(o))
        a, b = self.c, o.c
        shape = tuple(x + y - 1 for x, y in zip(a.shape, b.shape))
        out = np.zeros(shape)
        for idx in np.ndindex(*a.shape):
            if a[idx] != 0.0:
                out[tuple(slice(i, i + s) for i, s in zip(idx, b.shape))] += a[idx] * b
        return Poly(out)

    __rmul__ = __mul__

    def diff(self, axis: int) -> 'Poly':
        c = self.c
        n = c.shape[axis]
        if n == 1:
            return Poly(np.zeros((1,) * c.ndim))
        k = np.arange(1, n).reshape([-1 if a == axis else 1 for a in range(c.ndim)])
        sl = [slice(None)] * c.ndim
        sl[axis] = slice(1, None)
        return Poly(c[tuple(sl)] * k)

    def __call__(self, pts: np.ndarray) -> np.ndarray:
        d = self.c.ndim
        if d == 2:
            return np.polynomial.polynomial.polyval2d(pts[:, 0], pts[:, 1], self.c)
        return np.polynomial.polynomial.polyval3d(pts[:, 0], pts[:, 1], pts[:, 2], self.c)


def _scaled_legendre(n: int, x: Poly, t: Poly) -> List[Poly]:
    """P^s_k(x,t) = t^k P_k(x/t), k = 0..n."""
    d = x.c.ndim
    out = [Poly.const(d, 1.0)]
    if n >= 1:
        out.append(x)
    for k in range(1, n):
        out.append(((2 * k + 1) / (k + 1)) * (x * out[k]) - (k / (k + 1)) * (t * t * out[k - 1]))
    return out[:n + 1]


def _compose_1d(coef_high_first: np.ndarray, z: Poly) -> Poly:
    d = z.c.ndim
    out = Poly.const(d, 0.0)
    for c in coef_high_first:
        out = out * z + float(c)
    return out


def _barycentric(d: int) -> List[Poly]:
    xs = [Poly.var(d, a) for a in range(d)]
    lam0 = Poly.const(d, 1.0)
    for x in xs:
        lam0 = lam0 - x
    return [lam0] + xs


def _dubiner_tri(p: int, la: Poly, lb: Poly, lc: Poly) -> List[Poly]:
    """Orthogonal modes of total degree <= p on the triangle, constant first."""
    out = []
    if p < 0:
        return out
    leg = _scaled_legendre(p, lb - la, la + lb)
    for i in range(p + 1):
        for j in range(p + 1 - i):
            jac = _compose_1d(jacobi(j, 2 * i + 1, 0).coeffs, 2.0 * lc - 1.0) if j > 0 else Poly.const(2, 1.0)
            out.append(leg[i] * jac)
    return out


def _dubiner_tet(p: int, lam: List[Poly]) -> List[Poly]:
    out = []
    if p < 0:
        return out
    l0, l1, l2, l3 = lam
    leg = _scaled_legendre(p, l1 - l0, l0 + l1)
    for i in range(p + 1):
        for j in range(p + 1 - i):
            # scaled Jacobi in (l2 - (l0+l1)) with scale (l0+l1+l2)
            cj = jacobi(j, 2 * i + 1, 0).coeffs[::-1] if j > 0 else np.array([1.0])
            s = l0 + l1 + l2
            xj = l2 - l0 - l1
            pj = Poly.const(3, 0.0)
            for m, c in enumerate(cj):
                term = Poly.const(3, float(c))
                for _ in range(m):
                    term = term * xj
                for _ in range(len(cj) - 1 - m):
                    term = term * s
                pj = pj + term
            for k in range(p + 1 - i - j):
                jk = _compose_1d(jacobi(k, 2 * i + 2 * j + 2, 0).coeffs, 2.0 * l3 - 1.0) if k > 0 \
                    else Poly.const(3, 1.0)
                out.append(leg[i] * pj * jk)
    return out


# ---------------------------------------------------------------------------------------------------------------
class Basis:
    """kind: 'scalar' (value + gradient rows) or 'hdiv' (vector value + gradient rows, contravariant Piola).

    entity_dofs: list of (entity_dim, local_entity_index, count, tag) in LOCAL dof order, where entity_dim is
    0 vertex, 1 edge, 2 face (3-D), 'cell', 'facet'; tag distinguishes the low-order / high-order facet blocks.
    """
    kind = 'scalar'

    def __init__(self, cell_type: str, order: int):
        self.cell_type = cell_type
        self.order = order
        self.dim = local_topology(cell_type)['dim']
        self.entity_dofs: List[tuple] = []
        self.ndof = 0
        self._cache: dict = {}

    @property
    def ncomp(self) -> int:
        return 1 if self.kind == 'scalar' else self.dim

    @property
    def nrows(self) -> int:
        """reference rows per basis function: value rows followed by gradient rows."""
        return self.ncomp * (1 + self.dim)

    def tabulate(self, pts: np.ndarray) -> np.ndarray:
        """(npts, nrows, ndof): rows = [value comps..., d(comp c)/d(xi_a) at c*dim + a ...]."""
        raise NotImplementedError

    def tabulate_cell(self, deg: int) -> np.ndarray:
        """(nq, nrows, ndof) at the cell rule of degree ``deg``."""
        key = ('c', deg)
        if key not in self._cache:
            self._cache[key] = self.tabulate(cell_rule(self.cell_type, deg)[0])
        return self._cache[key]

    def tabulate_facets(self, deg: int) -> np.ndarray:
        """(nfc, nqf, nrows, ndof) at the facet rule of degree ``deg`` mapped onto every local facet."""
        key = ('f', deg)
        if key not in self._cache:
            from .quadrature import facet_rule_in_cell
            fp, _ = facet_rule_in_cell(self.cell_type, deg)
            self._cache[key] = np.stack([self.tabulate(fp[l]) for l in range(fp.shape[0])], axis=0)
        return self._cache[key]


class _PolyScalarBasis(Basis):
    def _finish(self, polys: List[Poly]):
        self._polys = polys
        self._grads = [[p.diff(a) for a in range(self.dim)] for p in polys]
        self.ndof = len(polys)

    def tabulate(self, pts):
        pts = np.asarray(pts, dtype=np.float64)
        out = np.zeros((pts.shape[0], 1 + self.dim, self.ndof))
        for i, p in enumerate(self._polys):
            out[:, 0, i] = p(pts)
            for a in range(self.dim):
                out[:, 1 + a, i] = self._grads[i][a](pts)
        return out


class H1Simplex(_PolyScalarBasis):
    def __init__(self, cell_type: str, order: int):
        super().__init__(cell_type, order)
        if order < 1:
            raise ValueError('H1 needs order >= 1')
        d = self.dim
        lam = _barycentric(d)
        loc = local_topology(cell_type)
        polys = list(lam)
        ent = [(0, v, 1, 'v') for v in range(d + 1)]
        p = order
        for le, (a, b) in enumerate(loc['edges']):
            if p >= 2:
                leg = _scaled_legendre(p - 2, lam[b] - lam[a], lam[a] + lam[b])
                polys += [lam[a] * lam[b] * q for q in leg]
                ent.append((1, le, p - 1, 'e'))
        if d == 2:
            if p >= 3:
                inner = _dubiner_tri(p - 3, lam[0], lam[1], lam[2])
                bub = lam[0] * lam[1] * lam[2]
                polys += [bub * q for q in inner]
                ent.append(('cell', 0, len(inner), 'c'))
        else:
            if p >= 3:
                for lf, (a, b, c) in enumerate(loc['facets']):
                    inner = _dubiner_face(p - 3, lam[a], lam[b], lam[c])
                    bub = lam[a] * lam[b] * lam[c]
                    polys += [bub * q for q in inner]
                    ent.append((2, lf, len(inner), 'f'))
            if p >= 4:
                inner = _dubiner_tet(p - 4, lam)
                bub = lam[0] * lam[1] * lam[2] * lam[3]
                polys += [bub * q for q in inner]
                ent.append(('cell', 0, len(inner), 'c'))
        self.entity_dofs = ent
        self._finish(polys)


def _dubiner_face(p: int, la: Poly, lb: Poly, lc: Poly) -> List[Poly]:
    """Dubiner-type modes on a tet face written with 3-D barycentrics (scaled so they extend into the cell)."""
    out = []
    leg = _scaled_legendre(p, lb - la, la + lb)
    s = la + lb + lc
    for i in range(p + 1):
        for j in range(p + 1 - i):
            cj = jacobi(j, 2 * i + 1, 0).coeffs[::-1] if j > 0 else np.array([1.0])
            xj = lc - la - lb
            pj = Poly.const(3, 0.0)
            for m, c in enumerate(cj):
                term = Poly.const(3, float(c))
                for _ in range(m):
                    term = term * xj
                for _ in range(len(cj) - 1 - m):
                    term = term * s
                pj = pj + term
            out.append(leg[i] * pj)
    return out


class L2Simplex(_PolyScalarBasis):
    def __init__(self, cell_type: str, order: int):
        super().__init__(cell_type, order)
        lam = _barycentric(self.dim)
        polys = _dubiner_tri(order, lam[0], lam[1], lam[2]) if self.dim == 2 else _dubiner_tet(order, lam)
        self.entity_dofs = [('cell', 0, len(polys), 'c')]
        self._finish(polys)


# ---- tensor-product cells -------------------------------------------------------------------------------------
def _h1_1d(p: int, x: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """values, derivatives (npts, p+1): 1-x, x, then x(1-x) P_{k-2}(2x-1)."""
    v = np.zeros((x.shape[0], p + 1))
    g = np.zeros_like(v)
    v[:, 0], g[:, 0] = 1 - x, -1.0
    v[:, 1], g[:, 1] = x, 1.0
    z = 2 * x - 1
    for k in range(2, p + 1):
        P = eval_legendre(k - 2, z)
        dP = legendre(k - 2).deriv()(z) if k > 2 else np.zeros_like(z)
        v[:, k] = x * (1 - x) * P
        g[:, k] = (1 - 2 * x) * P + x * (1 - x) * 2 * dP
    return v, g


def _leg_1d(p: int, x: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    z = 2 * x - 1
    v = np.stack([eval_legendre(k, z) for k in range(p + 1)], axis=1)
    g = np.stack([2 * legendre(k).deriv()(z) if k > 0 else np.zeros_like(z) for k in range(p + 1)], axis=1)
    return v, g


class _TensorScalarBasis(Basis):
    """Tensor product of 1-D bases; self._idx lists the (i,j[,k]) 1-D indices of every local dof."""

    def _fn1d(self, p, x):
        raise NotImplementedError

    def tabulate(self, pts):
        pts = np.asarray(pts, dtype=np.float64)
        d = self.dim
        vg = [self._fn1d(self.order, pts[:, a]) for a in range(d)]
        out = np.zeros((pts.shape[0], 1 + d, self.ndof))
        for n, idx in enumerate(self._idx):
            val = np.ones(pts.shape[0])
            for a in range(d):
                val = val * vg[a][0][:, idx[a]]
            out[:, 0, n] = val
            for a in range(d):
                g = vg[a][1][:, idx[a]]
                for b in range(d):
                    if b != a:
                        g = g * vg[b][0][:, idx[b]]
                out[:, 1 + a, n] = g
        return out


class H1Tensor(_TensorScalarBasis):
    def _fn1d(self, p, x):
        return _h1_1d(p, x)

    def __init__(self, cell_type: str, order: int):
        super().__init__(cell_type, order)
        p, d = order, self.dim
        loc = local_topology(cell_type)
        ref = loc['ref'].astype(int)
        idx, ent = [], []
        for v in range(loc['nv']):
            idx.append(tuple(ref[v]))
            ent.append((0, v, 1, 'v'))
        hi = list(range(2, p + 1))
        for le, (a, b) in enumerate(loc['edges']):
            if p >= 2:
                axis = int(np.nonzero(ref[b] - ref[a])[0][0])
                for k in hi:
                    t = list(ref[a])
                    t[axis] = k
                    idx.append(tuple(t))
                ent.append((1, le, p - 1, 'e'))
        if d == 3 and p >= 2:
            for lf, fv in enumerate(loc['facets']):
                fixed = lf // 2
                side = lf % 2
                free = [a for a in range(3) if a != fixed]
                for k0 in hi:
                    for k1 in hi:
                        t = [0, 0, 0]
                        t[fixed] = side
                        t[free[0]] = k0
                        t[free[1]] = k1
                        idx.append(tuple(t))
                ent.append((2, lf, (p - 1) ** 2, 'f'))
        if p >= 2:
            import itertools
            inner = list(itertools.product(hi, repeat=d))
            idx += [tuple(t) for t in inner]
            ent.append(('cell', 0, len(inner), 'c'))
        self._idx = idx
        self.entity_dofs = ent
        self.ndof = len(idx)


class L2Tensor(_TensorScalarBasis):
    def _fn1d(self, p, x):
        return _leg_1d(p, x)

    def __init__(self, cell_type: str, order: int):
        super().__init__(cell_type, order)
        import itertools
        # constant first, then by total degree
        idx = sorted(itertools.product(range(order + 1), repeat=self.dim), key=lambda t: (sum(t), t))
        self._idx = idx
        self.ndof = len(idx)
        self.entity_dofs = [('cell', 0, self.ndof, 'c')]


# ---- HDiv (BDM_k, RT_k) on triangles --------------------------------------------------------------------------
class HDivTri(Basis):
    """BDM_k = [P_k]^2 (NGSolve's default HDiv) or, with ``RT=True`` (reference models/ins.py:114-117),
    RT_k = [P_k]^2 + x P~_k (P~_k homogeneous of degree k): same k+1 normal moments per edge, interior moments against
    [P_{k-1}]^2 instead of grad P_{k-1} + curl(b P_{k-2}). Basis = dual basis of those functionals."""
    kind = 'hdiv'

    def __init__(self, cell_type: str, order: int, RT: bool = False):
        if cell_type != 'tri':
            raise NotImplementedError('HDiv is implemented on triangles')
        if order < 1:
            raise ValueError('HDiv needs order >= 1')
        super().__init__(cell_type, order)
        k = order
        self.RT = bool(RT)
        lam = _barycentric(2)
        scal = _dubiner_tri(k, lam[0], lam[1], lam[2])           # expansion basis of P_k
        ns = len(scal)
        # expansion functions e_m = (scal, 0) for m < ns, (0, scal) for m < 2 ns, then (x q, y q), q in P~_k (RT only)
        X, Y = Poly.var(2, 0), Poly.var(2, 1)
        extra = []
        if RT:
            for a in range(k + 1):
                q = Poly.const(2, 1.0)
                for _ in range(a):
                    q = q * X
                for _ in range(k - a):
                    q = q * Y
                extra.append((X * q, Y * q))
        nd = 2 * ns + len(extra)
        self._extra = extra
        self._egrad = [[[c.diff(a) for a in range(2)] for c in e] for e in extra]

        def expand(pts):
            """(npts, 2, nd): both components of every expansion function."""
            sv_ = np.stack([q(pts) for q in scal], axis=1)
            out_ = np.zeros((pts.shape[0], 2, nd))
            out_[:, 0, :ns] = sv_
            out_[:, 1, ns:2 * ns] = sv_
            for j, (ex, ey) in enumerate(extra):
                out_[:, 0, 2 * ns + j] = ex(pts)
                out_[:, 1, 2 * ns + j] = ey(pts)
            return out_
        loc = local_topology('tri')
        ref = loc['ref']
        rows = []
        ent_lo, ent_hi = [], []
        s, w = gauss_01(k + 2)
        # edge normal moments; low-order (l = 0) functionals of all edges first
        edge_rows = {}
        for le, (a, b) in enumerate(loc['edges']):
            t = ref[b] - ref[a]
            n = np.array([t[1], -t[0]])
            pts = ref[a][None, :] + s[:, None] * t[None, :]
            ev = expand(pts)                                       # (nq, 2, nd)
            for l in range(k + 1):
                ql = eval_legendre(l, 2 * s - 1) * w
                edge_rows[(le, l)] = ql @ (n[0] * ev[:, 0, :] + n[1] * ev[:, 1, :])
        for le in range(3):
            rows.append(edge_rows[(le, 0)])
            ent_lo.append(('facet', le, 1, 'lo'))
        for le in range(3):
            for l in range(1, k + 1):
                rows.append(edge_rows[(le, l)])
            ent_hi.append(('facet', le, k, 'hi'))
        # interior moments: against grad(P_{k-1} \ const) and curl(bubble * P_{k-2})
        cp, cw = cell_rule('tri', 2 * k + 2)
        ev = expand(cp)
        tests = []
        if RT:
            zero = Poly.const(2, 0.0)
            for q in _dubiner_tri(k - 1, lam[0], lam[1], lam[2]):
                tests.append((q, zero))
                tests.append((zero, q))
        else:
            for q in _dubiner_tri(k - 1, lam[0], lam[1], lam[2])[1:]:
                tests.append((q.diff(0), q.diff(1)))
            bub = lam[0] * lam[1] * lam[2]
            for q in _dubiner_tri(k - 2, lam[0], lam[1], lam[2]):
                f = bub * q
                tests.append((f.diff(1), -f.diff(0)))
        for tx, ty in tests:
            rows.append((cw * tx(cp)) @ ev[:, 0, :] + (cw * ty(cp)) @ ev[:, 1, :])
        nint = len(tests)
        V = np.stack(rows, axis=0)
        assert V.shape == (nd, nd), V.shape
        self._coef = np.linalg.inv(V)                            # column j = expansion coefficients of basis fn j
        self._scal = scal
        self._sgrad = [[q.diff(a) for a in range(2)] for q in scal]
        self.ndof = nd
        self.entity_dofs = ent_lo + ent_hi + ([('cell', 0, nint, 'c')] if nint else [])

    def tabulate(self, pts):
        pts = np.asarray(pts, dtype=np.float64)
        ns = len(self._scal)
        sv = np.stack([q(pts) for q in self._scal], axis=1)                         # (npts, ns)
        sg = np.stack([np.stack([g[a](pts) for g in self._sgrad], axis=1) for a in range(2)], axis=1)  # (npts,2,ns)
        C = self._coef
        out = np.zeros((pts.shape[0], 6, self.ndof))
        for c in range(2):
            Cc = C[c * ns:(c + 1) * ns, :]
            out[:, c, :] = sv @ Cc
            for a in range(2):
                out[:, 2 + c * 2 + a, :] = sg[:, a, :] @ Cc
        for j, (e, g) in enumerate(zip(self._extra, self._egrad)):       # RT: the x P~_k part
            Cj = C[2 * ns + j, :]
            for c in range(2):
                out[:, c, :] += np.outer(e[c](pts), Cj)
                for a in range(2):
                    out[:, 2 + c * 2 + a, :] += np.outer(g[c][a](pts), Cj)
        return out


# ---- HDiv (RT_[k]) on quadrilaterals -----------------------------------------------------------------------------
class HDivQuad(Basis):
    """Raviart-Thomas space RT_[k] = Q_{k+1,k} x Q_{k,k+1} on the reference square — the space NGSolve's
    ``HDiv(order=k)`` has on quadrilaterals (k+1 normal moments per edge like BDM_k on triangles, 2k(k+1) interior
    functions; dimension 2(k+1)(k+2)). The reference's diffuse-interface meshes are quadrilateral by default
    (config_functions/expanded_config_parser.py:79), so its HDiv-DG Stokes / INS DIM models live on this element
    (pytests/full_system/dim/dim_stokes_1). Basis = dual basis of: edge normal moments against Legendre polynomials
    (lowest order first, all edges), interior moments against Q_{k-1,k} x Q_{k,k-1}."""
    kind = 'hdiv'

    def __init__(self, cell_type: str, order: int, RT: bool = False):
        if cell_type != 'quad':
            raise NotImplementedError('HDivQuad lives on quadrilaterals')
        if order < 1:
            raise ValueError('HDiv needs order >= 1')
        super().__init__(cell_type, order)
        k = order
        self.RT = bool(RT)                    # RT_[k] already is the Raviart-Thomas element; the flag changes nothing
        # expansion functions: x-component L_a(x) L_b(y), a <= k+1, b <= k; y-component a <= k, b <= k+1
        self._ix = [(a, b) for a in range(k + 2) for b in range(k + 1)]
        self._iy = [(a, b) for a in range(k + 1) for b in range(k + 2)]
        nx, ny = len(self._ix), len(self._iy)
        nd = nx + ny
        loc = local_topology('quad')
        ref = loc['ref']
        s, w = gauss_01(k + 3)
        rows = []
        edge_rows = {}
        for le, (a, b) in enumerate(loc['edges']):
            t = ref[b] - ref[a]
            n = np.array([t[1], -t[0]])
            pts = ref[a][None, :] + s[:, None] * t[None, :]
            ev = self._expand(pts)[:, :2, :]                        # (nq, 2, nd)
            for l in range(k + 1):
                ql = eval_legendre(l, 2 * s - 1) * w
                edge_rows[(le, l)] = ql @ (n[0] * ev[:, 0, :] + n[1] * ev[:, 1, :])
        ent_lo, ent_hi = [], []
        for le in range(4):
            rows.append(edge_rows[(le, 0)])
            ent_lo.append(('facet', le, 1, 'lo'))
        for le in range(4):
            for l in range(1, k + 1):
                rows.append(edge_rows[(le, l)])
            ent_hi.append(('facet', le, k, 'hi'))
        cp, cw = cell_rule('quad', 2 * k + 3)
        ev = self._expand(cp)[:, :2, :]
        lx, _ = _leg_1d(k, cp[:, 0])
        ly, _ = _leg_1d(k, cp[:, 1])
        nint = 0
        for a in range(k):                                          # x-component tests: Q_{k-1,k}
            for b in range(k + 1):
                rows.append((cw * lx[:, a] * ly[:, b]) @ ev[:, 0, :])
                nint += 1
        for a in range(k + 1):                                      # y-component tests: Q_{k,k-1}
            for b in range(k):
                rows.append((cw * lx[:, a] * ly[:, b]) @ ev[:, 1, :])
                nint += 1
        V = np.stack(rows, axis=0)
        assert V.shape == (nd, nd), V.shape
        self._coef = np.linalg.inv(V)
        self.ndof = nd
        self.entity_dofs = ent_lo + ent_hi + [('cell', 0, nint, 'c')]

    def _expand(self, pts):
        """(npts, 6, nd): rows [ux, uy, dux/dx, dux/dy, duy/dx, duy/dy] of every expansion function."""
        k = self.order
        pts = np.asarray(pts, dtype=np.float64)
        vx1, gx1 = _leg_1d(k + 1, pts[:, 0])
        vy1, gy1 = _leg_1d(k + 1, pts[:, 1])
        nx = len(self._ix)
        out = np.zeros((pts.shape[0], 6, nx + len(self._iy)))
        for j, (a, b) in enumerate(self._ix):
            out[:, 0, j] = vx1[:, a] * vy1[:, b]
            out[:, 2, j] = gx1[:, a] * vy1[:, b]
            out[:, 3, j] = vx1[:, a] * gy1[:, b]
        for j, (a, b) in enumerate(self._iy):
            out[:, 1, nx + j] = vx1[:, a] * vy1[:, b]
            out[:, 4, nx + j] = gx1[:, a] * vy1[:, b]
            out[:, 5, nx + j] = vx1[:, a] * gy1[:, b]
        return out

    def tabulate(self, pts):
        return self._expand(pts) @ self._coef


@lru_cache(maxsize=None)
def make_basis(family: str, cell_type: str, order: int, RT: bool = False) -> Basis:
    simplex = cell_type in ('tri', 'tet')
    if family == 'H1':
        return H1Simplex(cell_type, order) if simplex else H1Tensor(cell_type, order)
    if family == 'L2':
        return L2Simplex(cell_type, order) if simplex else L2Tensor(cell_type, order)
    if family == 'HDiv':
        return HDivQuad(cell_type, order, RT) if cell_type == 'quad' else HDivTri(cell_type, order, RT)
    raise ValueError('unknown finite element family {}'.format(family))
