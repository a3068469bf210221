"""Oracle backend — TEST INFRASTRUCTURE ONLY.

Implements the backend protocol of opencmp_b200/ngs.py with NumPy/SciPy on the host so that tests can (a) run the
NGSolve-style front end (and, when /root/reference is mounted, unmodified OpenCMP models) against the reference's
golden error norms, and (b) produce the expected values the CUDA path is compared with. Never imported by the
product package.

Krylov restatements follow the textbook algorithms NGSolve's pure-Python ``ngsolve.solvers`` implements
(call sites reference opencmp/models/base_model.py:924-944): preconditioned CG with the
sqrt(<r, P r>) stopping rule, right-projected GMRES on the free DOFs with Givens rotations, MINRES, damped Richardson.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from . import fem


class OracleBackend:
    name = 'oracle'

    # ---- storage -------------------------------------------------------------------------------------------------
    def zeros(self, n):
        return np.zeros(int(n))

    def from_numpy(self, a):
        return np.array(a, dtype=np.float64)

    def to_numpy(self, a):
        return np.asarray(a)

    def numpy_view(self, a):
        return a

    def copy_into(self, dst, src):
        dst[:] = src

    def dot(self, a, b):
        return float(np.dot(a, b))

    def mdot(self, V, w):
        return np.asarray(V) @ np.asarray(w)

    def maxpy(self, V, coef, w):
        w += np.asarray(coef, dtype=np.float64) @ np.asarray(V)

    def masked_assign(self, dst, src, inv, mask):
        m = mask > 0
        dst[m] = src[m] * inv[m]

    # ---- assembly ------------------------------------------------------------------------------------------------
    def assemble_matrix(self, program, mat):
        mat.values[:] = fem.assemble_matrix(program)

    def assemble_vector(self, program, out):
        out[:] = fem.assemble_vector(program)

    def integrate(self, program):
        return fem.integrate(program)

    # ---- linear algebra ------------------------------------------------------------------------------------------
    def _csr(self, mat):
        return fem.csr_matrix(mat.space, mat.values)

    def spmv(self, mat, x, out):
        out[:] = self._csr(mat) @ x

    def solve_free(self, mat, r, out, freedofs):
        free = np.ones(mat.height, bool) if freedofs is None else np.asarray(freedofs, bool)
        out[:] = fem.solve_direct(self._csr(mat), r, np.zeros_like(r), free)

    def precond_setup(self, mat, kind, free, form=None, state=None, mask=None):
        A = self._csr(mat)
        free = np.asarray(free, bool)
        idx = np.nonzero(free)[0]
        if kind in ('local', 'jacobi'):
            d = A.diagonal()
            inv = np.where(free & (d != 0), 1.0 / np.where(d != 0, d, 1.0), 0.0)
            return lambda v: inv * v
        if kind == 'direct':
            lu = spla.splu(A[idx][:, idx].tocsc())

            def app(v):
                o = np.zeros_like(v)
                o[idx] = lu.solve(v[idx])
                return o
            return app
        raise NotImplementedError('oracle preconditioner {}'.format(kind))

    def krylov(self, kind, mat, b, x, pre, freedofs, tol, maxit, initialize, printrates, damp=1.0, restart=None):
        A = _FormAction(mat) if getattr(mat, 'matrix_free', False) else self._csr(mat)
        n = A.shape[0]
        if pre is not None and pre.state is None:
            pre.Update()
        P = pre.state if pre is not None else (lambda v: v)
        free = None if freedofs is None else np.asarray(freedofs, bool)
        proj = (lambda v: v) if free is None else (lambda v: np.where(free, v, 0.0))
        if kind == 'cg':
            x[:] = cg(A, b, x, P, tol, maxit, initialize)
        elif kind == 'gmres':
            x[:] = gmres(A, b, x, P, proj, tol, maxit)
        elif kind == 'minres':
            if initialize:
                x[:] = 0.0
            x[:] = minres(A, b, x, P, proj, tol, maxit)
        elif kind == 'richardson':
            x[:] = 0.0
            r = proj(b - A @ x)
            r0 = np.linalg.norm(r)
            for _ in range(maxit):
                x += damp * proj(P(r))
                r = proj(b - A @ x)
                if np.linalg.norm(r) < tol * r0:
                    break
        else:
            raise ValueError(kind)


class _FormAction:
    """``A @ v`` of a ``BilinearForm(nonassemble=True).mat``: the form's action, assembled as a linear form."""

    def __init__(self, op):
        self.op = op
        self.shape = (op.height, op.width)

    def __matmul__(self, v):
        out = np.zeros(self.shape[0])
        self.op.bf.apply_arrays(np.ascontiguousarray(v, dtype=np.float64), out)
        return out


def cg(A, b, x, P, tol, maxit, initialize):
    u = np.zeros_like(b) if initialize else x.copy()
    d = b - A @ u
    w = P(d)
    wdn = float(w @ d)
    err0 = np.sqrt(abs(wdn))
    if wdn == 0:
        return u
    s = w.copy()
    for _ in range(maxit):
        w = A @ s
        wd = wdn
        alpha = wd / float(s @ w)
        u += alpha * s
        d -= alpha * w
        w = P(d)
        wdn = float(w @ d)
        s = w + (wdn / wd) * s
        if np.sqrt(abs(wd)) < tol * err0:
            break
    return u


last_history = []


def minres(A, b, x, P, proj, tol, maxit):
    """Preconditioned MINRES (Paige & Saunders 1975; the form of Elman, Silvester & Wathen, Alg. 4.1): Lanczos in the
    P-inner product, two Givens rotations per step; stops on the recurrence value of the P-norm of the residual relative
    to the initial one — what ngsolve.solvers.MinRes monitors (reference base_model.py:929-932). Records that ratio per
    iteration in ``last_history``."""
    u = x.copy()
    v0 = np.zeros_like(b)
    v1 = proj(b - A @ u)
    z1 = P(v1)
    w0, w1 = np.zeros_like(b), np.zeros_like(b)
    gamma0, gamma1 = 1.0, np.sqrt(abs(float(z1 @ v1)))
    eta, s0, s1, c0, c1 = gamma1, 0.0, 0.0, 1.0, 1.0
    err0 = gamma1
    del last_history[:]
    if gamma1 == 0.0:
        return u
    for _ in range(maxit):
        z1 = z1 / gamma1
        Az = proj(A @ z1)
        delta = float(Az @ z1)
        v2 = Az - (delta / gamma1) * v1 - (gamma1 / gamma0) * v0
        z2 = P(v2)
        gamma2 = np.sqrt(abs(float(z2 @ v2)))
        a0 = c1 * delta - c0 * s1 * gamma1
        a1 = np.sqrt(a0 * a0 + gamma2 * gamma2)
        a2 = s1 * delta + c0 * c1 * gamma1
        a3 = s0 * gamma1
        c0, s0 = c1, s1
        c1, s1 = (1.0, 0.0) if a1 == 0.0 else (a0 / a1, gamma2 / a1)
        w2 = z1 - a3 * w0 - a2 * w1
        if a1 != 0.0:
            w2 = w2 / a1
        u += c1 * eta * w2
        eta = -s1 * eta
        last_history.append(abs(eta) / err0)
        if abs(eta) < tol * err0 or gamma2 == 0.0:
            break
        v0, v1, z1, w0, w1 = v1, v2, z2, w1, w2
        gamma0, gamma1 = gamma1, gamma2
    return u


def gmres(A, b, x, P, proj, tol, maxit):
    """Left-preconditioned GMRES without restart on the projected system."""
    u = x.copy()
    r = P(proj(b - A @ u))
    beta = np.linalg.norm(r)
    if beta == 0:
        return u
    m = maxit
    Q = [r / beta]
    H = np.zeros((m + 1, m))
    cs, sn = np.zeros(m), np.zeros(m)
    g = np.zeros(m + 1)
    g[0] = beta
    k_used = 0
    for k in range(m):
        w = P(proj(A @ Q[k]))
        for i in range(k + 1):
            H[i, k] = Q[i] @ w
            w = w - H[i, k] * Q[i]
        H[k + 1, k] = np.linalg.norm(w)
        for i in range(k):
            t = cs[i] * H[i, k] + sn[i] * H[i + 1, k]
            H[i + 1, k] = -sn[i] * H[i, k] + cs[i] * H[i + 1, k]
            H[i, k] = t
        den = np.hypot(H[k, k], H[k + 1, k])
        cs[k], sn[k] = (1.0, 0.0) if den == 0 else (H[k, k] / den, H[k + 1, k] / den)
        H[k, k] = cs[k] * H[k, k] + sn[k] * H[k + 1, k]
        H[k + 1, k] = 0.0
        g[k + 1] = -sn[k] * g[k]
        g[k] = cs[k] * g[k]
        k_used = k + 1
        if abs(g[k + 1]) < tol * beta or den == 0:
            break
        Q.append(w / np.linalg.norm(w))
    yk = np.linalg.solve(np.triu(H[:k_used, :k_used]), g[:k_used])
    for i in range(k_used):
        u += yk[i] * Q[i]
    return u


# ---- primitives of the element-partitioned multigrid driver, NumPy versions (gloo tests) --------------------------
def _vertex_patches(fes, vmask):
    m = fes.mesh
    v2c = [[] for _ in range(m.nv)]
    for c in range(m.ne):
        for v in m.cells[c]:
            v2c[v].append(c)
    keep = np.arange(m.nv) if vmask is None else np.nonzero(np.asarray(vmask, bool))[0]
    return [np.unique(np.concatenate([fes.cell_dofs[c] for c in v2c[v]])) for v in keep]


class _OraclePrimitives:
    def csr_handle(self, m):
        return m.tocsr()

    def csr_mult(self, h, x, y):
        y[:] = h @ x

    def patch_state(self, fes, vmask):
        import os
        kind = os.environ.get('OCMP_PATCH', 'vertex')
        if kind != 'vertex':
            # open-star / Vanka patch layouts are host logic of the product (opencmp_b200/patches.py); the oracle only
            # restates the arithmetic done on them
            from opencmp_b200.patches import vertex_patch_dofs
            return dict(dofs=[d[d >= 0] for d in vertex_patch_dofs(fes, kind, vmask)], inv=None)
        dofs = _vertex_patches(fes, vmask)
        if max(len(d) for d in dofs) > 160:          # same size rule as CudaBackend._patches
            from opencmp_b200.patches import vertex_patch_dofs
            dofs = [d[d >= 0] for d in vertex_patch_dofs(fes, 'star', vmask)]
        return dict(dofs=dofs, inv=None)

    def patch_setup(self, mat, pt, fm):
        A = self._csr(mat)
        free = np.asarray(fm) > 0
        inv = []
        for d in pt['dofs']:
            M = A[d][:, d].toarray()
            f = free[d]
            M[~f, :] = 0.0
            M[:, ~f] = 0.0
            M[~f, ~f] = 1.0
            inv.append(np.linalg.inv(M))
        pt['inv'] = inv

    def patch_apply(self, pt, r, z):
        z[:] = 0.0
        for d, Ai in zip(pt['dofs'], pt['inv']):
            z[d] += Ai @ r[d]

    def patch_count(self, pt, n):
        c = np.zeros(n)
        for d in pt['dofs']:
            c[d] += 1.0
        return c

    def dense_inverse(self, mat, fm):
        A = self._csr(mat).toarray()
        f = np.asarray(fm)
        A = A * f[:, None] * f[None, :] + np.diag(1.0 - f)
        return np.linalg.inv(A)


for _n in ('csr_handle', 'csr_mult', 'patch_state', 'patch_setup', 'patch_apply', 'patch_count', 'dense_inverse'):
    setattr(OracleBackend, _n, getattr(_OraclePrimitives, _n))
