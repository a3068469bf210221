"""Index-level emulation of csrc/ocmp_direct.cu (band LU with partial pivoting + substitution with a circular
shared-memory window) in NumPy, so that the storage arithmetic of the kernels is checked without a GPU: same band
layout (LAPACK dgbtrf), same look-ahead order of operations, same window slots and refill schedule."""
import numpy as np
import pytest
import scipy.sparse as sp

from opencmp_b200.direct import rcm_order

AHEAD = 32


def band_fill(A, perm, n, kl, ku):
    ld = 2 * kl + ku + 1
    kv = kl + ku
    ab = np.zeros(n * ld)
    A = A.tocsr()
    for row in range(A.shape[0]):
        pi = perm[row]
        if pi < 0:
            continue
        for k in range(A.indptr[row], A.indptr[row + 1]):
            pc = perm[A.indices[k]]
            if pc >= 0:
                ab[pc * ld + kv + pi - pc] = A.data[k]
    return ab


def band_lu(n, kl, ku, ab):
    """k_band_lu, column by column (finalise j+1 inside step j)."""
    kv, ld = kl + ku, 2 * kl + ku + 1
    ipiv = np.zeros(n, dtype=np.int64)
    info = [0, 0]

    def finalize(j):
        col = ab[j * ld + kv:]
        km = min(kl, n - 1 - j)
        a = np.abs(col[:km + 1])
        bi = int(np.argmax(a))
        if not a[bi] > 0:
            bi = 0
            if info[0] == 0:
                info[0] = j + 1
        ipiv[j] = j + bi
        col[0], col[bi] = col[bi], col[0]
        if col[0] != 0:
            col[1:km + 1] /= col[0]

    finalize(0)
    ju = umax = 0
    for j in range(n - 1):
        km = min(kl, n - 1 - j)
        colj = ab[j * ld + kv:]
        jp = ipiv[j] - j
        regular = colj[0] != 0
        if regular:
            ju = max(ju, min(j + ku + jp, n - 1))
        umax = max(umax, ju - j)
        sl = colj[1:1 + km].copy()
        cols = [j + 1] if (regular and j + 1 <= ju) else []
        for c in cols:
            cc = ab[c * ld + kv + j - c:]
            t, s = cc[0], cc[jp]
            for i in range(1, km + 1):
                cc[i] = (t if i == jp else cc[i]) - sl[i - 1] * s
            cc[0] = s
        finalize(j + 1)
        if regular:
            for c in range(j + 2, ju + 1):
                cc = ab[c * ld + kv + j - c:]
                t, s = cc[0], cc[jp]
                v = cc[1:km + 1].copy()
                if jp >= 1:
                    v[jp - 1] = t
                cc[1:km + 1] = v - sl * s
                cc[0] = s
    info[1] = max(umax, min(ku, n - 1))
    return ipiv, info


def band_solve(n, kl, ku, ubw, ab, ipiv, b, nt):
    """k_band_solve with nt 'threads': the circular window and its refill schedule."""
    kv, ld = kl + ku, 2 * kl + ku + 1
    if kl > 0:
        W = kl + 1 + 2 * AHEAD
        w = np.full(W, np.nan)
        for i in range(min(n, kl + 1 + AHEAD)):
            w[i % W] = b[i]
        for j in range(n):
            km = min(kl, n - 1 - j)
            col = ab[j * ld + kv:]
            p = ipiv[j]
            a, bp = w[j % W], w[p % W]
            for i in range(1, km + 1):
                s = (j + i) % W
                w[s] = (a if j + i == p else w[s]) - col[i] * bp
            b[j] = bp
            if j % AHEAD == 0:
                for tid in range(nt - AHEAD, nt):
                    idx = j + kl + 1 + AHEAD + (tid - (nt - AHEAD))
                    if idx < n:
                        w[idx % W] = b[idx]
    W = ubw + 1 + 2 * AHEAD
    w = np.full(W, np.nan)
    for i in range(min(n, ubw + 1 + AHEAD)):
        w[(n - 1 - i) % W] = b[n - 1 - i]
    for j in range(n - 1, -1, -1):
        km = min(ubw, j)
        base = j * ld + kv
        xj = w[j % W] / ab[base]
        for i in range(1, km + 1):
            w[(j - i) % W] -= ab[base - i] * xj
        b[j] = xj
        step = n - 1 - j
        if step % AHEAD == 0:
            for tid in range(nt - AHEAD, nt):
                idx = j - ubw - 1 - AHEAD - (tid - (nt - AHEAD))
                if idx >= 0:
                    w[idx % W] = b[idx]
    return b


def _saddle_point(nx, rng):
    """2-D Laplacian velocity-like block + a zero pressure-like block coupled by a random sparse B: forces pivoting."""
    n1 = nx * nx
    T = sp.diags([-1, 2.2, -1], [-1, 0, 1], shape=(nx, nx))
    K = sp.kron(sp.eye(nx), T) + sp.kron(T, sp.eye(nx))
    n2 = n1 // 3
    B = sp.random(n2, n1, density=3.0 / n1, random_state=np.random.RandomState(3), data_rvs=rng.standard_normal)
    B = B + sp.coo_matrix((np.ones(n2), (np.arange(n2), 3 * np.arange(n2))), shape=(n2, n1))
    A = sp.bmat([[K, B.T], [B, -1e-10 * sp.eye(n2)]]).tocsr()
    return A


@pytest.mark.parametrize('nx,constrain', [(6, False), (13, True), (20, True)])
def test_band_lu_emulation_matches_dense_solve(nx, constrain):
    rng = np.random.default_rng(nx)
    A = _saddle_point(nx, rng)
    ndof = A.shape[0]
    free = np.ones(ndof, bool)
    if constrain:
        free[rng.choice(ndof, ndof // 7, replace=False)] = False
    perm, n, kl, ku = rcm_order(A.indptr, A.indices, free)
    assert n == free.sum() and sorted(perm[free]) == list(range(n)) and (perm[~free] == -1).all()
    ab = band_fill(A, perm, n, kl, ku)
    ipiv, info = band_lu(n, kl, ku, ab)
    assert info[0] == 0
    assert (ipiv != np.arange(n)).any(), 'the case should exercise row interchanges'
    r = rng.standard_normal(ndof)
    b = np.zeros(n)
    b[perm[free]] = r[free]
    x = band_solve(n, kl, ku, info[1], ab, ipiv, b, nt=max(64, min(1024, (max(kl, info[1]) + 31) // 32 * 32)))
    out = np.zeros(ndof)
    out[free] = x[perm[free]]
    idx = np.nonzero(free)[0]
    ref = np.linalg.solve(A[idx][:, idx].toarray(), r[idx])
    assert np.abs(out[idx] - ref).max() <= 1e-9 * np.abs(ref).max()


def test_band_lu_reports_a_singular_matrix():
    A = sp.csr_matrix(np.array([[1.0, 2.0, 0.0], [2.0, 4.0, 0.0], [0.0, 0.0, 1.0]]))
    perm, n, kl, ku = rcm_order(A.indptr, A.indices, np.ones(3, bool))
    ab = band_fill(A, perm, n, kl, ku)
    _, info = band_lu(n, kl, ku, ab)
    assert info[0] != 0


def test_rcm_reduces_the_bandwidth_of_a_shuffled_grid():
    nx = 12
    T = sp.diags([-1, 2, -1], [-1, 0, 1], shape=(nx, nx))
    K = (sp.kron(sp.eye(nx), T) + sp.kron(T, sp.eye(nx))).tocsr()
    sh = np.random.default_rng(0).permutation(nx * nx)
    K = K[sh][:, sh].tocsr()
    perm, n, kl, ku = rcm_order(K.indptr, K.indices, np.ones(nx * nx, bool))
    assert max(kl, ku) <= 2 * nx
