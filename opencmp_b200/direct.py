"""Host side of the sparse direct solve behind ``a.mat.Inverse(freedofs)`` — the reference's default
``linear_solver = direct`` (reference opencmp/models/base_model.py:908-922, UMFPACK / PARDISO through NGSolve).

The free-free block of the assembled CSR matrix is put into a bandwidth-reducing order (reverse Cuthill-McKee on the
pattern graph, once per (space, free-dof mask)) and factorised on the device as a band matrix with partial pivoting
(csrc/ocmp_direct.cu: ``ocmp_band_fill / _factor / _solve``). ``BandLU.solve`` adds two steps of iterative refinement on
the true residual, which is what brings a pivoted LU to round-off level on the regularised saddle-point systems
(``-1e-10 p q``, reference opencmp/models/ins.py:243).

Only the ordering runs on the host (graph traversal, one-off per pattern). There is no CPU solve here."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np


class DirectSolveTooLarge(RuntimeError):
    pass


def band_budget_bytes() -> int:
    return int(float(os.environ.get('OCMP_DIRECT_MAX_GB', '32')) * (1 << 30))


def rcm_order(rowptr: np.ndarray, colidx: np.ndarray, free: np.ndarray):
    """perm (ndof,) int32: position of every free dof in the reverse Cuthill-McKee order of the free-free pattern
    graph, -1 for constrained dofs; plus (n, kl, ku) of the permuted band."""
    import scipy.sparse as sp
    from scipy.sparse.csgraph import reverse_cuthill_mckee
    ndof = len(rowptr) - 1
    free = np.asarray(free, dtype=bool)
    idx = np.nonzero(free)[0]
    n = len(idx)
    perm = -np.ones(ndof, dtype=np.int32)
    if n == 0:
        return perm, 0, 0, 0
    rows = np.repeat(np.arange(ndof, dtype=np.int64), np.diff(rowptr))
    keep = free[rows] & free[colidx]
    comp = -np.ones(ndof, dtype=np.int64)
    comp[idx] = np.arange(n)
    r, c = comp[rows[keep]], comp[np.asarray(colidx)[keep]]
    g = sp.coo_matrix((np.ones(len(r), dtype=np.int8), (r, c)), shape=(n, n)).tocsr()
    g = (g + g.T).tocsr()
    order = reverse_cuthill_mckee(g, symmetric_mode=True)
    pos = np.empty(n, dtype=np.int64)
    pos[order] = np.arange(n)
    perm[idx] = pos.astype(np.int32)
    d = pos[r] - pos[c]
    kl = int(max(d.max(), 0)) if len(d) else 0
    ku = int(max(-d.min(), 0)) if len(d) else 0
    return perm, n, kl, ku


class BandLU:
    """Factorisation of the free-free block of ``mat`` at the values it holds NOW (NGSolve's ``Inverse`` also
    factorises at construction; ``Update()`` re-factorises after a re-assembly)."""

    def __init__(self, be, mat, freedofs):
        self.be, self.mat = be, mat
        t = be.torch
        fes = mat.space
        free = np.ones(mat.height, bool) if freedofs is None else np.asarray(
            freedofs.a if hasattr(freedofs, 'a') else freedofs, dtype=bool)
        self.fm = be._mask(fes, free)
        sd = be.space_data(fes)
        cache = sd.setdefault('band_orders', {})
        key = np.packbits(free).tobytes()
        if key not in cache:
            pat = fes.pattern()
            perm, n, kl, ku = rcm_order(np.asarray(pat.rowptr), np.asarray(pat.colidx), free)
            cache[key] = (be._up(perm), n, kl, ku)
        self.perm, self.n, self.kl, self.ku = cache[key]
        self.length = int(be.lib.ocmp_band_len(self.n, self.kl, self.ku))
        need = 8 * self.length
        work = float(self.n) * self.kl * (self.kl + self.ku)
        if need > band_budget_bytes() or self.kl > 24000 or work > float(os.environ.get('OCMP_DIRECT_MAX_WORK', '4e14')):
            raise DirectSolveTooLarge(
                'band factorisation of {} free dofs with bandwidths ({}, {}) needs {:.1f} GB / {:.1e} multiply-adds '
                '(limits OCMP_DIRECT_MAX_GB, OCMP_DIRECT_MAX_WORK)'.format(self.n, self.kl, self.ku, need / 2 ** 30, work))
        self.ab = t.empty(max(1, self.length), dtype=t.float64, device=be.device)
        self.ipiv = t.zeros(max(1, self.n), dtype=t.int32, device=be.device)
        self.info = t.zeros(2, dtype=t.int32, device=be.device)
        self.rhs = t.zeros(max(1, self.n), dtype=t.float64, device=be.device)
        self.ubw = self.kl + self.ku
        self.Update()

    def Update(self):
        be, mat = self.be, self.mat
        pd = be.pattern_data(mat.space)
        st = be._stream()
        be._ck(be.lib.ocmp_band_fill(mat.height, pd['rowptr'].data_ptr(), pd['colidx'].data_ptr(),
                                     mat.values.data_ptr(), self.perm.data_ptr(), self.n, self.kl, self.ku,
                                     self.ab.data_ptr(), st))
        be._ck(be.lib.ocmp_band_factor(self.n, self.kl, self.ku, self.ab.data_ptr(), self.ipiv.data_ptr(),
                                       self.info.data_ptr(), st))
        be.launches += 3
        info = self.info.cpu().numpy()
        if int(info[0]) != 0:
            raise RuntimeError('opencmp_b200 direct solve: the matrix is singular on its free dofs (zero pivot in '
                               'column {} of {})'.format(int(info[0]), self.n))
        self.ubw = int(info[1]) if self.n > 1 else 0

    def handle(self):
        """ocmp_band_lu record for ocmp_system.direct (Preconditioner(a, 'direct') inside ocmp_krylov)."""
        from .backend import BandHandle
        h = BandHandle()
        h.n, h.kl, h.ku, h.ubw = self.n, self.kl, self.ku, self.ubw
        h.ab, h.ipiv, h.perm, h.rhs = (self.ab.data_ptr(), self.ipiv.data_ptr(), self.perm.data_ptr(),
                                       self.rhs.data_ptr())
        return h

    def _apply(self, r, out, accumulate: bool):
        be = self.be
        st = be._stream()
        n_all = self.mat.height
        be._ck(be.lib.ocmp_band_gather(n_all, self.perm.data_ptr(), r.data_ptr(), self.rhs.data_ptr(), st))
        be._ck(be.lib.ocmp_band_solve(self.n, self.kl, self.ku, self.ubw, self.ab.data_ptr(), self.ipiv.data_ptr(),
                                      self.rhs.data_ptr(), st))
        be._ck(be.lib.ocmp_band_scatter(n_all, self.perm.data_ptr(), self.rhs.data_ptr(), out.data_ptr(),
                                        1 if accumulate else 0, st))
        be.launches += 3

    def solve(self, r, out, refine: int = 2):
        """out = A_ff^{-1} r on the free dofs (0 elsewhere), then ``refine`` steps out += A_ff^{-1}(r - A out)."""
        be = self.be
        self._apply(r, out, False)
        if self.n == 0:
            return
        res = be.zeros(self.mat.height)
        for _ in range(refine):
            be.spmv(self.mat, out, res)                       # res = A out
            be._ck(be.lib.ocmp_axpby(res.numel(), 1.0, r.data_ptr(), -1.0, res.data_ptr(), be._stream()))
            be.launches += 1
            self._apply(res, out, True)                       # the gather only reads the free entries of res
