"""CPU oracle — TEST INFRASTRUCTURE ONLY (imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline).

NumPy/SciPy restatement of the arithmetic OpenCMP delegates to NGSolve on its hot path: evaluation
of the weak-form integrands at quadrature points, element / facet matrices and vectors, scatter into the global
CSR matrix, ``Integrate`` reductions, local L2 projection for ``GridFunction.Set`` and the sparse direct solve of
``Model.linear_solve`` (reference opencmp/models/base_model.py:886-947: ``x += A_ff^-1 (b - A x)``).

The oracle consumes the same lowered form programs (opencmp_b200/ir.py) and the same exported mesh / space arrays
as the CUDA path, but follows a different algebraic route: it builds *physical* basis tables per cell
(J^-T grad, Piola) and contracts them entry by entry with einsum, whereas the kernels fold the geometry into a
reference-row coefficient matrix. Agreement of the two is therefore a real check of the kernels.

PARITY STATUS: NGSolve itself is absent from this container and un-pinned in the reference (setup.cfg:17-22), so the
oracle is pinned through the reference's own manufactured-solution fixtures and golden error norms
(pytests/full_system/*) in tests/test_oracle_golden.py — not against NGSolve matrix entries,
which no reference test holds.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla
from scipy.special import erf as _erf

from opencmp_b200.ir import OPS, FormProgram, Integral
from opencmp_b200.mesh import Mesh
from opencmp_b200.quadrature import cell_rule, facet_rule_in_cell, facet_ref_geometry

_OPN = {i: n for i, n in enumerate(OPS)}


# ---- geometry ----------------------------------------------------------------------------------------------------
class Geo:
    """Per-item geometry at the quadrature points of one integral."""

    def __init__(self, mesh: Mesh, integral: Integral):
        dim = mesh.dim
        J_all = mesh.jacobians()
        det_all = np.linalg.det(J_all)
        Jinv_all = np.linalg.inv(J_all)
        org = mesh.origins()
        self.kind = integral.kind
        self.deg = integral.deg
        if integral.kind == 'cell':
            items = np.arange(mesh.ne) if integral.items is None else integral.items
            pts, w = cell_rule(mesh.cell_type, integral.deg)
            self.cells = [items]
            self.ref_pts = [np.broadcast_to(pts, (len(items),) + pts.shape)]
            self.lf = [None]
            self.w = w
            self.scale = np.abs(det_all[items])
            self.h = np.abs(det_all[items]) ** (1.0 / dim)
            self.normal = None
            self.x = org[items][:, None, :] + np.einsum('eia,qa->eqi', J_all[items], pts)
        else:
            f = integral.items
            fpts, w = facet_rule_in_cell(mesh.cell_type, integral.deg)       # (nfc, nq, dim)
            tang, nref = facet_ref_geometry(mesh.cell_type)
            c0 = mesh.facet_cells[f, 0]
            l0 = mesh.facet_local[f, 0]
            self.cells = [c0]
            self.lf = [l0]
            self.ref_pts = [fpts[l0]]
            if integral.kind == 'ifacet':
                c1 = mesh.facet_cells[f, 1]
                l1 = mesh.facet_local[f, 1]
                self.cells.append(c1)
                self.lf.append(l1)
                self.ref_pts.append(fpts[l1])
            self.w = w
            T = np.einsum('eia,eka->eki', J_all[c0], tang[l0])               # physical tangents (nf, dim-1, dim)
            if dim == 2:
                meas = np.linalg.norm(T[:, 0, :], axis=1)
            else:
                meas = np.linalg.norm(np.cross(T[:, 0, :], T[:, 1, :]), axis=1)
            self.scale = meas
            n = np.einsum('eai,ea->ei', Jinv_all[c0], nref[l0])              # J^-T n_ref
            self.normal = n / np.linalg.norm(n, axis=1)[:, None]
            self.h = np.abs(det_all[c0]) / meas
            self.x = org[c0][:, None, :] + np.einsum('eia,eqa->eqi', J_all[c0], self.ref_pts[0])
        self.J = [J_all[c] for c in self.cells]
        self.Jinv = [Jinv_all[c] for c in self.cells]
        self.det = [det_all[c] for c in self.cells]
        self.nitems = len(self.cells[0])
        self.nq = len(self.w)


def phys_table(block, geo: Geo, side: int) -> np.ndarray:
    """(nitems, nq, nrows, nloc) physical operator rows of one block on one side."""
    basis = block.basis
    dim = basis.dim
    if geo.kind == 'cell':
        ref = np.broadcast_to(basis.tabulate_cell(geo.deg)[None], (geo.nitems, geo.nq, basis.nrows, basis.ndof))
    else:
        ref = basis.tabulate_facets(geo.deg)[geo.lf[side]]
    J, Jinv, det = geo.J[side], geo.Jinv[side], geo.det[side]
    out = np.zeros_like(ref)
    if basis.kind == 'scalar':
        out[:, :, 0, :] = ref[:, :, 0, :]
        out[:, :, 1:, :] = np.einsum('eba,eqbi->eqai', Jinv, ref[:, :, 1:, :])
    else:
        val = ref[:, :, :dim, :]
        gr = ref[:, :, dim:, :].reshape(ref.shape[0], ref.shape[1], dim, dim, -1)   # [b, m]
        out[:, :, :dim, :] = np.einsum('ecb,eqbi->eqci', J, val) / det[:, None, None, None]
        g = np.einsum('ecb,eqbmi,ema->eqcai', J, gr, Jinv) / det[:, None, None, None, None]
        out[:, :, dim:, :] = g.reshape(ref.shape[0], ref.shape[1], dim * dim, -1)
    return out


# ---- bytecode interpreter (NumPy, vectorised over items x quadrature points) ------------------------------------
def run_bytecode(prog, geo: Geo, params: np.ndarray, fields: List[np.ndarray]) -> np.ndarray:
    shape = (geo.nitems, geo.nq)
    reg: List[Optional[np.ndarray]] = [None] * prog.nreg
    out = np.zeros((prog.nout,) + shape)
    ones = np.ones(shape)
    for op_dst, a, b, c in prog.code:
        op = _OPN[int(op_dst) & 0xff]
        d = int(op_dst) >> 8
        if op == 'CONST':
            v = prog.consts_arr[a] * ones
        elif op == 'PARAM':
            v = params[a] * ones
        elif op == 'COORD':
            v = geo.x[:, :, a] if a < geo.x.shape[2] else 0.0 * ones
        elif op == 'NORMAL':
            v = geo.normal[:, a][:, None] * ones
        elif op == 'MESHSIZE':
            v = geo.h[:, None] * ones
        elif op == 'MEASURE':
            v = geo.scale[:, None] * ones
        elif op == 'FIELD':
            v = fields[a]
        elif op == 'ADD':
            v = reg[a] + reg[b]
        elif op == 'SUB':
            v = reg[a] - reg[b]
        elif op == 'MUL':
            v = reg[a] * reg[b]
        elif op == 'DIV':
            v = reg[a] / reg[b]
        elif op == 'NEG':
            v = -reg[a]
        elif op == 'ABS':
            v = np.abs(reg[a])
        elif op == 'SQRT':
            v = np.sqrt(reg[a])
        elif op == 'SIN':
            v = np.sin(reg[a])
        elif op == 'COS':
            v = np.cos(reg[a])
        elif op == 'TAN':
            v = np.tan(reg[a])
        elif op == 'EXP':
            v = np.exp(reg[a])
        elif op == 'LOG':
            v = np.log(reg[a])
        elif op == 'POW':
            v = np.power(reg[a], reg[b])
        elif op == 'IFPOS':
            v = np.where(reg[a] > 0, reg[b], reg[c])
        elif op == 'MIN':
            v = np.minimum(reg[a], reg[b])
        elif op == 'MAX':
            v = np.maximum(reg[a], reg[b])
        elif op == 'TANH':
            v = np.tanh(reg[a])
        elif op == 'ERF':
            v = _erf(reg[a])
        elif op == 'FLOOR':
            v = np.floor(reg[a])
        elif op == 'CEIL':
            v = np.ceil(reg[a])
        elif op == 'ROUND':
            v = np.round(reg[a])
        elif op == 'TRUNC':
            v = np.trunc(reg[a])
        elif op == 'SGN':
            v = np.sign(reg[a])
        elif op == 'ATAN':
            v = np.arctan(reg[a])
        elif op == 'OUT':
            out[a] = reg[b]
            continue
        elif op == 'MOV':
            v = reg[a]
        else:
            raise ValueError(op)
        reg[d] = v
    return out


def _field_values(prog, geo: Geo, tables: Dict[tuple, np.ndarray]) -> List[np.ndarray]:
    vals = []
    for (gf, blk, row, side) in prog.fields:
        fes = gf.space
        block = fes.blocks[blk]
        key = (id(block.basis), side)
        if key not in tables:
            tables[key] = phys_table(block, geo, side)
        tab = tables[key]
        dofs = block.cell_dofs[geo.cells[side]] + fes.block_offsets[blk]
        coef = np.asarray(gf.vec_numpy())[dofs]
        vals.append(np.einsum('eqi,ei->eq', tab[:, :, row, :], coef))
    return vals


def _row_lookup(fes):
    """row -> (block index, row inside block)."""
    out = []
    for b, blk in enumerate(fes.blocks):
        out += [(b, r) for r in range(blk.basis.nrows)]
    return out


def element_tensors(program: FormProgram, integral: Integral):
    """Dense local matrices (arity 2): (nitems, nside*nloc, nside*nloc); vectors (arity 1); values (arity 0)."""
    fes = program.fes
    mesh = fes.mesh
    geo = Geo(mesh, integral)
    tables: Dict[tuple, np.ndarray] = {}
    params = program.param_values(integral)
    fvals = _field_values(integral.prog, geo, tables)
    D = run_bytecode(integral.prog, geo, params, fvals)                       # (nout, nitems, nq)
    wq = geo.w[None, :] * geo.scale[:, None]
    nloc, nrows = fes.nloc, fes.nrows
    nside = 2 if integral.kind == 'ifacet' else 1
    rows = _row_lookup(fes)
    lo = fes.loc_offsets

    def tab(r):
        side, r = divmod(int(r), nrows)
        b, rr = rows[r]
        key = (id(fes.blocks[b].basis), side)
        if key not in tables:
            tables[key] = phys_table(fes.blocks[b], geo, side)
        return tables[key][:, :, rr, :], side * nloc + lo[b], fes.blocks[b].nloc

    if program.arity == 2:
        A = np.zeros((geo.nitems, nside * nloc, nside * nloc))
        for tr, ur, k in integral.entries:
            Bt, ot, nt = tab(tr)
            Bu, ou, nu = tab(ur)
            A[:, ot:ot + nt, ou:ou + nu] += np.einsum('eq,eqi,eqj->eij', wq * D[k], Bt, Bu)
        return geo, A
    if program.arity == 1:
        b = np.zeros((geo.nitems, nside * nloc))
        for tr, _, k in integral.entries:
            Bt, ot, nt = tab(tr)
            b[:, ot:ot + nt] += np.einsum('eq,eqi->ei', wq * D[k], Bt)
        return geo, b
    tot = np.zeros(geo.nitems)
    for _, _, k in integral.entries:
        tot += (wq * D[k]).sum(axis=1)
    return geo, tot


def assemble_matrix(program: FormProgram) -> np.ndarray:
    """CSR values (pattern = fes.pattern()) of a bilinear form (reference: ``a.Assemble()``, base_solver.py:374)."""
    fes = program.fes
    pat = fes.pattern()
    nloc = fes.nloc
    vals = np.zeros(pat.nnz)
    for integ in program.integrals:
        geo, A = element_tensors(program, integ)
        if integ.kind == 'ifacet':
            fpos = np.searchsorted(fes.mesh.interior_facets, integ.items)
            for s in (0, 1):
                idx = pat.cell2nnz[geo.cells[s]]
                blk = A[:, s * nloc:(s + 1) * nloc, s * nloc:(s + 1) * nloc].reshape(geo.nitems, -1)
                vals += np.bincount(idx.ravel(), weights=blk.ravel(), minlength=pat.nnz)
                idx = pat.facet2nnz[fpos, s]
                blk = A[:, s * nloc:(s + 1) * nloc, (1 - s) * nloc:(2 - s) * nloc].reshape(geo.nitems, -1)
                vals += np.bincount(idx.ravel(), weights=blk.ravel(), minlength=pat.nnz)
        else:
            idx = pat.cell2nnz[geo.cells[0]]
            vals += np.bincount(idx.ravel(), weights=A.reshape(geo.nitems, -1).ravel(), minlength=pat.nnz)
    return vals


def assemble_vector(program: FormProgram) -> np.ndarray:
    fes = program.fes
    nloc = fes.nloc
    out = np.zeros(fes.ndof)
    cd = fes.cell_dofs
    for integ in program.integrals:
        geo, b = element_tensors(program, integ)
        nside = 2 if integ.kind == 'ifacet' else 1
        for s in range(nside):
            out += np.bincount(cd[geo.cells[s]].ravel(), weights=b[:, s * nloc:(s + 1) * nloc].ravel(),
                               minlength=fes.ndof)
    return out


def integrate(program: FormProgram) -> float:
    """``ngs.Integrate(cf, mesh)`` (reference helpers/error.py:66-77)."""
    tot = 0.0
    for integ in program.integrals:
        _, v = element_tensors(program, integ)
        tot += float(v.sum())
    return tot


# ---- linear algebra ------------------------------------------------------------------------------------------------
def csr_matrix(fes, vals: np.ndarray) -> sp.csr_matrix:
    pat = fes.pattern()
    return sp.csr_matrix((vals, pat.colidx, pat.rowptr), shape=(pat.n, pat.n))


def solve_direct(A: sp.csr_matrix, b: np.ndarray, x: np.ndarray, free: np.ndarray) -> np.ndarray:
    """``gfu.vec += A_ff^-1 (L.vec - A gfu.vec)`` — reference base_model.py:918-922."""
    r = b - A @ x
    idx = np.nonzero(free)[0]
    Aff = A[idx][:, idx].tocsc()
    lu = spla.splu(Aff)
    rf = r[idx]
    d = lu.solve(rf)
    for _ in range(2):            # two steps of iterative refinement, PARDISO's default behaviour
        d += lu.solve(rf - Aff @ d)
    out = x.copy()
    out[idx] += d
    return out
