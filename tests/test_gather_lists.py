"""Contributor lists of the deterministic (two-phase) assembly, built on the host side of the CUDA backend
(``CudaBackend._gather_lists``; host tensors stand in for device buffers): every valid local tile entry / local vector
entry appears exactly once, the targets are strictly increasing (one writer per non-zero / dof), and scattering a known
item-local buffer through the lists reproduces the scatter through the element -> nnz map / the cell dof lists."""
import numpy as np
import pytest
import torch

import cases
import test_gpu_paths_dry as T


def _plans(build):
    import opencmp_b200.ngs as ngs
    be = T.DryCudaBackend()
    old = ngs._backend
    ngs.set_backend(be)
    try:
        c = build()
        progs = (c['a'].program(), c['L'].program())
        return be, c, progs, [be._plans(p) for p in progs]
    finally:
        ngs.set_backend(old)


CASES = {
    'poisson_h1_p2': lambda: cases.poisson(cases.square_mesh(4), 2, False),
    'stokes_hdiv_dg_p2': lambda: cases.stokes(cases.channel_mesh(4), 2, True),
    'stokes_3d_hex': lambda: cases.stokes_3d('hex', 2),
}


@pytest.mark.parametrize('name', sorted(CASES))
def test_matrix_gather_lists_equal_the_scatter_map(name):
    be, c, (pa, pl), (plans_a, plans_l) = _plans(CASES[name])
    fes = c['fes']
    pat = fes.pattern()
    nloc, n2 = fes.nloc, fes.nloc ** 2
    for integ, plan in zip(pa.integrals, plans_a):
        xp = plan['contract']
        n = plan['nitems']
        seg_ptr, seg_tgt, order = (a.numpy().astype(np.int64) for a in be._gather_lists(pa, integ, plan, 0, n, 'matrix'))
        assert (np.diff(seg_tgt) > 0).all()                       # one segment per touched non-zero, ascending
        assert (np.diff(seg_ptr) > 0).all() and seg_ptr[0] == 0 and seg_ptr[-1] == len(order)
        assert len(np.unique(order)) == len(order)                # every buffer entry is used at most once
        # reference: walk the tiles the way the kernel's atomicAdd epilogue does
        tiles = plan['tables']['tiles']
        rng = np.random.default_rng(0)
        abuf = rng.standard_normal(n * xp.ntiles * 16)
        ref = np.zeros(pat.nnz)
        items = np.arange(n) if integ.items is None else np.asarray(integ.items)
        fc = fes.mesh.facet_cells
        nvalid = 0
        for it in range(n):
            cells = (items[it], items[it]) if integ.kind == 'cell' else fc[items[it]]
            for t, (boff, j0, s0, ns, sides, row, col, rem) in enumerate(tiles):
                st, su = sides & 1, (sides >> 1) & 1
                for a in range(rem & 0xff):
                    for b in range(rem >> 8):
                        ij = (row + a) * nloc + col + b
                        pos = pat.cell2nnz.reshape(-1)[cells[st] * n2 + ij] if st == su else \
                            pat.facet2nnz.reshape(-1)[(it * 2 + st) * n2 + ij]
                        ref[pos] += abuf[(it * xp.ntiles + t) * 16 + a * 4 + b]
                        nvalid += 1
        assert len(order) == nvalid == n * xp.nact
        got = np.zeros(pat.nnz)
        for s in range(len(seg_tgt)):
            got[seg_tgt[s]] += abuf[order[seg_ptr[s]:seg_ptr[s + 1]]].sum()
        assert np.abs(got - ref).max() < 1e-12


@pytest.mark.parametrize('name', sorted(CASES))
def test_vector_gather_lists_equal_the_cell_dof_lists(name):
    be, c, (pa, pl), (plans_a, plans_l) = _plans(CASES[name])
    fes = c['fes']
    cd = fes.cell_dofs
    seen = 0
    for integ, plan in zip(pl.integrals, plans_l):
        xp = plan['contract']
        n = plan['nitems']
        if n == 0:
            continue
        seg_ptr, seg_tgt, order = (a.numpy().astype(np.int64) for a in be._gather_lists(pl, integ, plan, 0, n, 'vector'))
        assert (np.diff(seg_tgt) > 0).all() and seg_ptr[-1] == len(order)
        items = np.arange(n) if integ.items is None else np.asarray(integ.items)
        fc = fes.mesh.facet_cells
        lbuf = np.random.default_rng(1).standard_normal(n * xp.nside * fes.nloc)
        ref = np.zeros(fes.ndof)
        for it in range(n):
            for s in range(xp.nside):
                cell = items[it] if integ.kind == 'cell' else fc[items[it]][s]
                if cell < 0:
                    continue
                np.add.at(ref, cd[cell], lbuf[(it * xp.nside + s) * fes.nloc:(it * xp.nside + s + 1) * fes.nloc])
        got = np.zeros(fes.ndof)
        for s in range(len(seg_tgt)):
            got[seg_tgt[s]] += lbuf[order[seg_ptr[s]:seg_ptr[s + 1]]].sum()
        assert np.abs(got - ref).max() < 1e-12
        seen += 1
    assert seen > 0
