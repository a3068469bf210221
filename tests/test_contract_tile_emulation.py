"""Emulation (NumPy) of the tile arithmetic of k_contract (csrc/ocmp_assembly.cu) from the descriptor tables
``backend.contract_tables`` builds: Z rows from (zdesc, ent), 4 x 4 register tiles from (tiles, seg), scatter positions
from the tile's row / column. Random physical tables B and coefficient values D stand in for one quadrature point of
one item; the result must equal the direct evaluation of the lowered integrand
``A[(st, i), (su, j)] = sum_entries D[slot] B[st][bt][rt][i] B[su][bu][ru][j]``. Pins the descriptor construction
(padding, offsets, segment ranges) without a GPU; the kernel arithmetic is compared with the oracle on the GPU."""
import numpy as np
import pytest

from opencmp_b200.backend import contract_tables

pad4 = lambda v: (v + 3) // 4 * 4


def _setup(nloc, nrows, nside, nent, seed):
    rng = np.random.default_rng(seed)
    nblk = len(nloc)
    ro = np.concatenate([[0], np.cumsum(nrows)])
    nrows_tot = int(ro[-1])

    def decode(row):
        side, rr = divmod(int(row), nrows_tot)
        b = max(i for i in range(nblk) if ro[i] <= rr)
        return side, b, rr - int(ro[b])
    sb_off, off = [], 0
    for b in range(nblk):
        sb_off.append(off)
        off += nrows[b] * pad4(nloc[b])
    sbsz = off
    loc_off = [int(v) for v in np.concatenate([[0], np.cumsum(nloc)[:-1]])]
    entries = []
    for slot in range(nent):
        tr = int(rng.integers(0, nside * nrows_tot))
        ur = int(rng.integers(0, nside * nrows_tot))
        entries.append((tr, ur, slot % max(1, nent // 2)))           # slots are shared between entries
    tb = contract_tables(entries, decode, nloc, [0] * nblk, nrows, [0] * nblk, sb_off, sbsz, loc_off, nside)
    return rng, decode, sb_off, sbsz, loc_off, entries, tb


def emulate(tb, sB, D, nside, nloc_tot):
    """One quadrature point: phases 2 and 3 of the kernel, then the scatter into the dense local matrix."""
    sZ = np.zeros(tb['zsz'])
    for k0, k1, bbase, zi in tb['zdesc']:
        sZ[zi] = sum(D[tb['ent'][k, 0]] * sB[bbase + tb['ent'][k, 1]] for k in range(k0, k1))
    A = np.zeros((nside * nloc_tot, nside * nloc_tot))
    for boff, j0, s0, ns, sides, row, col, rem in tb['tiles']:
        acc = np.zeros((4, 4))
        for sg in range(s0, s0 + ns):
            ro_, zo = tb['seg'][sg]
            b = sB[boff + ro_: boff + ro_ + 4]
            z = sZ[j0 + zo: j0 + zo + 4]
            acc += np.outer(b, z)
        st, su = sides & 1, (sides >> 1) & 1
        ni, nj = rem & 0xff, rem >> 8
        for a in range(ni):
            for c in range(nj):
                A[st * nloc_tot + row + a, su * nloc_tot + col + c] += acc[a, c]
    return A


@pytest.mark.parametrize('nloc,nrows,nside,nent', [
    ([27, 27, 27, 8], [4, 4, 4, 4], 1, 40), ([20, 6], [6, 3], 2, 60), ([6], [3], 1, 5), ([10, 3], [3, 3], 2, 25),
    ([1], [3], 2, 4)])
def test_tile_tables_reproduce_the_lowered_integrand(nloc, nrows, nside, nent):
    rng, decode, sb_off, sbsz, loc_off, entries, tb = _setup(nloc, nrows, nside, nent, seed=sum(nloc) + nent)
    nloc_tot = sum(nloc)
    # physical tables with the kernel's padded layout; the padding is zero like in shared memory
    sB = np.zeros(nside * sbsz + 8)
    Bfull = {}
    for s in range(nside):
        for b, nl in enumerate(nloc):
            for r in range(nrows[b]):
                v = rng.standard_normal(nl)
                Bfull[(s, b, r)] = v
                o = s * sbsz + sb_off[b] + r * pad4(nl)
                sB[o: o + nl] = v
    D = rng.standard_normal(nent)
    got = emulate(tb, sB, D, nside, nloc_tot)
    ref = np.zeros_like(got)
    for tr, ur, slot in entries:
        st, bt, rt = decode(tr)
        su, bu, ru = decode(ur)
        ref[st * nloc_tot + loc_off[bt]: st * nloc_tot + loc_off[bt] + nloc[bt],
            su * nloc_tot + loc_off[bu]: su * nloc_tot + loc_off[bu] + nloc[bu]] += \
            D[slot] * np.outer(Bfull[(st, bt, rt)], Bfull[(su, bu, ru)])
    assert np.abs(got - ref).max() < 1e-12 * max(1.0, np.abs(ref).max())
    assert tb['nact'] == int(((tb['tiles'][:, 7] & 0xff) * (tb['tiles'][:, 7] >> 8)).sum())
    assert tb['a_fma'] > 0 and tb['z_fma'] > 0
