"""Index-level emulation (NumPy) of the smoother application (csrc/ocmp_patch.cu): ``k_patch_apply_stream`` — persistent
CTAs walking their patches as one sequence of column chunks through a ring of NSTAGE shared-memory stages, a thread
owning VEC = 16 / sizeof(T) consecutive rows and every VEC-th column, the column groups summed in a fixed order when the
patch ends — followed by ``k_patch_gather`` over the incidence list built by ``backend.patch_incidence``. Same chunk
geometry (host-side sizing of ``patch_apply_stream``), same chunk -> (patch, first column, byte offset) arithmetic, same
stage / parity schedule. The result must equal ``z[dofs] += A_p^-1 r[dofs]`` computed directly from the
(transposed-stored, rounded) inverses; the arithmetic itself is checked on the GPU."""
import numpy as np
import pytest
import torch

from opencmp_b200.backend import patch_incidence, _STORAGE_ALIGN

NSTAGE = 4
ELEM = {'fp64': 8, 'fp32': 4, 'bf16': 2}


def _round(a, storage):
    t = torch.from_numpy(np.ascontiguousarray(a))
    if storage == 'fp32':
        return t.to(torch.float32).to(torch.float64).numpy()
    if storage == 'bf16':
        return t.to(torch.bfloat16).to(torch.float64).numpy()
    return a


def chunk_geometry(bs, storage):
    """Host-side sizing in patch_apply_stream."""
    colbytes = bs * ELEM[storage]
    assert colbytes % 16 == 0, 'columns must stay 16-byte aligned for the bulk copies'
    cc = max(1, min(bs, 8192 // colbytes))
    nchunk = (bs + cc - 1) // cc
    stage_bytes = (cc * colbytes + 127) // 128 * 128
    return cc, nchunk, stage_bytes


def emulate_stream(npatch, bs, pdofs, inv_flat, r, storage, grid):
    """inv_flat: one flat array of npatch * bs * bs stored entries (A_p[j * bs + i] = (A_p^-1)_{ij})."""
    cc, nchunk, stage_bytes = chunk_geometry(bs, storage)
    es = ELEM[storage]
    vec = 16 // es
    assert bs % vec == 0
    nq = bs // vec
    nthreads = (bs + 31) // 32 * 32
    y = np.full(npatch * bs, np.nan)
    for block in range(min(grid, npatch)):
        mine = (npatch - block + grid - 1) // grid
        total = mine * nchunk
        stages = [None] * NSTAGE
        filled = [0] * NSTAGE                      # how often each stage's barrier completed a phase
        waited = [0] * NSTAGE

        def issue(g):
            k, c = divmod(g, nchunk)
            p = block + k * grid
            cols = bs - c * cc if c == nchunk - 1 else cc
            nbytes = cols * bs * es
            assert nbytes % 16 == 0 and nbytes <= stage_bytes
            src = p * bs * bs + c * cc * bs                        # element offset of the chunk
            assert (src * es) % 16 == 0
            s = g % NSTAGE
            stages[s] = inv_flat[src: src + cols * bs].copy()
            filled[s] += 1
        for g in range(min(NSTAGE, total)):
            issue(g)
        p = block
        rl = np.array([r[d] if d >= 0 else 0.0 for d in pdofs[p]])
        acc = np.zeros((nthreads, vec))
        part = np.full(vec * bs, np.nan)
        c = 0
        for g in range(total):
            s = g % NSTAGE
            rnext = None
            if c == 0 and p + grid < npatch:
                rnext_vals = np.array([r[d] if d >= 0 else 0.0 for d in pdofs[p + grid]])
            parity = (g // NSTAGE) & 1
            assert filled[s] == g // NSTAGE + 1 and (filled[s] - 1) & 1 == parity      # the phase being waited on
            waited[s] += 1
            cols = bs - c * cc if c == nchunk - 1 else cc
            A = stages[s]
            for tid in range(nthreads):
                q, grp = tid % nq, tid // nq
                if grp >= vec:                         # threads of the rounded-up last warp do not compute
                    assert tid >= bs
                    continue
                for j in range(grp, cols, vec):
                    off = j * bs * es + q * 16         # byte offset of this thread's 128-bit load in the stage
                    assert off % 16 == 0 and off + 16 <= cols * bs * es
                    for v in range(vec):
                        acc[tid, v] += A[off // es + v] * rl[c * cc + j]
                if c == nchunk - 1:
                    part[grp * bs + q * vec: grp * bs + (q + 1) * vec] = acc[tid]
            if g + NSTAGE < total:
                issue(g + NSTAGE)
            c += 1
            if c == nchunk:
                assert not np.isnan(part).any()
                for tid in range(bs):
                    y[p * bs + tid] = sum(part[v * bs + tid] for v in range(vec))
                if p + grid < npatch:
                    rl = rnext_vals
                acc = np.zeros((nthreads, vec))
                part = np.full(vec * bs, np.nan)
                c = 0
                p += grid
    return y


def emulate_gather(n, inc_ptr, inc_idx, y, w, m, scale, z, accumulate):
    out = z.copy()
    for d in range(n):
        s = 0.0
        for k in range(inc_ptr[d], inc_ptr[d + 1]):
            s += y[inc_idx[k]]
        s *= w[d] * m[d] * scale
        out[d] = out[d] + s if accumulate else s
    return out


def _direct(pdofs, inv, r, n, bs):
    z = np.zeros(n)
    for p, d in enumerate(pdofs):
        rl = np.array([r[dj] if dj >= 0 else 0.0 for dj in d])
        A = inv[p].reshape(bs, bs)                    # A[j, i]
        for i, di in enumerate(d):
            if di >= 0:
                z[di] += A[:, i] @ rl
    return z


def _case(npatch, nvalid, storage, seed, n=211):
    rng = np.random.default_rng(seed)
    align = _STORAGE_ALIGN[storage]
    bs = nvalid + (-nvalid % align)
    pdofs = -np.ones((npatch, bs), dtype=np.int32)
    inv = np.zeros((npatch, bs * bs))
    for p in range(npatch):
        k = nvalid - (p % 3)                          # ragged: some patches carry extra padding
        pdofs[p, :k] = np.sort(rng.choice(n, k, replace=False))
        M = rng.standard_normal((bs, bs))
        M[k:, :] = 0.0
        M[:, k:] = 0.0
        M[np.arange(k, bs), np.arange(k, bs)] = 1.0   # identity rows / columns on the padding, like the set-up kernel
        inv[p] = _round(M.ravel(), storage)
    r = rng.standard_normal(n)
    return bs, pdofs, inv, r, n


@pytest.mark.parametrize('storage,nvalid,npatch,grid', [
    ('fp64', 13, 7, 3), ('fp64', 132, 5, 2), ('fp32', 90, 6, 4), ('fp32', 132, 3, 5), ('bf16', 89, 5, 2),
    ('fp64', 7, 9, 2)])
def test_stream_apply_and_gather_match_direct_patch_products(storage, nvalid, npatch, grid):
    bs, pdofs, inv, r, n = _case(npatch, nvalid, storage, seed=nvalid + npatch)
    y = emulate_stream(npatch, bs, pdofs, inv.ravel(), r, storage, grid)
    assert not np.isnan(y).any()                      # every patch-local entry was written exactly once
    inc_ptr, inc_idx = patch_incidence(pdofs, n)
    assert inc_ptr[-1] == (pdofs >= 0).sum()
    for d in range(n):                                # ascending positions = fixed summation order
        seg = inc_idx[inc_ptr[d]: inc_ptr[d + 1]]
        assert (np.diff(seg) > 0).all() and (pdofs.ravel()[seg] == d).all()
    w = np.random.default_rng(1).uniform(0.2, 1.0, n)
    m = (np.random.default_rng(2).uniform(size=n) > 0.2).astype(float)
    ref = _direct(pdofs, inv, r, n, bs)
    z0 = np.random.default_rng(3).standard_normal(n)
    got = emulate_gather(n, inc_ptr, inc_idx, y, w, m, 0.7, z0, False)
    assert np.abs(got - 0.7 * w * m * ref).max() < 1e-12 * max(1.0, np.abs(ref).max())
    got = emulate_gather(n, inc_ptr, inc_idx, y, w, m, 0.7, z0, True)
    assert np.abs(got - (z0 + 0.7 * w * m * ref)).max() < 1e-12 * max(1.0, np.abs(ref).max())


def test_chunk_geometry_fits_the_shared_memory_budget():
    for storage, align in _STORAGE_ALIGN.items():
        for bs in range(align, 257, align):
            cc, nchunk, stage_bytes = chunk_geometry(bs, storage)
            assert 1 <= cc <= bs and (nchunk - 1) * cc < bs <= nchunk * cc
            vec = 16 // ELEM[storage]
            smem = 128 + (1 + vec) * ((bs * 8 + 127) // 128 * 128) + NSTAGE * stage_bytes
            assert smem <= 64 * 1024, (storage, bs, smem)


def test_packed_bfloat16_pairs_widen_by_bit_moves():
    """RowVec<__nv_bfloat16>::load: a 32-bit word holds two stored entries; the low one becomes an FP32 by a left shift
    of 16 bits, the high one by masking the low half away — the same values torch's bfloat16 -> float32 conversion gives
    (little-endian packing: entry 2k in the low half of word k)."""
    rng = np.random.default_rng(5)
    vals = torch.from_numpy(rng.standard_normal(64) * 10.0 ** rng.integers(-12, 12, 64)).to(torch.bfloat16)
    vals[3], vals[10] = 0.0, -0.0
    words = vals.view(torch.int16).numpy().view(np.uint16).astype(np.uint32)
    packed = words[0::2] | (words[1::2] << 16)                       # what one LDS.128 returns, word by word
    lo = (packed << 16).astype(np.uint32).view(np.float32)
    hi = (packed & np.uint32(0xffff0000)).view(np.float32)
    ref = vals.to(torch.float32).numpy()
    assert np.array_equal(lo.view(np.uint32), ref[0::2].view(np.uint32))
    assert np.array_equal(hi.view(np.uint32), ref[1::2].view(np.uint32))
