"""Thread-by-thread emulation (NumPy) of the index logic of the reduced-precision smoother kernels
``k_patch_apply_f32`` and ``k_patch_apply_bf16`` (csrc/ocmp_patch.cu), written while no GPU was available: every lane
follows the same column / row / chunk assignment, unrolled main loop, remainder loop and partial-sum layout as the CUDA
source, phase by phase between the barriers. The result must equal ``z[dofs] += A_p^-1 r[dofs]`` computed directly
from the (transposed-stored, rounded) inverses. This pins the thread mapping; the arithmetic itself is checked on the
GPU (tests/test_zzz_gpu_unmeasured_kernels.py)."""
import numpy as np
import pytest
import torch


def _round(a, storage):
    t = torch.from_numpy(np.ascontiguousarray(a))
    if storage == 'fp32':
        return t.to(torch.float32).to(torch.float64).numpy()
    return t.to(torch.bfloat16).to(torch.float64).numpy()


def emulate_f32(npatch, bs, pdofs, inv, r, n, MR, NW, UC, grid):
    """k_patch_apply_f32<MR, NW, UC>: lane owns rows 4*lane .. +3 (+ 128 m), one column per warp and iteration slot."""
    z = np.zeros(n)
    for block in range(grid):
        for p in range(block, npatch, grid):
            d = pdofs[p]
            rl = np.array([r[dj] if dj >= 0 else 0.0 for dj in d])
            A = inv[p]                                            # flat: A[j * bs + i]
            part = np.zeros((NW, bs))
            for warp in range(NW):
                for lane in range(32):
                    s = np.zeros((MR, 4))
                    j = warp
                    while j + (UC - 1) * NW < bs:
                        for u in range(UC):
                            col = (j + u * NW) * bs
                            for m in range(MR):
                                i4 = lane + 32 * m
                                if 4 * i4 < bs:
                                    s[m] += A[col + 4 * i4: col + 4 * i4 + 4] * rl[j + u * NW]
                        j += UC * NW
                    while j < bs:
                        for m in range(MR):
                            i4 = lane + 32 * m
                            if 4 * i4 < bs:
                                s[m] += A[j * bs + 4 * i4: j * bs + 4 * i4 + 4] * rl[j]
                        j += NW
                    for m in range(MR):
                        i = 4 * (lane + 32 * m)
                        if i < bs:
                            part[warp, i:i + 4] = s[m]
            for i in range(bs):
                if d[i] >= 0:
                    z[d[i]] += part[:, i].sum()
    return z


def emulate_bf16(npatch, bs, pdofs, inv, r, n, MR, NW, UC, grid):
    """k_patch_apply_bf16<MR, NW, UC>: a half warp per column group, lane owns rows 8*(lane % 16) .. +7 (+ 128 m)."""
    G = 2 * NW
    z = np.zeros(n)
    for block in range(grid):
        for p in range(block, npatch, grid):
            d = pdofs[p]
            rl = np.array([r[dj] if dj >= 0 else 0.0 for dj in d])
            A = inv[p]
            part = np.zeros((G, bs))
            for warp in range(NW):
                for lane in range(32):
                    l16, g = lane & 15, 2 * warp + (lane >> 4)
                    s = np.zeros((MR, 8))
                    j = g
                    while j + (UC - 1) * G < bs:
                        for u in range(UC):
                            col = (j + u * G) * bs
                            for m in range(MR):
                                i8 = l16 + 16 * m
                                if 8 * i8 < bs:
                                    s[m] += A[col + 8 * i8: col + 8 * i8 + 8] * rl[j + u * G]
                        j += UC * G
                    while j < bs:
                        for m in range(MR):
                            i8 = l16 + 16 * m
                            if 8 * i8 < bs:
                                s[m] += A[j * bs + 8 * i8: j * bs + 8 * i8 + 8] * rl[j]
                        j += G
                    for m in range(MR):
                        i = 8 * (l16 + 16 * m)
                        if i < bs:
                            part[g, i:i + 8] = s[m]
            for i in range(bs):
                if d[i] >= 0:
                    z[d[i]] += part[:, i].sum()
    return z


def _case(bs, nreal, npatch=3, n=400, seed=0):
    rng = np.random.default_rng(seed)
    pdofs = -np.ones((npatch, bs), dtype=np.int64)
    for p in range(npatch):
        pdofs[p, :nreal] = np.sort(rng.choice(n, nreal, replace=False))
    inv = rng.uniform(-1, 1, (npatch, bs * bs))
    r = rng.uniform(-1, 1, n)
    return pdofs, inv, r, n


def _direct(pdofs, inv, r, n, bs):
    z = np.zeros(n)
    for p in range(pdofs.shape[0]):
        d = pdofs[p]
        ok = d >= 0
        rl = np.where(ok, r[np.maximum(d, 0)], 0.0)
        At = inv[p].reshape(bs, bs)                               # At[j, i] = (A^-1)_{ij}
        zl = At.T @ rl
        np.add.at(z, d[ok], zl[ok])
    return z


# (bs, real dofs, MR) as launched by ocmp_patch_apply_cta_f32 / _bf16: MR = 1 up to 128 rows, 2 above
@pytest.mark.parametrize('bs,nreal', [(92, 89), (132, 132), (48, 41), (128, 128), (256, 250)])
def test_f32_kernel_thread_mapping(bs, nreal):
    pdofs, inv, r, n = _case(bs, nreal)
    inv = _round(inv, 'fp32')
    MR, UC = (1, 4) if bs <= 128 else (2, 2)
    got = emulate_f32(pdofs.shape[0], bs, pdofs, inv, r, n, MR, 4, UC, grid=2)
    ref = _direct(pdofs, inv, r, n, bs)
    assert np.abs(got - ref).max() < 1e-12 * max(1.0, np.abs(ref).max())


@pytest.mark.parametrize('bs,nreal', [(96, 89), (136, 132), (48, 41), (128, 128), (256, 250)])
def test_bf16_kernel_thread_mapping(bs, nreal):
    pdofs, inv, r, n = _case(bs, nreal, seed=1)
    inv = _round(inv, 'bf16')
    MR = 1 if bs <= 128 else 2
    got = emulate_bf16(pdofs.shape[0], bs, pdofs, inv, r, n, MR, 4, 4, grid=2)
    ref = _direct(pdofs, inv, r, n, bs)
    assert np.abs(got - ref).max() < 1e-12 * max(1.0, np.abs(ref).max())


def test_bf16_word_unpacking_convention():
    """bf16x2_fma takes element 2k of a 16-byte load from the LOW half of word k and element 2k+1 from the HIGH half
    (little endian), and widens a bf16 by a 16-bit shift."""
    vals = torch.tensor([1.5, -2.25, 3.0e-8, 1.0e10], dtype=torch.float64).to(torch.bfloat16)
    words = vals.view(torch.int16).numpy().view(np.uint16).astype(np.uint32)
    w0 = words[0] | (words[1] << 16)                              # as the 32-bit load sees two consecutive elements
    lo = np.array([w0 << 16], dtype=np.uint32).view(np.float32)[0]
    hi = np.array([w0 & 0xffff0000], dtype=np.uint32).view(np.float32)[0]
    assert lo == float(vals[0]) and hi == float(vals[1])
