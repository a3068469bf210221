"""Diffuse-interface phase-field generation on the device.

Same functions, arguments and return values as the voxel pipeline of reference
``opencmp/diffuse_interface/interface.py`` — ``get_binary_2d`` (:31-57), ``get_phi`` (:137-180) — and the ``edt.edt``
call inside it (:167), computed by csrc/ocmp_dim.cu: ray tracing of every grid node, 3^d erosion, exact Euclidean
distance transform, erf profile. Arrays go in and come out as NumPy arrays like in the reference (its callers index
them on the host, ``helpers/ngsolve_.py:212-296``); everything in between stays on the device. One-off per run,
repeated per time step only with rigid-body motion.

Needs the CUDA backend (no CPU fallback here: CPU runs keep the SciPy stand-in of ``netgen_shim.edt``)."""
from __future__ import annotations

from typing import List, Sequence

import numpy as np


def _cuda():
    from . import ngs
    be = ngs._backend
    if getattr(be, 'name', '') != 'cuda':
        raise RuntimeError('opencmp_b200.dimgen needs the CUDA backend (ngs.set_backend(CudaBackend()))')
    return be


def get_binary_2d(boundary_lst: Sequence, N: List[int], scale: List[float], offset: List[float]) -> np.ndarray:
    """interface.py:31-57: 1 at the grid nodes inside the polygon, 0 elsewhere; shape (N[0] + 1, N[1] + 1)."""
    be = _cuda()
    t = be.torch
    n0, n1 = int(N[0] + 1), int(N[1] + 1)
    poly = be._up(np.asarray(boundary_lst, dtype=np.float64).reshape(-1, 2))
    out = t.empty(n0 * n1, dtype=t.float64, device=be.device)
    be._ck(be.lib.ocmp_dim_raytrace_2d(n0, n1, float(scale[0]), float(scale[1]), float(offset[0]), float(offset[1]),
                                       int(N[0]), int(N[1]), int(poly.shape[0]), poly.data_ptr(), out.data_ptr(),
                                       be._stream()))
    be.launches += 1
    return be.to_numpy(out).reshape(n0, n1)


def _shape3(shape):
    if len(shape) not in (2, 3):
        raise ValueError('Only works with 2D or 3D meshes.')
    return (int(shape[0]), int(shape[1]), int(shape[2]) if len(shape) == 3 else 1)


def edt(data, anisotropy=None, black_border: bool = False, **_) -> np.ndarray:
    """``edt.edt(array)``: Euclidean distance (FP32, in voxels) of every non-zero voxel to the nearest zero voxel."""
    if anisotropy is not None or black_border:
        raise NotImplementedError('edt on the device: isotropic voxels without a black border (what interface.py uses)')
    be = _cuda()
    t = be.torch
    a = np.ascontiguousarray(np.asarray(data) != 0, dtype=np.uint8)
    n0, n1, n2 = _shape3(a.shape)
    fg = be._up(a.reshape(-1))
    wa = t.empty(a.size, dtype=t.int64, device=be.device)
    wb = t.empty(a.size, dtype=t.int64, device=be.device)
    dist = t.empty(a.size, dtype=t.float32, device=be.device)
    be._ck(be.lib.ocmp_dim_edt(a.ndim, n0, n1, n2, fg.data_ptr(), wa.data_ptr(), wb.data_ptr(), dist.data_ptr(),
                               be._stream()))
    be.launches += a.ndim + 2
    return dist.cpu().numpy().reshape(a.shape)


def get_phi(binary: np.ndarray, lmbda: float, N: List[int], scale: List[float], offset: List[float], dim: int = 2) \
        -> np.ndarray:
    """interface.py:137-180: phase field running from 1 inside the geometry to 0 outside with an erf profile of width
    ``lmbda`` across the border of ``binary``. border -> distance transform -> profile without leaving the device."""
    if dim not in (2, 3):
        raise ValueError('Only works with 2D or 3D meshes.')
    be = _cuda()
    t = be.torch
    b = np.ascontiguousarray(binary, dtype=np.float64)
    if b.ndim != dim:
        raise ValueError('binary must be a {}-dimensional array'.format(dim))
    n0, n1, n2 = _shape3(b.shape)
    bd = be._up(b.reshape(-1))
    fg = t.empty(b.size, dtype=t.uint8, device=be.device)
    wa = t.empty(b.size, dtype=t.int64, device=be.device)
    wb = t.empty(b.size, dtype=t.int64, device=be.device)
    dist = t.empty(b.size, dtype=t.float32, device=be.device)
    phi = t.empty(b.size, dtype=t.float64, device=be.device)
    st = be._stream()
    be._ck(be.lib.ocmp_dim_border(dim, n0, n1, n2, bd.data_ptr(), fg.data_ptr(), st))
    be._ck(be.lib.ocmp_dim_edt(dim, n0, n1, n2, fg.data_ptr(), wa.data_ptr(), wb.data_ptr(), dist.data_ptr(), st))
    h = min(scale) / min(N)
    be._ck(be.lib.ocmp_dim_phi(b.size, dist.data_ptr(), bd.data_ptr(), float(h), float(lmbda), phi.data_ptr(), st))
    be.launches += dim + 4
    return be.to_numpy(phi).reshape(b.shape)


def rigid_body_motion(values: np.ndarray, inv_R: np.ndarray, N: List[int], scale: List[float], offset: List[float]) \
        -> np.ndarray:
    """The node array ``tmp_arr`` of reference helpers/ngsolve_.py:236-262 / :264-292 (gridfunction_rigid_body_motion):
    ``values[i, j(, k)]`` = the field at node x = -offset + scale * (i, j, k) / N; returns the field carried along the
    rotation, out[node] = field(inv_R x_node) (multilinear interpolation on the grid — exact for the multilinear
    GridFunctions ``numpy_to_NGSolve`` produces), 1 where the pre-image lies outside the box. One launch instead of the
    reference's Python loop over all nodes with a point search per node."""
    import ctypes as C
    be = _cuda()
    t = be.torch
    v = np.ascontiguousarray(values, dtype=np.float64)
    dim = v.ndim
    if dim not in (2, 3) or tuple(v.shape) != tuple(int(n) + 1 for n in N[:dim]):
        raise ValueError('values must have shape N + 1 in 2 or 3 dimensions')
    R = np.ascontiguousarray(np.asarray(inv_R, dtype=np.float64).reshape(dim, dim))
    vd = be._up(v.reshape(-1))
    out = t.empty(v.size, dtype=t.float64, device=be.device)
    arr = lambda a: (C.c_double * len(a))(*[float(x) for x in a])
    n0, n1, n2 = _shape3(v.shape)
    be._ck(be.lib.ocmp_dim_rigid_motion(dim, n0, n1, n2, arr(scale[:dim]), arr(offset[:dim]), arr(R.ravel()),
                                        vd.data_ptr(), out.data_ptr(), be._stream()))
    be.launches += 1
    return be.to_numpy(out).reshape(v.shape)


def gridfunction_rigid_body_motion(t, orig_gfu, gfu, inv_R, mesh, N, scale, offset, fallback=None):
    """Drop-in for reference helpers/ngsolve_.py:212-296 (same arguments and return value). When ``orig_gfu`` is a
    multilinear H1 field whose mesh vertices are the nodes of the N grid — what ``numpy_to_NGSolve`` / the DIM
    pre-processing produce on the default quadrilateral / hexahedral meshes — the node loop with a point search per node
    becomes one kernel launch (``rigid_body_motion``); otherwise ``fallback`` (the reference's own function) is called."""
    from . import ngs
    dim = mesh.dim
    fes = orig_gfu.space
    ok = len(fes.blocks) == 1 and fes.blocks[0].family == 'H1' and fes.mesh.cell_type in ('quad', 'hex')
    grid = None
    if ok:
        vals = ngs.get_backend().to_numpy(orig_gfu.vec.a)
        nv = fes.mesh.nv
        ok = not vals[nv:].size or float(np.abs(vals[nv:]).max()) <= 1e-12 * max(1.0, float(np.abs(vals[:nv]).max()))
    if ok:
        pts = fes.mesh.points
        idx = []
        for a in range(dim):
            u = (pts[:, a] + offset[a]) / scale[a] * N[a]
            k = np.rint(u)
            ok = ok and float(np.abs(u - k).max()) < 1e-8 and k.min() >= 0 and k.max() <= N[a]
            idx.append(k.astype(np.int64))
        if ok and len(pts) == int(np.prod([n + 1 for n in N[:dim]])):
            grid = np.empty(tuple(int(n) + 1 for n in N[:dim]))
            grid[tuple(idx)] = vals[:nv]
    if grid is None:
        if fallback is None:
            raise NotImplementedError('rigid-body motion on the device needs a multilinear field on the N grid')
        return fallback(t, orig_gfu, gfu, inv_R, mesh, N, scale, offset)
    moved = rigid_body_motion(grid, np.asarray(inv_R(t.Get())), N, scale, offset)
    lo = tuple(-offset[a] for a in range(dim))
    hi = tuple(scale[a] - offset[a] for a in range(dim))
    gfu.Set(ngs.VoxelCoefficient(lo, hi, moved.transpose(), linear=True))       # ngsolve_.py:260 / :288
    return gfu
