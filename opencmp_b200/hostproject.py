"""Host-side ``GridFunction.Set`` for data that do not live on the target mesh (DIM pre-processing).

The device projection (project.py) evaluates its integrand with the assembly kernels, which only know fields on the
mesh being integrated over. The reference's diffuse-interface set-up also projects

* voxel arrays (``VoxelCoefficient``: phase fields / masks computed on a regular grid) whose grid is not the mesh, and
* GridFunctions of a *finer* structured mesh onto the simulation mesh (``phi_gfu.Set(phi_gfu_tmp)``,
  ``Set(Grad(phi_gfu_tmp))``, ``Set(Norm(Grad(phi_gfu_tmp)))``; reference diffuse_interface/dim.py:393-409),

once per run. Both are handled here with NumPy by the same algorithm as the device path — local L2 projection with the
order-2p rule, then averaging of shared DOFs — with the integrand evaluated point-wise: voxel data by multilinear
interpolation, foreign GridFunctions by locating the quadrature points in their own mesh.
"""
from __future__ import annotations

import numpy as np

from .ir import coef_leaves
from .quadrature import cell_rule
from .vtk import _NP_BINARY, _NP_UNARY, _erf


def locate_points(mesh, pts: np.ndarray):
    """(cell index, reference coordinates) of every point. Vectorised for meshes of axis-aligned boxes (the structured
    quad / hex meshes of the DIM generator); point-by-point search otherwise."""
    d = mesh.dim
    pts = np.asarray(pts, dtype=np.float64)
    if mesh.cell_type in ('quad', 'hex'):
        J = mesh.jacobians()
        off = J - np.einsum('eii->ei', J)[:, :, None] * np.eye(d)[None]
        if np.abs(off).max() < 1e-12 * np.abs(J).max():
            P = mesh.points
            axes = [np.unique(np.round(P[:, a], 12)) for a in range(d)]
            lo = mesh.origins()
            idx = [np.searchsorted(axes[a], np.round(lo[:, a], 12)) for a in range(d)]
            table = -np.ones([len(ax) for ax in axes], dtype=np.int64)
            table[tuple(idx)] = np.arange(mesh.ne)
            pi = [np.clip(np.searchsorted(axes[a], pts[:, a], side='right') - 1, 0, len(axes[a]) - 2)
                  for a in range(d)]
            cells = table[tuple(pi)]
            if (cells < 0).any():
                raise ValueError('locate_points: a point lies in a hole of the structured mesh')
            h = np.einsum('eii->ei', J)[cells]
            ref = (pts - lo[cells]) / h
            if ref.min() < -1e-9 or ref.max() > 1 + 1e-9:
                raise ValueError('locate_points: a point lies outside the mesh')
            return cells, np.clip(ref, 0.0, 1.0)
    from .ngs import MeshPoint
    cells = np.empty(len(pts), dtype=np.int64)
    ref = np.empty((len(pts), d))
    for k, p in enumerate(pts):
        c, xi = MeshPoint(mesh, tuple(p)).locate()
        cells[k], ref[k] = c, xi
    return cells, ref


def evaluate_at_points(cf, pts: np.ndarray) -> np.ndarray:
    """(ncomp, npts) values of a coefficient function (coordinates, parameters, functions, GridFunctions of any mesh)."""
    pts = np.asarray(pts, dtype=np.float64)
    n, d = pts.shape
    cache, fcache, lcache = {}, {}, {}

    def field(gf, blk):
        key = (id(gf), blk)
        if key in fcache:
            return fcache[key]
        fes = gf.space
        mesh = fes.mesh
        if id(mesh.points) not in lcache:
            lcache[id(mesh.points)] = locate_points(mesh, pts)
        cells, ref = lcache[id(mesh.points)]
        b = fes.blocks[blk]
        lo = fes.loc_offsets[blk]
        coef = np.asarray(gf.vec_numpy())[fes.cell_dofs[cells, lo:lo + b.nloc]]            # (n, nloc)
        out = None
        J = mesh.jacobians()
        for s in range(0, n, 65536):                                                       # bound the table size
            sl = slice(s, min(n, s + 65536))
            tab = b.basis.tabulate(ref[sl])                                                # (m, nrows, nloc)
            r = np.einsum('krl,kl->kr', tab, coef[sl])
            Jc = J[cells[sl]]
            if b.kind == 'scalar':
                g = np.einsum('kab,ka->kb', np.linalg.inv(Jc), r[:, 1:])
                phys = np.concatenate([r[:, :1], g], axis=1)
            else:
                det = np.linalg.det(Jc)
                val = np.einsum('kia,ka->ki', Jc, r[:, :d]) / det[:, None]
                gg = np.einsum('kia,kab,kbj->kij', Jc, r[:, d:].reshape(-1, d, d), np.linalg.inv(Jc)) / det[:, None, None]
                phys = np.concatenate([val, gg.reshape(-1, d * d)], axis=1)
            out = phys if out is None else np.concatenate([out, phys], axis=0)
        fcache[key] = out
        return out

    def ev(c):
        if id(c) in cache:
            return cache[id(c)]
        if c.op == 'const':
            v = np.full(n, float(c.val))
        elif c.op == 'param':
            v = np.full(n, float(c.val.Get()))
        elif c.op == 'coord':
            v = pts[:, c.val] if c.val < d else np.zeros(n)
        elif c.op == 'field':
            gf, blk, row, _side = c.val
            v = field(gf, blk)[:, row]
        elif c.op == 'ifpos':
            v = np.where(ev(c.args[0]) > 0, ev(c.args[1]), ev(c.args[2]))
        elif c.op == 'erf':
            v = _erf(ev(c.args[0]))
        elif c.op in _NP_UNARY:
            v = _NP_UNARY[c.op](ev(c.args[0]))
        elif c.op in _NP_BINARY:
            v = _NP_BINARY[c.op](ev(c.args[0]), ev(c.args[1]))
        else:
            raise NotImplementedError('host evaluation of {}'.format(c.op))
        cache[id(c)] = v
        return v

    return np.stack([ev(s.as_coef()) for s in cf.arr.reshape(-1)], axis=0)


def voxel_values(vox, pts: np.ndarray) -> np.ndarray:
    """Multilinear interpolation of VoxelCoefficient node data at points (clamped to the box)."""
    from scipy.interpolate import RegularGridInterpolator
    shape = vox.values.shape[::-1]
    axes = [np.linspace(vox.start[a], vox.end[a], shape[a]) for a in range(len(shape))]
    data = np.transpose(vox.values)                                  # indexed [x][y][z]
    f = RegularGridInterpolator(axes, data, method='linear', bounds_error=False, fill_value=None)
    q = np.stack([np.clip(pts[:, a], vox.start[a], vox.end[a]) for a in range(len(shape))], axis=1)
    return f(q)


def foreign_fields(cf, mesh) -> bool:
    leaves = coef_leaves([s.as_coef() for s in cf.arr.reshape(-1) if s.is_coef()], 'field')
    return any(lf.val[0].space.mesh.points is not mesh.points for lf in leaves)


def set_from_points(gf, source, definedon=None) -> None:
    """Local L2 projection + averaging (same algorithm as project.set_gridfunction) with a point-wise integrand:
    ``source`` is a VoxelCoefficient or a CoefficientFunction that may contain GridFunctions of other meshes."""
    from . import ngs
    if definedon is not None:
        raise NotImplementedError('host-side Set with definedon')
    root = gf._root_space
    mesh = root.mesh
    blocks = [root.blocks[b] for b in gf._blocks]
    if any(b.kind != 'scalar' for b in blocks):
        raise NotImplementedError('host-side Set on HDiv blocks')
    order = max(b.order for b in blocks)
    qp, w = cell_rule(mesh.cell_type, 2 * order)
    J = mesh.jacobians()
    x = (mesh.origins()[:, None, :] + np.einsum('eia,ka->eki', J, qp)).reshape(-1, mesh.dim)
    if isinstance(source, ngs.VoxelCoefficient):
        vals = voxel_values(source, x)[None, :]
    else:
        vals = evaluate_at_points(source, x)
    if vals.shape[0] != len(blocks):
        raise ValueError('Set: {} components for {} blocks'.format(vals.shape[0], len(blocks)))
    out_root = np.asarray(gf._root.vec.NumPy(), dtype=np.float64).copy()
    for k, (bi, b) in enumerate(zip(gf._blocks, blocks)):
        phi = b.basis.tabulate(qp)[:, 0, :]                              # (nq, nloc)
        M = phi.T @ (w[:, None] * phi)
        dual = (w[:, None] * phi) @ np.linalg.inv(M)                      # (nq, nloc)
        loc = vals[k].reshape(mesh.ne, len(w)) @ dual                     # (ne, nloc)
        lo = root.loc_offsets[bi]
        dofs = root.cell_dofs[:, lo:lo + b.nloc]
        acc = np.zeros(root.ndof)
        cnt = np.zeros(root.ndof)
        np.add.at(acc, dofs.ravel(), loc.ravel())
        np.add.at(cnt, dofs.ravel(), 1.0)
        sel = cnt > 0
        out_root[sel] = acc[sel] / cnt[sel]
    gf._root.vec.data = ngs.BaseVector(ngs.get_backend().from_numpy(out_root))
    gf._set_source = None
