"""Serialise lowered form programs (ir.FormProgram) together with everything they read — mesh, spaces, parameter
values, DOF vectors of the captured GridFunctions — into one ``.npz`` file, and rebuild them under any backend.

Used to carry the weak forms built by the *reference's own model classes* (lowered in the build container, where
/root/reference is mounted) to the GPU box, where the CUDA path assembles exactly those forms
(tests/golden/make_golden.py writes, tests/test_golden_programs.py replays)."""
from __future__ import annotations

import json
from typing import Dict, List

import numpy as np

from .ir import Bytecode, FormProgram, Integral
from .mesh import Mesh


def space_descriptor(fes) -> dict:
    if fes.components:
        return dict(compound=[space_descriptor(c) for c in fes.components], dgjumps=bool(fes.dgjumps))
    b = fes.blocks[0]
    return dict(family=fes.name, order=int(fes.order), dirichlet=b.dirichlet or '', dgjumps=bool(fes.dgjumps))


def build_space(ngs, mesh, d: dict):
    if 'compound' in d:
        return ngs.FESpace([build_space(ngs, mesh, c) for c in d['compound']], dgjumps=d['dgjumps'])
    return getattr(ngs, d['family'])(mesh, order=d['order'], dirichlet=d['dirichlet'], dgjumps=d['dgjumps'])


def dump(path: str, programs: Dict[str, FormProgram], extra: Dict[str, np.ndarray]) -> None:
    """programs: name -> FormProgram, all over the same mesh."""
    first = next(iter(programs.values()))
    mesh = first.fes.mesh
    arrays: Dict[str, np.ndarray] = dict(points=mesh.points, cells=mesh.cells, bnd_facets=mesh.facets[mesh.bnd_facets],
                                         bnd_region=mesh.bnd_region)
    meta = dict(dim=mesh.dim, cell_type=mesh.cell_type, bnd_names=list(mesh.bnd_names), programs={}, gfs=[])
    gfs: List = []

    def gf_index(gf):
        for i, g in enumerate(gfs):
            if g is gf:
                return i
        gfs.append(gf)
        return len(gfs) - 1

    for name, prog in programs.items():
        pm = dict(arity=prog.arity, space=space_descriptor(prog.fes), integrals=[])
        for k, integ in enumerate(prog.integrals):
            key = '{}__{}'.format(name, k)
            bc = integ.prog
            arrays[key + '_entries'] = integ.entries
            arrays[key + '_code'] = bc.code
            arrays[key + '_consts'] = bc.consts_arr
            arrays[key + '_params'] = np.array([float(p.Get()) for p in bc.params], dtype=np.float64)
            if integ.items is not None:
                arrays[key + '_items'] = integ.items
            pm['integrals'].append(dict(kind=integ.kind, deg=integ.deg, nreg=bc.nreg, nout=bc.nout,
                                        has_items=integ.items is not None,
                                        fields=[[gf_index(gf), int(blk), int(row), int(side)]
                                                for (gf, blk, row, side) in bc.fields]))
        meta['programs'][name] = pm
    for i, gf in enumerate(gfs):
        meta['gfs'].append(space_descriptor(gf.space))
        arrays['gf_{}'.format(i)] = np.asarray(gf.vec_numpy(), dtype=np.float64)
    for k, v in extra.items():
        arrays['x_' + k] = np.asarray(v)
    arrays['meta'] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(path, **arrays)


class _Const:
    def __init__(self, v):
        self.v = float(v)

    def Get(self):
        return self.v


def load(path: str, ngs):
    """Rebuild (mesh, {name: FormProgram}, extra arrays) under the backend currently installed in ``ngs``."""
    z = np.load(path)
    meta = json.loads(bytes(z['meta']).decode())
    mesh = ngs.Mesh(Mesh(meta['dim'], meta['cell_type'], z['points'], z['cells'], z['bnd_facets'], z['bnd_region'],
                         meta['bnd_names']))
    gfs = []
    for i, d in enumerate(meta['gfs']):
        gf = ngs.GridFunction(build_space(ngs, mesh, d))
        gf.vec.data = ngs.BaseVector(ngs.get_backend().from_numpy(z['gf_{}'.format(i)]))
        gfs.append(gf)
    spaces: Dict[str, object] = {}
    programs: Dict[str, FormProgram] = {}
    for name, pm in meta['programs'].items():
        skey = json.dumps(pm['space'], sort_keys=True)
        if skey not in spaces:
            spaces[skey] = build_space(ngs, mesh, pm['space'])
        fes = spaces[skey]
        integrals = []
        for k, im in enumerate(pm['integrals']):
            key = '{}__{}'.format(name, k)
            bc = Bytecode.__new__(Bytecode)
            bc.code = np.ascontiguousarray(z[key + '_code'], dtype=np.int32)
            bc.consts_arr = np.ascontiguousarray(z[key + '_consts'], dtype=np.float64)
            bc.consts = list(bc.consts_arr)
            bc.params = [_Const(v) for v in z[key + '_params']]
            bc.fields = [(gfs[g], blk, row, side) for g, blk, row, side in im['fields']]
            bc.nreg, bc.nout = im['nreg'], im['nout']
            items = z[key + '_items'] if im['has_items'] else None
            integrals.append(Integral(im['kind'], im['deg'], z[key + '_entries'], bc, items))
        programs[name] = FormProgram(fes, pm['arity'], integrals)
    extra = {k[2:]: z[k] for k in z.files if k.startswith('x_')}
    return mesh, programs, extra
