"""CUDA backend: ctypes binding of the C ABI (include/opencmp_b200.h) with PyTorch tensors as device buffers.

This is the product path behind opencmp_b200/ngs.py. It raises if the shared library or a CUDA device is missing —
there is deliberately no CPU fallback (the NumPy oracle lives under oracle/ and is test infrastructure only).

Per mesh / space the exported arrays are uploaded once (geometry, connectivity, DOF lists, CSR pattern, scatter
maps, reference tables); per ``Assemble()`` only the run-time parameter array and the field-vector pointers change.
"""
from __future__ import annotations

import ctypes as C
import os
import weakref
from typing import Dict, List, Optional

import numpy as np

from .ir import FormProgram, Integral
from .quadrature import cell_rule, facet_rule_in_cell, facet_ref_geometry

_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'lib', 'libopencmp_b200.so')
MAX_FVEC = 8


class CoefPlan(C.Structure):
    _fields_ = [('dim', C.c_int), ('kind', C.c_int), ('nq', C.c_int), ('nout', C.c_int), ('ninstr', C.c_int),
                ('nreg', C.c_int), ('nfgroups', C.c_int), ('nfslots', C.c_int),
                ('items', C.c_void_p), ('geo', C.c_void_p), ('facet_cells', C.c_void_p), ('facet_local', C.c_void_p),
                ('qpts', C.c_void_p), ('qw', C.c_void_p), ('fref', C.c_void_p), ('code', C.c_void_p),
                ('consts', C.c_void_p), ('params', C.c_void_p), ('fgroup', C.c_void_p), ('fgroup2', C.c_void_p),
                ('fslot', C.c_void_p), ('ftab', C.c_void_p),
                ('fvec', C.c_void_p * MAX_FVEC), ('fdof', C.c_void_p * MAX_FVEC)]


class ContractPlan(C.Structure):
    _fields_ = [('dim', C.c_int), ('kind', C.c_int), ('nq', C.c_int), ('nside', C.c_int), ('nblk', C.c_int),
                ('nloc', C.c_int), ('sbsz', C.c_int), ('zsz', C.c_int), ('nact', C.c_int), ('nslots', C.c_int),
                ('eb', C.c_int), ('ntiles', C.c_int), ('maxt', C.c_int), ('nzd', C.c_int), ('nent', C.c_int),
                ('nseg', C.c_int), ('qb', C.c_int),
                ('items', C.c_void_p), ('geo', C.c_void_p), ('facet_cells', C.c_void_p), ('facet_local', C.c_void_p),
                ('blk', C.c_void_p), ('tab', C.c_void_p), ('zdesc', C.c_void_p), ('ent', C.c_void_p),
                ('dofdesc', C.c_void_p), ('tiles', C.c_void_p), ('seg', C.c_void_p),
                ('cell2nnz', C.c_void_p), ('facet2nnz', C.c_void_p), ('cell_dofs', C.c_void_p), ('abuf', C.c_void_p)]


class System(C.Structure):
    _fields_ = [('nrows', C.c_int), ('rowptr', C.c_void_p), ('colidx', C.c_void_p), ('vals', C.c_void_p),
                ('freemask', C.c_void_p), ('pre_kind', C.c_int), ('dinv', C.c_void_p), ('npatch', C.c_int),
                ('bs', C.c_int), ('patch_dofs', C.c_void_p), ('inv_blocks', C.c_void_p), ('patch_weight', C.c_void_p),
                ('nlevels', C.c_int), ('levels', C.c_void_p), ('inv_rowptr', C.c_void_p), ('inv_colidx', C.c_void_p),
                ('inv_vals', C.c_void_p), ('owned', C.c_void_p), ('halo_fwd', C.c_int), ('halo_sum', C.c_int),
                ('inv_storage', C.c_int), ('vals32', C.c_void_p), ('apply_fn', C.c_void_p), ('apply_ctx', C.c_void_p),
                ('direct', C.c_void_p), ('patch_inc_ptr', C.c_void_p), ('patch_inc_idx', C.c_void_p),
                ('patch_ybuf', C.c_void_p), ('spmv_rows', C.c_void_p), ('n_spmv_rows', C.c_int),
                ('run_len', C.c_void_p), ('run_shift', C.c_int), ('run_nc', C.c_int), ('run_grouped', C.c_int),
                ('n_spmv_groups', C.c_int)]


class BandHandle(C.Structure):
    _fields_ = [('n', C.c_int), ('kl', C.c_int), ('ku', C.c_int), ('ubw', C.c_int), ('ab', C.c_void_p),
                ('ipiv', C.c_void_p), ('perm', C.c_void_p), ('rhs', C.c_void_p)]


# ocmp_apply_fn (include/opencmp_b200.h): y = A x of a matrix-free operator, called back by ocmp_krylov
APPLY_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p)


class MGLevel(C.Structure):
    _fields_ = [('sys', System), ('ncoarse', C.c_int), ('p_rowptr', C.c_void_p), ('p_colidx', C.c_void_p),
                ('p_vals', C.c_void_p), ('r_rowptr', C.c_void_p), ('r_colidx', C.c_void_p), ('r_vals', C.c_void_p),
                ('work', C.c_void_p), ('nu', C.c_int), ('omega', C.c_double), ('restrict_sum', C.c_int),
                ('handover', C.c_int)]


def load_library() -> C.CDLL:
    if not os.path.exists(_LIB_PATH):
        raise RuntimeError('opencmp_b200: {} is missing — build it with `python -c "import __graft_entry__ as g; '
                           'g.build()"`; there is no CPU fallback'.format(_LIB_PATH))
    lib = C.CDLL(_LIB_PATH)
    lib.ocmp_last_error.restype = C.c_char_p
    lib.ocmp_krylov_work_len.restype = C.c_longlong
    lib.ocmp_krylov_work_len.argtypes = [C.c_int, C.c_int, C.c_int]
    P = C.c_void_p
    lib.ocmp_eval_coefficients.argtypes = [C.POINTER(CoefPlan), C.c_int, C.c_int, P, P]
    lib.ocmp_contract_matrix.argtypes = [C.POINTER(ContractPlan), C.c_int, C.c_int, P, P, P]
    lib.ocmp_contract_vector.argtypes = [C.POINTER(ContractPlan), C.c_int, C.c_int, P, P, P]
    lib.ocmp_sum.argtypes = [P, C.c_longlong, P, P]
    lib.ocmp_gather_add.argtypes = [C.c_int, P, P, P, P, P, P]
    lib.ocmp_spmv.argtypes = [C.c_int, P, P, P, P, P, P]
    lib.ocmp_spmv_runs.argtypes = [C.c_int, P, P, C.c_int, C.c_int, P, P, P]
    lib.ocmp_spmv_compressed.argtypes = [C.c_int, P, P, P, P, C.c_int, C.c_int, C.c_int, P, P, P]
    lib.ocmp_dot.argtypes = [C.c_longlong, P, P, P, P]
    lib.ocmp_axpby.argtypes = [C.c_longlong, C.c_double, P, C.c_double, P, P]
    lib.ocmp_masked_assign.argtypes = [C.c_longlong, P, P, P, P, P]
    lib.ocmp_mdot.argtypes = [C.c_longlong, P, C.c_longlong, C.c_int, P, P, P]
    lib.ocmp_maxpy.argtypes = [C.c_longlong, P, C.c_longlong, C.c_int, P, P, P]
    lib.ocmp_jacobi_setup.argtypes = [C.c_int, P, P, P, P, P]
    lib.ocmp_asm_setup.argtypes = [C.c_int, C.c_int, P, P, P, P, P, P, P, P]
    lib.ocmp_patch_positions.argtypes = [C.c_int, C.c_int, P, P, P, P, P]
    lib.ocmp_asm_apply.argtypes = [C.c_int, C.c_int, P, P, P, P, P, P, P, C.c_longlong, P]
    lib.ocmp_asm_setup_f32.argtypes = lib.ocmp_asm_setup_bf16.argtypes = lib.ocmp_asm_setup.argtypes
    lib.ocmp_to_f32.argtypes = [C.c_longlong, P, P, P]
    lib.ocmp_asm_apply_f32.argtypes = lib.ocmp_asm_apply_bf16.argtypes = lib.ocmp_asm_apply.argtypes
    lib.ocmp_krylov.argtypes = [C.POINTER(System), C.c_int, P, P, C.c_double, C.c_int, C.c_int, C.c_double, P,
                                C.c_longlong, C.POINTER(C.c_int), C.POINTER(C.c_double), P]
    lib.ocmp_krylov_history.argtypes = [C.POINTER(C.c_double), C.c_int]
    lib.ocmp_comm_unique_id.argtypes = [C.c_char_p]
    lib.ocmp_comm_init.argtypes = [C.c_char_p, C.c_int, C.c_int]
    lib.ocmp_halo_plan.argtypes = [C.c_int, P, P, P, P, P, P, P]
    lib.ocmp_halo_run.argtypes = [C.c_int, P, C.c_int, P]
    lib.ocmp_allreduce_sum.argtypes = [P, C.c_int, P]
    lib.ocmp_band_len.restype = C.c_longlong
    lib.ocmp_band_len.argtypes = [C.c_int, C.c_int, C.c_int]
    lib.ocmp_band_fill.argtypes = [C.c_int, P, P, P, P, C.c_int, C.c_int, C.c_int, P, P]
    lib.ocmp_band_factor.argtypes = [C.c_int, C.c_int, C.c_int, P, P, P, P]
    lib.ocmp_band_solve.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, P, P, P, P]
    lib.ocmp_band_gather.argtypes = [C.c_int, P, P, P, P]
    lib.ocmp_band_scatter.argtypes = [C.c_int, P, P, P, C.c_int, P]
    lib.ocmp_dim_raytrace_2d.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int,
                                         C.c_int, C.c_int, P, P, P]
    lib.ocmp_dim_border.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, P, P, P]
    lib.ocmp_dim_edt.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, P, P, P, P, P]
    lib.ocmp_dim_phi.argtypes = [C.c_longlong, P, P, C.c_double, C.c_double, P, P]
    lib.ocmp_dim_rigid_motion.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double),
                                          C.POINTER(C.c_double), C.POINTER(C.c_double), P, P, P]
    lib.ocmp_profile_enable.argtypes = [C.c_int]
    lib.ocmp_profile_enable.restype = None
    lib.ocmp_profile_reset.restype = None
    lib.ocmp_profile_read.argtypes = [C.c_int, C.POINTER(C.c_longlong), C.POINTER(C.c_double)]
    lib.ocmp_launch_count.restype = C.c_longlong
    lib.ocmp_profile_bytes.restype = C.c_double
    lib.ocmp_profile_bytes.argtypes = [C.c_int]
    return lib


PROFILE_CATEGORIES = ['spmv', 'asm_apply', 'coef', 'contract_matrix', 'contract_vector', 'multi_dot', 'multi_axpy',
                      'vector', 'precond_setup', 'spmv_multigrid', 'halo_exchange']


def read_profile(lib) -> dict:
    out = {}
    for i, name in enumerate(PROFILE_CATEGORIES):
        cnt, ms = C.c_longlong(0), C.c_double(0.0)
        lib.ocmp_profile_read(i, C.byref(cnt), C.byref(ms))
        out[name] = dict(count=int(cnt.value), ms=float(ms.value), bytes=float(lib.ocmp_profile_bytes(i)))
    return out


EXPORTED = ['ocmp_mdot', 'ocmp_maxpy', 'ocmp_krylov_history', 'ocmp_comm_unique_id', 'ocmp_comm_init', 'ocmp_halo_plan', 'ocmp_halo_run', 'ocmp_allreduce_sum',
            'ocmp_patch_positions', 'ocmp_profile_bytes', 'ocmp_profile_enable', 'ocmp_profile_reset', 'ocmp_profile_read', 'ocmp_launch_count','ocmp_eval_coefficients', 'ocmp_contract_matrix', 'ocmp_contract_vector', 'ocmp_sum', 'ocmp_spmv',
            'ocmp_dot', 'ocmp_axpby', 'ocmp_masked_assign', 'ocmp_jacobi_setup', 'ocmp_asm_setup', 'ocmp_asm_apply',
            'ocmp_asm_setup_f32', 'ocmp_asm_apply_f32', 'ocmp_asm_setup_bf16', 'ocmp_asm_apply_bf16', 'ocmp_to_f32',
            'ocmp_krylov', 'ocmp_krylov_work_len', 'ocmp_last_error', 'ocmp_version',
            'ocmp_band_len', 'ocmp_band_fill', 'ocmp_band_factor', 'ocmp_band_solve', 'ocmp_band_gather',
            'ocmp_band_scatter', 'ocmp_spmv_runs', 'ocmp_spmv_compressed', 'ocmp_gather_add',
            'ocmp_dim_raytrace_2d', 'ocmp_dim_border', 'ocmp_dim_edt', 'ocmp_dim_phi', 'ocmp_dim_rigid_motion']


# storage type of the smoother's patch inverses (ocmp_system.inv_storage): arithmetic is FP64 in every case
STORAGE_ID = {'fp64': 0, 'fp32': 1, 'bf16': 2}
_STORAGE_ALIGN = {'fp64': 2, 'fp32': 4, 'bf16': 8}          # patch stride: every stored column 16-byte aligned


def deterministic_assembly() -> bool:
    """OCMP_DETERMINISTIC=0 selects the atomicAdd scatter; default: two-phase assembly (item-local tiles + ordered gather),
    bit-reproducible between runs and between the ranks of an element-partitioned run."""
    return os.environ.get('OCMP_DETERMINISTIC', '1') != '0'


def patch_storage() -> str:
    """OCMP_PATCH_STORAGE = fp64 (default) | fp32 | bf16; OCMP_PATCH_FP32=1 is a synonym of fp32."""
    kind = os.environ.get('OCMP_PATCH_STORAGE', 'fp32' if os.environ.get('OCMP_PATCH_FP32', '0') == '1' else 'fp64')
    if kind not in STORAGE_ID:
        raise ValueError('OCMP_PATCH_STORAGE must be one of {}'.format(sorted(STORAGE_ID)))
    return kind


def patch_incidence(dofs: np.ndarray, ndof: int):
    """Incidence list of a patch table (npatch x bs, -1 padded): for every dof the flat positions p * bs + i of its
    patch entries in ascending order — the fixed summation order of the deterministic smoother gather
    (ocmp_asm_apply). Returns (inc_ptr (ndof + 1) int32, inc_idx int32)."""
    flat = dofs.ravel()
    pos = np.nonzero(flat >= 0)[0]
    order = np.argsort(flat[pos], kind='stable')          # stable: positions stay ascending within a dof
    inc_idx = pos[order].astype(np.int32)
    counts = np.bincount(flat[pos], minlength=ndof)
    inc_ptr = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    return inc_ptr, inc_idx


def contract_smem_bytes(xp, eb: int, gs: int) -> int:
    """Mirror of contract_smem() in csrc/ocmp_assembly.cu."""
    dbl = eb * xp.qb * (xp.zsz + xp.nside * xp.sbsz + xp.nslots) + eb * xp.nside * gs
    dbl += dbl % 2
    ints = 4 * eb + 4 * xp.nside * xp.nloc + 4 * xp.nzd + 2 * xp.nent + 2 * xp.nseg
    return 8 * dbl + 4 * ints + 16


def contract_tables(entries, decode, nloc, kind, nrows, tab_off, sb_off, sbsz, loc_off, nside):
    """Descriptor tables of k_contract for one bilinear integral. ``entries``: (test row, trial row, D slot) of the
    lowered integrand; ``decode(row) -> (side, block, physical row of the block)``. Per item and quadrature point the
    kernel forms Z_g[j] = sum_k D[slot_k] B_trial[row_k][j] for every group g = (test side/block/row, trial side/block)
    and then updates 4 x 4 register tiles A[i0:i0+4, j0:j0+4] += B_test[row_g][i] Z_g[j] over the groups ("segments")
    of the tile's block pair. Shared-memory rows are padded to multiples of 4 so that tile edges are 16-byte loads."""
    pad4 = lambda v: (v + 3) // 4 * 4
    nblk = len(nloc)
    segs: Dict[tuple, list] = {}
    for tr, ur, slot in entries:
        st, bt, rt = decode(tr)
        su, bu, ru = decode(ur)
        segs.setdefault((st, bt, rt, su, bu), []).append((slot, ru))
    # per (side, local dof): how to build its physical table column
    dofdesc = []
    for s_ in range(nside):
        for b in range(nblk):
            for il in range(nloc[b]):
                dofdesc.append([kind[b] | (nrows[b] << 8) | (pad4(nloc[b]) << 16), nloc[b], tab_off[b] + il,
                                sb_off[b] + il])
    # Z rows: one (padded) segment per (test row, trial side-block)
    zdesc, ent, seg_z = [], [], {}
    zoff = z_fma = 0
    for key in sorted(segs):
        st, bt, rt, su, bu = key
        k0 = len(ent)
        nl = nloc[bu]
        ent += [[slot, ru * pad4(nl)] for slot, ru in segs[key]]
        k1 = len(ent)
        seg_z[key] = zoff
        for j in range(nl):
            zdesc.append([k0, k1, su * sbsz + sb_off[bu] + j, zoff + j])
        z_fma += (k1 - k0) * nl
        zoff += pad4(nl)
    # active block pairs -> segments and 4 x 4 tiles
    pairs: Dict[tuple, list] = {}
    for key in sorted(segs):
        st, bt, rt, su, bu = key
        pairs.setdefault((st, bt, su, bu), []).append((rt, seg_z[key]))
    seg_arr, tiles = [], []
    nact = a_fma = 0
    for (st, bt, su, bu), lst in sorted(pairs.items()):
        ni, nj = nloc[bt], nloc[bu]
        s0 = len(seg_arr)
        seg_arr += [[rt * pad4(ni), z] for rt, z in lst]
        for i0 in range(0, ni, 4):
            for j0 in range(0, nj, 4):
                tiles.append([st * sbsz + sb_off[bt] + i0, j0, s0, len(lst), st | (su << 1), loc_off[bt] + i0,
                              loc_off[bu] + j0, min(4, ni - i0) | (min(4, nj - j0) << 8)])
        nact += ni * nj
        a_fma += ni * nj * len(lst)
    i32 = lambda a, w: np.asarray(a, dtype=np.int32).reshape(-1, w)
    return dict(zdesc=i32(zdesc, 4), ent=i32(ent, 2), seg=i32(seg_arr, 2), dofdesc=i32(dofdesc, 4),
                tiles=i32(tiles, 8), zsz=zoff, nact=nact, z_fma=z_fma, a_fma=a_fma)


def component_runs(fes):
    """(shift, nc) when the space starts with nc = 2 or 3 scalar blocks of equal size numbered one after the other
    (the components of a VectorH1 block): the CSR rows then hold nc shifted copies of the same scalar columns and the
    SpMV kernels read only the first copy's indices (ocmp_spmv_runs verifies it row by row). None otherwise."""
    blocks, offs = fes.blocks, list(fes.block_offsets)
    if len(blocks) < 2 or blocks[0].kind != 'scalar' or offs[0] != 0:
        return None
    size = blocks[0].ndof
    nc = 1
    while nc < min(3, len(blocks)) and blocks[nc].kind == 'scalar' and blocks[nc].ndof == size and \
            blocks[nc].nloc == blocks[0].nloc and offs[nc] == nc * size:
        nc += 1
    if nc < 2 or size <= 0:
        return None
    return int(size), int(nc)


def _ptr(t) -> Optional[int]:
    return None if t is None else t.data_ptr()


class _Precond:
    def __init__(self, kind, **kw):
        self.kind = kind
        self.__dict__.update(kw)


class CudaBackend:
    name = 'cuda'

    def __init__(self, device: Optional[int] = None):
        import torch
        self.torch = torch
        self.lib = load_library()
        if not torch.cuda.is_available():
            raise RuntimeError('opencmp_b200: no CUDA device visible — the backend has no CPU fallback')
        self.device = torch.device('cuda', torch.cuda.current_device() if device is None else device)
        self._mesh_cache: Dict[int, dict] = {}
        self._space_cache: Dict[int, dict] = {}
        self._plan_cache = weakref.WeakKeyDictionary()     # plans die with their FormProgram (Integrate builds one per call)
        self._dbuf = None
        self._abuf = None
        self._scal = torch.zeros(64, dtype=torch.float64, device=self.device)
        self.launches = 0
        self.last_iters = 0
        self.last_resid = 0.0
        self.chunk_bytes = 48 << 20

    # ---- helpers -----------------------------------------------------------------------------------------------
    def comm_init(self) -> bool:
        """Create the library's own NCCL communicator (id broadcast through torch.distributed). Idempotent."""
        import torch.distributed as dist
        if getattr(self, '_comm_ready', False):
            return True
        if not (dist.is_initialized() and dist.get_world_size() > 1):
            return False
        buf = C.create_string_buffer(128)
        if dist.get_rank() == 0:
            self._ck(self.lib.ocmp_comm_unique_id(buf))
        t = self.torch.frombuffer(bytearray(buf.raw), dtype=self.torch.uint8).to(self.device)
        dist.broadcast(t, src=0)
        raw = bytes(t.cpu().numpy().tobytes())
        self._ck(self.lib.ocmp_comm_init(raw, dist.get_world_size(), dist.get_rank()))
        self._comm_ready = True
        return True

    def halo_plan(self, peers, send_lists, recv_lists):
        """Build a C-side halo plan from per-neighbour local index arrays; returns (handle, keep-alive tensors)."""
        t = self.torch
        cat = lambda ls: np.concatenate(ls).astype(np.int32) if ls else np.zeros(0, np.int32)
        si, ri = self._up(cat(send_lists)), self._up(cat(recv_lists))
        sb = t.empty(max(1, si.numel()), dtype=t.float64, device=self.device)
        rb = t.empty(max(1, ri.numel()), dtype=t.float64, device=self.device)
        arr = lambda v: (C.c_int * len(v))(*[int(a) for a in v])
        h = self.lib.ocmp_halo_plan(len(peers), arr(peers), arr([len(a) for a in send_lists]),
                                    arr([len(a) for a in recv_lists]), si.data_ptr(), ri.data_ptr(), sb.data_ptr(),
                                    rb.data_ptr())
        if h < 0:
            self._ck(h)
        return h, (si, ri, sb, rb)

    def halo_run(self, handle, x, add: bool):
        self._ck(self.lib.ocmp_halo_run(handle, x.data_ptr(), 1 if add else 0, self._stream()))

    def _stream(self):
        return self.torch.cuda.current_stream(self.device).cuda_stream

    def _ck(self, rc):
        if rc != 0:
            raise RuntimeError('opencmp_b200 C ABI error {}: {}'.format(rc, self.lib.ocmp_last_error().decode()))

    def _up(self, a, dtype=None):
        t = self.torch.from_numpy(np.ascontiguousarray(a if dtype is None else np.asarray(a, dtype=dtype)))
        return t.to(self.device)

    def zeros(self, n):
        return self.torch.zeros(int(n), dtype=self.torch.float64, device=self.device)

    def from_numpy(self, a):
        return self._up(np.asarray(a, dtype=np.float64))

    def to_numpy(self, a):
        return a.detach().cpu().numpy()

    def numpy_view(self, a):
        return a.detach().cpu().numpy()

    def copy_into(self, dst, src):
        if isinstance(src, np.ndarray):        # e.g. the reference's host-side mixers hand back NumPy arrays
            src = self.torch.from_numpy(np.ascontiguousarray(src, dtype=np.float64))
        dst.copy_(src)

    def mdot(self, V, w):
        """Host array of <V_j, w> for the rows of the 2-D device array V (one batched launch + one read-back)."""
        k, n = V.shape
        if V.stride(1) != 1 or not w.is_contiguous():
            raise ValueError('mdot needs unit-stride rows and a contiguous vector')
        out = self._scal if k <= self._scal.numel() else self.zeros(k)     # e.g. Anderson mixing with a long history
        self._ck(self.lib.ocmp_mdot(n, V.data_ptr(), V.stride(0), k, w.data_ptr(), out.data_ptr(), self._stream()))
        self.launches += 1
        return out[:k].cpu().numpy()

    def maxpy(self, V, coef, w):
        """w += sum_j coef[j] V_j (coef: host array)."""
        k, n = V.shape
        c = self._up(np.asarray(coef, dtype=np.float64))
        self._ck(self.lib.ocmp_maxpy(n, V.data_ptr(), V.stride(0), k, c.data_ptr(), w.data_ptr(), self._stream()))
        self.launches += 1

    def dot(self, a, b):
        if not (a.is_contiguous() and b.is_contiguous()):      # BaseVector slices with a step
            a, b = a.contiguous(), b.contiguous()
        self._ck(self.lib.ocmp_dot(a.numel(), a.data_ptr(), b.data_ptr(), self._scal.data_ptr(), self._stream()))
        self.launches += 1
        return float(self._scal[0].item())

    def masked_assign(self, dst, src, inv, mask):
        self._ck(self.lib.ocmp_masked_assign(dst.numel(), dst.data_ptr(), src.data_ptr(), inv.data_ptr(),
                                             mask.data_ptr(), self._stream()))
        self.launches += 1

    # ---- exported arrays ---------------------------------------------------------------------------------------
    def mesh_data(self, mesh) -> dict:
        key = id(mesh)
        d = self._mesh_cache.get(key)
        if d is None or d['ne'] != mesh.ne:
            J = mesh.jacobians()
            geo = np.concatenate([mesh.origins(), J.reshape(mesh.ne, -1), np.linalg.inv(J).reshape(mesh.ne, -1),
                                  np.linalg.det(J)[:, None]], axis=1)
            tang, nrm = facet_ref_geometry(mesh.cell_type)
            fref = np.concatenate([nrm, tang.reshape(tang.shape[0], -1)], axis=1)
            d = dict(ne=mesh.ne, geo=self._up(geo), facet_cells=self._up(mesh.facet_cells),
                     facet_local=self._up(mesh.facet_local), fref=self._up(fref), keep=mesh)
            self._mesh_cache[key] = d
        return d

    def space_data(self, fes) -> dict:
        key = id(fes)
        d = self._space_cache.get(key)
        if d is None or d['ndof'] != fes.ndof:
            d = dict(ndof=fes.ndof, cell_dofs=self._up(fes.cell_dofs), keep=fes)
            self._space_cache[key] = d
        return d

    def pattern_data(self, fes) -> dict:
        d = self.space_data(fes)
        if 'rowptr' not in d:
            pat = fes.pattern()
            d.update(rowptr=self._up(pat.rowptr.astype(np.int32)), colidx=self._up(pat.colidx),
                     cell2nnz=self._up(pat.cell2nnz), facet2nnz=self._up(pat.facet2nnz), diag=self._up(pat.diag),
                     nnz=pat.nnz, runs=None)
            comp = component_runs(fes)
            if comp is not None and os.environ.get('OCMP_SPMV_RUNS', '1') != '0':
                shift, nc = comp
                runlen = self.torch.zeros(fes.ndof, dtype=self.torch.int32, device=self.device)
                bad = self.torch.zeros(1, dtype=self.torch.int32, device=self.device)
                self._ck(self.lib.ocmp_spmv_runs(fes.ndof, d['rowptr'].data_ptr(), d['colidx'].data_ptr(), shift, nc,
                                                 runlen.data_ptr(), bad.data_ptr(), self._stream()))
                grouped = int(bad.item()) == 0 and os.environ.get('OCMP_SPMV_GROUPED', '1') != '0'
                d['runs'] = (runlen, shift, nc, grouped)
        return d

    # ---- plan construction -------------------------------------------------------------------------------------
    def _tables(self, basis, kind: str, deg: int) -> np.ndarray:
        return (basis.tabulate_cell(deg) if kind == 'cell' else basis.tabulate_facets(deg)).reshape(-1)

    def _build_plans(self, program: FormProgram, integ: Integral) -> dict:
        fes = program.fes
        mesh = fes.mesh
        dim = mesh.dim
        md = self.mesh_data(mesh)
        kind_id = {'cell': 0, 'ifacet': 1, 'bfacet': 2}[integ.kind]
        prog = integ.prog
        if integ.kind == 'cell':
            qp, qw = cell_rule(mesh.cell_type, integ.deg)
        else:
            qp, qw = facet_rule_in_cell(mesh.cell_type, integ.deg)
        nq = len(qw)
        keep: List = []

        def up(a, dtype):
            t = self._up(np.asarray(a, dtype=dtype))
            keep.append(t)
            return t

        nitems = mesh.ne if integ.items is None else len(integ.items)
        items = None if integ.items is None else up(integ.items, np.int32)
        cp = CoefPlan()
        cp.dim, cp.kind, cp.nq, cp.nout = dim, kind_id, nq, prog.nout
        cp.ninstr, cp.nreg = prog.code.shape[0], prog.nreg
        cp.items = _ptr(items)
        cp.geo, cp.facet_cells, cp.facet_local = md['geo'].data_ptr(), md['facet_cells'].data_ptr(), \
            md['facet_local'].data_ptr()
        cp.qpts, cp.qw, cp.fref = up(qp, np.float64).data_ptr(), up(qw, np.float64).data_ptr(), md['fref'].data_ptr()
        cp.code = up(prog.code, np.int32).data_ptr()
        cp.consts = up(prog.consts_arr, np.float64).data_ptr()
        params = up(np.zeros(max(1, len(prog.params))), np.float64)
        cp.params = params.data_ptr()
        # field groups
        roots: List = []
        spaces: List = []
        groups: Dict[tuple, int] = {}
        fgroup, fgroup2, fslot, ftab = [], [], [], []
        toff = 0
        for (gf, blk, row, side) in prog.fields:
            rs = gf.space
            if gf not in roots:
                roots.append(gf)
            if rs not in spaces:
                spaces.append(rs)
            gkey = (id(gf), blk, side)
            if gkey not in groups:
                groups[gkey] = len(fgroup)
                b = rs.blocks[blk]
                tab = self._tables(b.basis, 'cell' if integ.kind == 'cell' else 'facet', integ.deg)
                fgroup.append([roots.index(gf), spaces.index(rs), side, 0 if b.kind == 'scalar' else 1, b.nloc,
                               b.basis.nrows, toff, rs.loc_offsets[blk]])
                fgroup2.append([rs.nloc, 0])
                ftab.append(tab)
                toff += tab.size
            fslot.append([groups[gkey], row])
        if len(roots) > MAX_FVEC:
            raise ValueError('an integral reads more than {} distinct DOF vectors'.format(MAX_FVEC))
        cp.nfgroups, cp.nfslots = len(fgroup), len(fslot)
        cp.fgroup = up(np.array(fgroup if fgroup else [[0] * 8]), np.int32).data_ptr()
        cp.fgroup2 = up(np.array(fgroup2 if fgroup2 else [[0, 0]]), np.int32).data_ptr()
        cp.fslot = up(np.array(fslot if fslot else [[0, 0]]), np.int32).data_ptr()
        cp.ftab = up(np.concatenate(ftab) if ftab else np.zeros(1), np.float64).data_ptr()
        for i, rs in enumerate(spaces):
            cp.fdof[i] = self.space_data(rs)['cell_dofs'].data_ptr()
        plan = dict(coef=cp, params=params, roots=roots, nitems=nitems, nq=nq, keep=keep, contract=None)
        if program.arity == 0:
            return plan
        # ---- contraction plan ------------------------------------------------------------------------------
        nside = 2 if integ.kind == 'ifacet' else 1
        blocks = fes.blocks
        nblk = len(blocks)
        tkind = 'cell' if integ.kind == 'cell' else 'facet'
        pad4 = lambda v: (v + 3) // 4 * 4
        blk_tab, tabs = [], []
        toff = sboff = 0
        sb_off = []
        for b, lo in zip(blocks, fes.loc_offsets):
            tab = self._tables(b.basis, tkind, integ.deg)
            blk_tab.append([0 if b.kind == 'scalar' else 1, b.nloc, b.basis.nrows, toff, lo, sboff])
            sb_off.append(sboff)
            tabs.append(tab)
            toff += tab.size
            sboff += b.basis.nrows * pad4(b.nloc)       # shared-memory rows padded: 16-byte loads of 4 x 4 tile edges
        sbsz = sboff
        nrows_tot = fes.nrows
        ro = fes.row_offsets

        def decode(row):
            side, rr = divmod(int(row), nrows_tot)
            b = max(i for i in range(nblk) if ro[i] <= rr)
            return side, b, rr - ro[b]

        xp = ContractPlan()
        xp.dim, xp.kind, xp.nq, xp.nside, xp.nblk, xp.nloc = dim, kind_id, nq, nside, nblk, fes.nloc
        xp.sbsz, xp.nslots = sbsz, prog.nout
        xp.items = _ptr(items)
        xp.geo, xp.facet_cells, xp.facet_local = cp.geo, cp.facet_cells, cp.facet_local
        xp.blk = up(np.array(blk_tab), np.int32).data_ptr()
        xp.tab = up(np.concatenate(tabs), np.float64).data_ptr()
        sd = self.space_data(fes)
        xp.cell_dofs = sd['cell_dofs'].data_ptr()
        if program.arity == 1:
            ent = []
            for tr, _, slot in integ.entries:
                s, b, r = decode(tr)
                ent.append([slot, ((s * nblk + b) << 8) | r])
            xp.ent = up(np.array(ent), np.int32).data_ptr()
            xp.nent = len(ent)
            xp.eb = 1
            plan['contract'] = xp
            return plan
        pd = self.pattern_data(fes)
        xp.cell2nnz, xp.facet2nnz = pd['cell2nnz'].data_ptr(), pd['facet2nnz'].data_ptr()
        tb = contract_tables(integ.entries, decode, [b.nloc for b in blocks],
                             [0 if b.kind == 'scalar' else 1 for b in blocks], [b.basis.nrows for b in blocks],
                             [t[3] for t in blk_tab], sb_off, sbsz, list(fes.loc_offsets), nside)
        xp.zsz, xp.nact, xp.ntiles = tb['zsz'], tb['nact'], len(tb['tiles'])
        xp.nzd, xp.nent, xp.nseg = len(tb['zdesc']), len(tb['ent']), len(tb['seg'])
        xp.zdesc = up(tb['zdesc'], np.int32).data_ptr()
        xp.ent = up(tb['ent'], np.int32).data_ptr()
        xp.seg = up(tb['seg'], np.int32).data_ptr()
        xp.dofdesc = up(tb['dofdesc'], np.int32).data_ptr()
        xp.tiles = up(tb['tiles'], np.int32).data_ptr()
        gs = dim + 2 * dim * dim + 1
        xp.qb = min(4, nq)                              # quadrature points staged per barrier round
        # items per 256-thread CTA: as many as leave every thread at most one 4 x 4 tile (16 FP64 accumulators in
        # registers); items with more than 256 tiles take 2 or 4 tiles per thread
        eb_pick, maxt = 1, 1
        for eb in (16, 8, 4, 2, 1):
            if xp.ntiles <= 256 // eb and contract_smem_bytes(xp, eb, gs) <= 96 * 1024:
                eb_pick = eb
                break
        else:
            maxt = 2 if xp.ntiles <= 512 else 4
            if xp.ntiles > 1024:
                raise ValueError('a local matrix of {} 4 x 4 tiles does not fit k_contract (limit 1024)'
                                 .format(xp.ntiles))
        xp.eb, xp.maxt = eb_pick, maxt
        plan['contract'] = xp
        plan['tables'] = tb
        plan['flops_per_item'] = 2.0 * nq * (tb['z_fma'] + tb['a_fma'])
        return plan

    def _plans(self, program: FormProgram) -> list:
        hit = self._plan_cache.get(program)
        if hit is None:
            hit = [self._build_plans(program, integ) for integ in program.integrals]
            self._plan_cache[program] = hit
        return hit

    def _run(self, program: FormProgram, out, mode: str):
        torch = self.torch
        st = self._stream()
        for integ, plan in zip(program.integrals, self._plans(program)):
            cp = plan['coef']
            vals = program.param_values(integ)
            plan['params'].copy_(torch.from_numpy(vals))
            for i, gf in enumerate(plan['roots']):
                cp.fvec[i] = gf.vec.a.data_ptr()
            nq, nout, nitems = plan['nq'], cp.nout, plan['nitems']
            chunk = max(256, min(nitems, self.chunk_bytes // (8 * nq * nout)))
            need = chunk * nq * nout
            if self._dbuf is None or self._dbuf.numel() < need:
                self._dbuf = torch.empty(need, dtype=torch.float64, device=self.device)
            det = mode in ('matrix', 'vector') and deterministic_assembly()
            for i0 in range(0, nitems, chunk):
                n = min(chunk, nitems - i0)
                self._ck(self.lib.ocmp_eval_coefficients(C.byref(cp), i0, n, self._dbuf.data_ptr(), st))
                self.launches += 1
                if mode == 'sum':
                    self._ck(self.lib.ocmp_sum(self._dbuf.data_ptr(), n * nq * nout, out.data_ptr(), st))
                    self.launches += 1
                    continue
                xp = plan['contract']
                fn = self.lib.ocmp_contract_matrix if mode == 'matrix' else self.lib.ocmp_contract_vector
                if not det:
                    xp.abuf = None
                    self._ck(fn(C.byref(xp), i0, n, self._dbuf.data_ptr(), out.data_ptr(), st))
                    self.launches += 1
                    continue
                # deterministic: local tiles / vectors into the item-local buffer, then the ordered gather
                per_item = xp.ntiles * 16 if mode == 'matrix' else xp.nside * xp.nloc
                need_a = n * per_item
                if self._abuf is None or self._abuf.numel() < need_a:
                    self._abuf = torch.empty(max(need_a, 1 << 20), dtype=torch.float64, device=self.device)
                if mode == 'vector':
                    self._abuf[:need_a].zero_()               # k_lin skips sides without a cell / untested blocks
                lists = plan.setdefault('gather_' + mode, {})
                if i0 not in lists:
                    lists[i0] = self._gather_lists(program, integ, plan, i0, n, mode)
                seg_ptr, seg_tgt, order = lists[i0]
                xp.abuf = self._abuf.data_ptr()
                self._ck(fn(C.byref(xp), i0, n, self._dbuf.data_ptr(), out.data_ptr(), st))
                self._ck(self.lib.ocmp_gather_add(seg_tgt.numel(), seg_ptr.data_ptr(), seg_tgt.data_ptr(),
                                                  order.data_ptr(), self._abuf.data_ptr(), out.data_ptr(), st))
                xp.abuf = None
                self.launches += 2

    def _gather_lists(self, program, integ, plan, i0: int, n: int, mode: str):
        """Contributor lists of one chunk of one integral for ocmp_gather_add: the flat index of every local tile entry
        (matrices) / local vector entry in the item-local buffer, stably sorted by its target (CSR position / dof).
        Returns (seg_ptr, seg_tgt, order) as int32 device tensors. Built once; a pure function of the mesh, the space
        and the plan tables — identical on every rank that holds the same cells."""
        t = self.torch
        fes = program.fes
        mesh = fes.mesh
        md = self.mesh_data(mesh)
        dev = self.device
        xp = plan['contract']
        ids = t.arange(i0, i0 + n, device=dev, dtype=t.int64)             # item index (facet2nnz is indexed by it)
        gid = ids if integ.items is None else self._up(np.asarray(integ.items[i0:i0 + n], dtype=np.int64))
        if integ.kind == 'cell':
            cells = t.stack([gid, gid], dim=1)
        else:
            cells = md['facet_cells'][gid].to(t.int64)                    # (n, 2), -1 where there is no second cell
        nloc = fes.nloc
        if mode == 'vector':
            cd = self.space_data(fes)['cell_dofs'].to(t.int64)            # (ncells, nloc)
            nside = xp.nside
            c = cells[:, :nside]                                          # (n, nside)
            tgt = cd[c.clamp(min=0)]                                      # (n, nside, nloc)
            tgt = t.where(c[:, :, None] >= 0, tgt, t.full_like(tgt, -1)).reshape(-1)
        else:
            pd = self.pattern_data(fes)
            n2 = nloc * nloc
            tiles = plan['tables']['tiles'].astype(np.int64)
            a = np.arange(4)
            ii = tiles[:, 5][:, None, None] + a[None, :, None]            # (ntiles, 4, 4) local row
            jj = tiles[:, 6][:, None, None] + a[None, None, :]
            ok = (a[None, :, None] < (tiles[:, 7] & 0xff)[:, None, None]) & \
                 (a[None, None, :] < (tiles[:, 7] >> 8)[:, None, None])
            ij = self._up((ii * nloc + jj).reshape(-1))                   # (ntiles * 16)
            okd = self._up(ok.reshape(-1))
            st_ = self._up(np.repeat(tiles[:, 4] & 1, 16))
            su_ = self._up(np.repeat((tiles[:, 4] >> 1) & 1, 16))
            c_st = t.where(st_[None, :] == 0, cells[:, 0:1], cells[:, 1:2])           # (n, ntiles*16)
            same = (st_ == su_)[None, :]
            idx_same = c_st.clamp(min=0) * n2 + ij[None, :]
            idx_cross = (ids[:, None] * 2 + st_[None, :]) * n2 + ij[None, :]
            c2n = pd['cell2nnz'].reshape(-1)
            f2n = pd['facet2nnz'].reshape(-1)
            if f2n.numel() == 0:
                f2n = c2n
            tgt = t.where(same, c2n[idx_same.clamp(max=c2n.numel() - 1)].to(t.int64),
                          f2n[idx_cross.clamp(max=f2n.numel() - 1)].to(t.int64))
            tgt = t.where(okd[None, :] & (c_st >= 0), tgt, t.full_like(tgt, -1)).reshape(-1)
        order = t.sort(tgt, stable=True).indices
        ninv = int((tgt < 0).sum().item())
        order = order[ninv:]
        sorted_tgt = tgt[order]
        seg_tgt, counts = t.unique_consecutive(sorted_tgt, return_counts=True)
        seg_ptr = t.zeros(seg_tgt.numel() + 1, dtype=t.int64, device=dev)
        if seg_tgt.numel():
            seg_ptr[1:] = t.cumsum(counts, 0)
        return seg_ptr.to(t.int32), seg_tgt.to(t.int32), order.to(t.int32)

    # ---- assembly ----------------------------------------------------------------------------------------------
    def assemble_matrix(self, program, mat):
        mat.values.zero_()
        self._run(program, mat.values, 'matrix')

    def matrix_flops(self, program) -> float:
        """FP64 flops of one assembly of ``program`` in the contraction kernel (for the assembly roofline)."""
        return float(sum(p['nitems'] * p.get('flops_per_item', 0.0) for p in self._plans(program)))

    def assemble_vector(self, program, out):
        out.zero_()
        self._run(program, out, 'vector')

    def integrate(self, program):
        self._scal[1].zero_()
        self._run(program, self._scal[1:2], 'sum')
        return float(self._scal[1].item())

    # ---- linear algebra ----------------------------------------------------------------------------------------
    def spmv(self, mat, x, out):
        pd = self.pattern_data(mat.space)
        self._ck(self.lib.ocmp_spmv(mat.height, pd['rowptr'].data_ptr(), pd['colidx'].data_ptr(),
                                    mat.values.data_ptr(), x.data_ptr(), out.data_ptr(), self._stream()))
        self.launches += 1

    def _mask(self, fes_or_n, free):
        """Device copy (1.0 free / 0.0 constrained) of a FreeDofs mask. The callers hand over a fresh host array on
        every solve (``fes.FreeDofs()``, reference base_model.py:906), so the upload is cached on the packed bits —
        compared exactly, a handful of entries."""
        if free is None:
            return None
        arr = np.asarray(free.a if hasattr(free, 'a') else free, dtype=bool)
        packed = np.packbits(arr).tobytes()
        cache = self.__dict__.setdefault('_mask_cache', [])
        for i, (n, bits, dev) in enumerate(cache):
            if n == arr.size and bits == packed:
                cache.append(cache.pop(i))                 # most recently used last
                return dev
        dev = self._up(arr.astype(np.float64))
        cache.append((arr.size, packed, dev))
        del cache[:-8]
        return dev

    def precond_setup(self, mat, kind, free, form=None, state=None, mask=None):
        pd = self.pattern_data(mat.space)
        fm = mask if mask is not None else self._mask(mat.space, free)
        st = self._stream()
        if kind == 'multigrid':
            from .multigrid import MultigridState
            if state is None or getattr(state, 'kind', 0) != 3:
                state = MultigridState(self, form, nu=int(os.environ.get('OCMP_MG_NU', '1')),
                                       omega=float(os.environ.get('OCMP_MG_OMEGA', '0.7')))
            return state.update(mat)
        if kind in ('local', 'jacobi'):
            dinv = self.zeros(mat.height)
            self._ck(self.lib.ocmp_jacobi_setup(mat.height, pd['diag'].data_ptr(), mat.values.data_ptr(), _ptr(fm),
                                                dinv.data_ptr(), st))
            self.launches += 1
            return _Precond(1, dinv=dinv, fm=fm)
        if kind in ('asm', 'asm_cell', 'block'):
            # overlapping additive Schwarz: vertex-star patches (all dofs of the cells around a vertex) by default,
            # cell patches for kind 'asm_cell'; averaged by the patch multiplicity of every dof
            fes = mat.space
            pt = self._patches(fes, 'cell' if kind == 'asm_cell' else os.environ.get('OCMP_PATCH', 'vertex'))
            # the patch TOPOLOGY is shared per space; the stored inverses belong to this preconditioner state alone
            # (several preconditioners on one space — adaptive_two_step.py:50-59 — must not overwrite each other)
            old = getattr(state, 'inv', None) if getattr(state, 'kind', 0) == 2 else None
            if old is not None and (state.pdofs is not pt['dofs'] or state.storage != pt['storage']):
                old = None
            inv = self._invert_patches(pt, pd, mat, fm, old)
            self.launches += 1
            return _Precond(2, inv=inv, npatch=pt['npatch'], bs=pt['bs'], pdofs=pt['dofs'], fm=fm,
                            wgt=pt['wgt'], storage=pt['storage'], inc_ptr=pt['inc_ptr'], inc_idx=pt['inc_idx'],
                            ybuf=pt['ybuf'])
        if kind == 'direct':
            # NGSolve's 'direct' preconditioner is a sparse factorisation of the assembled matrix
            # (reference base_model.py:365-383): the band LU of direct.py, applied as an exact inverse
            fact = self.factorize(mat, free) if getattr(state, 'kind', 0) != 5 else state.fact
            if getattr(state, 'kind', 0) == 5:
                fact.Update()
            return _Precond(5, fact=fact, fm=fm)
        if kind in ('h1amg', 'bddc'):
            raise NotImplementedError(
                "opencmp_b200: preconditioner type '{}' (reference base_model.py:365-383) is not implemented; "
                "available: local, direct, multigrid, and the additive-Schwarz types asm / asm_cell".format(kind))
        raise NotImplementedError('preconditioner type {}'.format(kind))

    def _invert_patches(self, pt, pd, mat, fm, inv=None):
        """Patch inverses of the patch table ``pt`` for the current matrix values, written into ``inv`` (allocated when
        None) and returned. With ``pt['storage']`` = fp32 / bf16 (OCMP_PATCH_STORAGE) they are stored in that type —
        inverted and applied in FP64 arithmetic."""
        t = self.torch
        npatch, bs = pt['npatch'], pt['bs']
        st = self._stream()
        if inv is None:
            dtype = {'fp64': t.float64, 'fp32': t.float32, 'bf16': t.bfloat16}[pt['storage']]
            inv = t.empty(npatch * bs * bs, dtype=dtype, device=self.device)
        if pt.get('pos') is None and bs <= 160:
            npad = 16 * ((bs + 15) // 16)
            pt['pos'] = t.empty(npatch * npad * npad, dtype=t.int32, device=self.device)
            self._ck(self.lib.ocmp_patch_positions(npatch, bs, pt['dofs'].data_ptr(), pd['rowptr'].data_ptr(),
                                                   pd['colidx'].data_ptr(), pt['pos'].data_ptr(), st))
        setup = {'fp64': self.lib.ocmp_asm_setup, 'fp32': self.lib.ocmp_asm_setup_f32,
                 'bf16': self.lib.ocmp_asm_setup_bf16}[pt['storage']]
        self._ck(setup(npatch, bs, pt['dofs'].data_ptr(), pd['rowptr'].data_ptr(), pd['colidx'].data_ptr(),
                       mat.values.data_ptr(), _ptr(fm), inv.data_ptr(), _ptr(pt.get('pos')), st))
        return inv

    def fp32_copy(self, values, buf=None):
        """FP32 copy of a matrix value array for the operator applications inside the multigrid cycle
        (OCMP_SPMV_FP32=1); ``buf`` is reused when it fits."""
        t = self.torch
        if buf is None or buf.numel() != values.numel():
            buf = t.empty(values.numel(), dtype=t.float32, device=self.device)
        self._ck(self.lib.ocmp_to_f32(values.numel(), values.data_ptr(), buf.data_ptr(), self._stream()))
        self.launches += 1
        return buf

    def _patches(self, fes, kind: str, vmask=None) -> dict:
        """vmask: optional boolean mask over the mesh vertices — keep only the patches of those vertices (the
        element-partitioned smoother applies the patches of the vertices a rank owns)."""
        sd = self.space_data(fes)
        key = 'patch_' + kind + ('' if vmask is None else '_owned')
        if key in sd:
            return sd[key]
        cd = fes.cell_dofs
        if kind == 'cell':
            dofs = cd
        elif kind in ('star', 'vanka'):
            from .patches import vertex_patch_dofs
            dofs = vertex_patch_dofs(fes, kind, vmask)
        else:
            m = fes.mesh
            ne, nvc = m.cells.shape
            order = np.argsort(m.cells.ravel(), kind='stable')
            vsorted = m.cells.ravel()[order]
            cell_of = order // nvc
            start = np.searchsorted(vsorted, np.arange(m.nv + 1))
            cnt = np.diff(start)
            maxc = int(cnt.max())
            # padded (nv, maxc) table of star cells, -1 padded
            star = -np.ones((m.nv, maxc), dtype=np.int64)
            pos = np.arange(len(vsorted)) - start[vsorted]
            star[vsorted, pos] = cell_of
            vbig = int(np.argmax(cnt))                 # largest closed star: skip the full table if it cannot fit
            if len(np.unique(cd[star[vbig, :cnt[vbig]]])) > 160:
                out = self._patches(fes, 'star', vmask)
                sd[key] = out
                return out
            big = np.where(star[:, :, None] >= 0, cd[np.maximum(star, 0)], np.iinfo(np.int32).max)
            big = np.sort(big.reshape(m.nv, -1), axis=1)
            dup = np.concatenate([np.zeros((m.nv, 1), bool), big[:, 1:] == big[:, :-1]], axis=1)
            big[dup] = np.iinfo(np.int32).max
            big = np.sort(big, axis=1)
            bs = int((big < np.iinfo(np.int32).max).sum(axis=1).max())
            dofs = big[:, :bs].copy()
            dofs[dofs == np.iinfo(np.int32).max] = -1
            if vmask is not None:
                dofs = dofs[np.asarray(vmask, dtype=bool)]
        if dofs.shape[1] > 160 and kind != 'cell':
            # closed vertex stars can exceed what the register-tiled inversion holds (bs <= 160) — always in 3-D
            # (Q2/Q1 Taylor-Hood: 402 DOFs): take the open star (82-89 DOFs there), and cell patches if even that
            # is too large
            out = self._patches(fes, 'star' if kind == 'vertex' else 'cell', vmask if kind == 'vertex' else None)
            sd[key] = out
            return out
        storage = patch_storage()
        align = _STORAGE_ALIGN[storage]
        if dofs.shape[1] % align:
            # patch stride padded so that the columns of the stored inverses stay 16-byte aligned: k_patch_apply_stream
            # moves whole columns with the bulk-copy engine (FP64: even stride, FP32 / bf16: multiple of 4 / 8)
            npad = -dofs.shape[1] % align
            dofs = np.concatenate([dofs, -np.ones((dofs.shape[0], npad), dtype=dofs.dtype)], axis=1)
        dofs = np.ascontiguousarray(dofs, dtype=np.int32)
        mult = np.bincount(dofs[dofs >= 0].ravel(), minlength=fes.ndof).astype(np.float64)
        inc_ptr, inc_idx = patch_incidence(dofs, fes.ndof)
        out = dict(npatch=dofs.shape[0], bs=dofs.shape[1], dofs=self._up(dofs),
                   wgt=self._up(1.0 / np.maximum(mult, 1.0)), storage=storage,
                   inc_ptr=self._up(inc_ptr), inc_idx=self._up(inc_idx),
                   ybuf=self.zeros(max(1, dofs.size)))
        sd[key] = out
        return out

    def _system(self, mat, fm, pre) -> System:
        s = System()
        s.nrows = mat.height
        if not getattr(mat, 'matrix_free', False):    # a matrix-free operator has no CSR arrays (krylov sets apply_fn)
            pd = self.pattern_data(mat.space)
            s.rowptr, s.colidx, s.vals = pd['rowptr'].data_ptr(), pd['colidx'].data_ptr(), mat.values.data_ptr()
            if pd.get('runs') is not None:
                s.run_len, s.run_shift, s.run_nc = pd['runs'][0].data_ptr(), pd['runs'][1], pd['runs'][2]
                s.run_grouped = pd['runs'][1] if pd['runs'][3] else 0
        s.freemask = _ptr(fm)
        s.pre_kind = 0
        if pre is not None:
            s.pre_kind = pre.kind
            if pre.kind == 1:
                s.dinv = pre.dinv.data_ptr()
            elif pre.kind == 5:
                if not hasattr(pre.fact, 'handle'):
                    raise NotImplementedError("Preconditioner type 'direct': the system is too large to factorise "
                                              '(OCMP_DIRECT_MAX_GB)')
                pre.handle = pre.fact.handle()            # kept alive by the state for the duration of the solve
                s.direct = C.addressof(pre.handle)
            elif pre.kind == 3:
                top = pre.levels[pre.nlevels - 1].sys
                for name in ('npatch', 'bs', 'patch_dofs', 'inv_blocks', 'patch_weight', 'inv_storage', 'vals32',
                             'patch_inc_ptr', 'patch_inc_idx', 'patch_ybuf'):
                    setattr(s, name, getattr(top, name))
                s.nlevels = pre.nlevels
                s.levels = C.addressof(pre.levels)
            else:
                s.npatch, s.bs = pre.npatch, pre.bs
                s.patch_dofs, s.inv_blocks = pre.pdofs.data_ptr(), pre.inv.data_ptr()
                s.patch_weight = _ptr(pre.wgt)
                s.inv_storage = STORAGE_ID[getattr(pre, 'storage', 'fp64')]
                s.patch_inc_ptr, s.patch_inc_idx = pre.inc_ptr.data_ptr(), pre.inc_idx.data_ptr()
                s.patch_ybuf = pre.ybuf.data_ptr()
        return s

    def krylov(self, kind, mat, b, x, pre, freedofs, tol, maxit, initialize, printrates, damp=1.0, restart=None):
        state = None
        if pre is not None:
            if pre.state is None:
                pre.Update()
            state = pre.state
        fm = self._mask(mat.space, freedofs)
        if fm is None and state is not None:
            fm = state.fm
        if initialize:
            x.zero_()
        kid = {'cg': 0, 'gmres': 1, 'minres': 3, 'richardson': 2}[kind]
        restart = min(maxit, 200) if restart is None else restart
        sys_ = self._system(mat, fm, state)
        wl = self.lib.ocmp_krylov_work_len(mat.height, kid, restart)
        work = self.torch.empty(wl, dtype=self.torch.float64, device=self.device)
        it, res = C.c_int(0), C.c_double(0.0)
        keep = None
        if getattr(mat, 'matrix_free', False):
            # BilinearForm(nonassemble=True).mat as the Krylov operator: ocmp_krylov calls back with device pointers
            # into its work array / x / b; the form's action runs on them (k_coef + k_lin on the same stream)
            n, spans, errors = mat.height, (work, x, b), []

            def view(ptr):
                for tns in spans:
                    off = ptr - tns.data_ptr()
                    if 0 <= off and off + 8 * n <= 8 * tns.numel() and off % 8 == 0:
                        return tns[off // 8: off // 8 + n]
                raise RuntimeError('matrix-free callback: pointer outside the Krylov vectors')

            def callback(_ctx, xp, yp, _stream):
                try:
                    mat.bf.apply_arrays(view(xp), view(yp))
                    return 0
                except Exception as exc:                # an exception must not unwind through the C frames
                    errors.append(exc)
                    return 1
            keep = APPLY_FN(callback)
            sys_.apply_fn = C.cast(keep, C.c_void_p).value
            rc = self.lib.ocmp_krylov(C.byref(sys_), kid, b.data_ptr(), x.data_ptr(), float(tol), int(maxit),
                                      int(restart), float(damp), work.data_ptr(), wl, C.byref(it), C.byref(res),
                                      self._stream())
            if errors:
                raise errors[0]
            self._ck(rc)
        else:
            self._ck(self.lib.ocmp_krylov(C.byref(sys_), kid, b.data_ptr(), x.data_ptr(), float(tol), int(maxit),
                                      int(restart), float(damp), work.data_ptr(), wl, C.byref(it), C.byref(res),
                                      self._stream()))
        self.last_iters, self.last_resid = it.value, res.value
        if hasattr(state, 'note_solve'):
            state.note_solve(it.value)
        if printrates:
            print('{}: {} iterations, residual {:.3e}'.format(kind, it.value, res.value))

    def krylov_history(self, cap: int = 4096):
        """Relative preconditioned residuals of the last Krylov solve (one per iteration)."""
        buf = (C.c_double * cap)()
        n = self.lib.ocmp_krylov_history(buf, cap)
        return [buf[i] for i in range(min(n, cap))]

    def factorize(self, mat, freedofs):
        """``a.mat.Inverse(freedofs)`` (reference base_model.py:908-916): band LU with partial pivoting on the device
        (direct.BandLU). Systems whose band does not fit the budget get the Krylov stand-in, which checks the TRUE
        residual and raises when it cannot reach direct-solver quality."""
        from .direct import BandLU, DirectSolveTooLarge
        try:
            return BandLU(self, mat, freedofs)
        except DirectSolveTooLarge as exc:
            import warnings
            warnings.warn('opencmp_b200: {} — using GMRES + additive Schwarz with true-residual refinement '
                          'instead'.format(exc))
            return _KrylovInverse(self, mat, freedofs)

    def solve_free(self, mat, r, out, freedofs):
        """``mat.Inverse(freedofs) * r`` in one call."""
        self.factorize(mat, freedofs).solve(r, out)


class _KrylovInverse:
    """Stand-in for a factorisation that does not fit: GMRES(100) + vertex-patch additive Schwarz, restarted on the
    true free-dof residual until it is at round-off level; raises instead of returning an inexact answer."""

    def __init__(self, be, mat, freedofs):
        self.be, self.mat = be, mat
        self.free = np.ones(mat.height, bool) if freedofs is None else freedofs
        self.Update()

    def Update(self):
        class _P:
            pass
        self.p = _P()
        self.p.state = self.be.precond_setup(self.mat, 'asm', self.free)

    def solve(self, r, out, target: float = 1e-13, accept: float = 1e-10):
        be = self.be
        fm = self.p.state.fm
        out.zero_()
        res = be.zeros(self.mat.height)
        rnorm = float((r * fm).norm()) if fm is not None else float(r.norm())
        if rnorm == 0.0:
            return
        rel, best = 1.0, 1.0
        for _ in range(8):
            be.krylov('gmres', self.mat, r, out, self.p, self.free, max(1e-3 * target / rel, 1e-14), 4000, False,
                      False, restart=100)
            be.spmv(self.mat, out, res)
            res.mul_(-1.0).add_(r)
            rel = float((res * fm).norm() if fm is not None else res.norm()) / rnorm
            if rel <= target or rel > 0.5 * best:
                break
            best = rel
        if not rel <= accept:
            raise RuntimeError('opencmp_b200: the Krylov stand-in for mat.Inverse stalled at a true relative residual '
                               'of {:.2e} (needs {:.0e}); raise OCMP_DIRECT_MAX_GB to factorise instead'
                               .format(rel, accept))


# ---- primitives used by the element-partitioned multigrid driver (dist_mg.py) -----------------------------------
def _cuda_csr_handle(self, m):
    m = m.tocsr()
    m.sort_indices()
    return dict(nrows=m.shape[0], rowptr=self._up(m.indptr.astype(np.int32)), colidx=self._up(m.indices.astype(np.int32)),
                vals=self._up(m.data.astype(np.float64)))


def _cuda_csr_mult(self, h, x, y):
    self._ck(self.lib.ocmp_spmv(h['nrows'], h['rowptr'].data_ptr(), h['colidx'].data_ptr(), h['vals'].data_ptr(),
                                x.data_ptr(), y.data_ptr(), self._stream()))


def _cuda_patch_state(self, fes, vmask):
    # the caller owns the returned record and with it the inverse buffer; the topology arrays are shared
    return dict(self._patches(fes, os.environ.get('OCMP_PATCH', 'vertex'), vmask), inv=None)


def _cuda_patch_setup(self, mat, pt, fm):
    pt['inv'] = self._invert_patches(pt, self.pattern_data(mat.space), mat, fm, pt.get('inv'))


def _cuda_patch_apply(self, pt, r, z):
    apply = {'fp64': self.lib.ocmp_asm_apply, 'fp32': self.lib.ocmp_asm_apply_f32,
             'bf16': self.lib.ocmp_asm_apply_bf16}[pt['storage']]
    self._ck(apply(pt['npatch'], pt['bs'], pt['dofs'].data_ptr(), pt['inv'].data_ptr(), pt['inc_ptr'].data_ptr(),
                   pt['inc_idx'].data_ptr(), pt['ybuf'].data_ptr(), r.data_ptr(), z.data_ptr(), z.numel(),
                   self._stream()))


def _cuda_patch_count(self, pt, n):
    d = pt['dofs'].reshape(-1).to(self.torch.int64)
    d = d[d >= 0]
    return self.torch.bincount(d, minlength=n).to(self.torch.float64)


def _cuda_dense_inverse(self, mat, fm):
    t = self.torch
    pat = mat.space.pattern()
    n = mat.height
    rows = self._up(np.repeat(np.arange(n), np.diff(pat.rowptr)).astype(np.int64))
    cols = self._up(pat.colidx.astype(np.int64))
    dense = t.zeros((n, n), dtype=t.float64, device=self.device)
    dense[rows, cols] = mat.values
    dense = dense * fm[:, None] * fm[None, :] + t.diag(1.0 - fm)
    inv = t.linalg.inv(dense).contiguous()
    return dict(nrows=n, rowptr=self._up((np.arange(n + 1, dtype=np.int64) * n).astype(np.int32)),
                colidx=self._up(np.tile(np.arange(n, dtype=np.int32), n)), vals=inv.view(-1))


for _name, _fn in (('csr_handle', _cuda_csr_handle), ('csr_mult', _cuda_csr_mult), ('patch_state', _cuda_patch_state),
                   ('patch_setup', _cuda_patch_setup), ('patch_apply', _cuda_patch_apply),
                   ('patch_count', _cuda_patch_count), ('dense_inverse', _cuda_dense_inverse)):
    setattr(CudaBackend, _name, _fn)
