"""Host-side check of the launch plans the CUDA backend builds (no GPU needed): for every GPU parity case the plan
builder runs with host tensors standing in for device buffers, and the compile-time limits and alignment assumptions
of the kernels in csrc/ocmp_assembly.cu are checked on the resulting structs — shared-memory budget of k_contract,
MAX_ROWS / MAX_FSLOTS / OCMP_MAX_REGS, tile coverage and the 16-byte alignment of every tile edge. The numbers are computed by the oracle backend; only the plan construction of
``CudaBackend`` runs (it is pure host logic)."""
import weakref

import numpy as np
import pytest
import torch

import cases
from oracle.backend import OracleBackend
from opencmp_b200.backend import CudaBackend

MAX_ROWS, MAX_FSLOTS, MAX_REGS = 12, 40, 48          # csrc/ocmp_assembly.cu, include/opencmp_b200.h


def _dry_cuda_backend():
    be = object.__new__(CudaBackend)
    be.torch = torch
    be.device = torch.device('cpu')
    be._mesh_cache, be._space_cache = {}, {}
    be._plan_cache = weakref.WeakKeyDictionary()
    from test_gpu_paths_dry import _NullLib
    be.lib = _NullLib()                      # pattern_data marks component runs through the C ABI (a no-op here)
    be._stream = lambda: None
    return be


class PlanProbe(OracleBackend):
    """Oracle backend that also builds (and inspects) the CUDA launch plans of everything it assembles."""

    def __init__(self):
        self.dry = _dry_cuda_backend()
        self.seen = []

    def _probe(self, program):
        for integ, plan in zip(program.integrals, self.dry._plans(program)):
            self.seen.append(check_plan(program, integ, plan))

    def assemble_matrix(self, program, mat):
        self._probe(program)
        return super().assemble_matrix(program, mat)

    def assemble_vector(self, program, out):
        self._probe(program)
        return super().assemble_vector(program, out)

    def integrate(self, program):
        self._probe(program)
        return super().integrate(program)


def check_plan(program, integ, plan):
    from opencmp_b200.backend import contract_smem_bytes
    cp = plan['coef']
    assert cp.nfslots <= MAX_FSLOTS
    assert cp.nreg <= MAX_REGS
    for (gf, blk, row, side) in integ.prog.fields:
        assert gf.space.blocks[blk].basis.nrows <= MAX_ROWS
        assert row < gf.space.blocks[blk].basis.nrows
    xp = plan['contract']
    if xp is None:
        return 0
    fes = program.fes
    for b in fes.blocks:
        assert b.basis.nrows <= MAX_ROWS and b.basis.nrows < 256          # row count packed in 8 bits
        assert b.nloc + 3 < (1 << 15)                                     # padded block size packed above bit 16
    if program.arity == 1:
        return 0
    gs = cp.dim + 2 * cp.dim * cp.dim + 1
    smem = contract_smem_bytes(xp, xp.eb, gs)
    assert smem <= 220 * 1024, 'k_contract needs {} bytes of shared memory'.format(smem)
    assert xp.eb in (1, 2, 4, 8, 16) and xp.maxt in (1, 2, 4)
    assert xp.ntiles <= xp.maxt * (256 // xp.eb)                             # every tile has a thread
    assert xp.sbsz % 4 == 0 and xp.zsz % 4 == 0                              # tile edges are read as double2 pairs
    tb = plan['tables']
    t = tb['tiles']
    assert (t[:, 0] % 4 == 0).all() and (t[:, 1] % 4 == 0).all()             # 16-byte aligned B and Z offsets
    assert (tb['seg'] % 4 == 0).all()
    assert (t[:, 3] > 0).all() and (t[:, 2] + t[:, 3] <= len(tb['seg'])).all()
    rem_i, rem_j = t[:, 7] & 0xff, t[:, 7] >> 8
    assert ((rem_i >= 1) & (rem_i <= 4) & (rem_j >= 1) & (rem_j <= 4)).all()
    assert (t[:, 5] + rem_i <= fes.nloc).all() and (t[:, 6] + rem_j <= fes.nloc).all()
    assert int((rem_i * rem_j).sum()) == xp.nact                             # the tiles cover the active entries once
    return smem


def _programs(build, assemble=True):
    import opencmp_b200.ngs as ngs
    be = PlanProbe()
    old = ngs._backend
    ngs.set_backend(be)
    try:
        c = build()
        g = c['gfu']
        if c.get('noset'):
            pass
        elif 'exact' in c:
            g.components[0].Set(c['exact'], definedon=c['mesh'].Boundaries(c['dnames']))
        else:
            g.components[0].Set(c['uex'], definedon=c['mesh'].Boundaries(c['walls']))
        c['a'].Assemble()
        c['L'].Assemble()
    finally:
        ngs.set_backend(old)
    return be.seen


def _all_cases():
    from test_gpu_parity import CASES as established
    from test_zz_gpu_late_additions import CASES as late
    out = dict(established)
    out.update(late)
    return out


ALL = _all_cases()
# the 3-D cases take a while on the oracle; one hex and one tet case are enough for the plan limits
SLOW = {'ins_dim_3d_hex_q2q1'}


@pytest.mark.parametrize('name', sorted(n for n in ALL if n not in SLOW))
def test_launch_plans_respect_kernel_limits(name):
    seen = _programs(ALL[name])
    assert len(seen) > 0
    assert max(seen) > 0                 # at least one bilinear contraction plan was built and checked


def test_golden_program_plans_respect_kernel_limits():
    """The replayed reference-model programs (tests/golden/prog_*.npz) go through the same plan builder."""
    import test_golden_programs as tg
    import opencmp_b200.ngs as ngs
    assert set(tg.LATE) <= set(tg.FIXTURES)
    for name in tg.FIXTURES:
        be = PlanProbe()
        old = ngs._backend
        ngs.set_backend(be)
        try:
            tg._replay(ngs, name)
        finally:
            ngs.set_backend(old)
        assert be.seen, name


def test_edge_case_plans_respect_kernel_limits():
    """The edge cases of tests/test_edge_cases.py (one-cell meshes, order-0 DG, empty regions) through the same probe."""
    import opencmp_b200.ngs as ngs
    import test_edge_cases as e
    for name in e.EDGE_CASES:
        be = PlanProbe()
        old = ngs._backend
        ngs.set_backend(be)
        try:
            e.assemble_edge_case(ngs, name)
        finally:
            ngs.set_backend(old)
        assert be.seen and max(be.seen) > 0, name


def test_binary_searches_do_not_overflow_int32():
    """CSR positions reach 1.42e9 at the 3-D N = 64 workload (> 2^30): a midpoint written as (lo + hi) >> 1 wraps to a
    negative index there (the illegal address of round 1's N = 64 run, in k_patch_positions). Every bisection in csrc/
    has to form its midpoint as lo + ((hi - lo) >> 1); the 32-bit arithmetic is replayed here on the offending range."""
    import glob
    import os
    import re
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'opencmp_b200', 'csrc')
    srcs = glob.glob(os.path.join(root, '*.cu')) + glob.glob(os.path.join(root, '*.cuh'))
    assert srcs
    nmid = 0
    for f in srcs:
        code = '\n'.join(ln.split('//')[0] for ln in open(f).read().splitlines())      # comments stripped
        assert not re.search(r'\(\s*lo\s*\+\s*hi\s*\)\s*(>>\s*1|/\s*2)', code), f
        nmid += len(re.findall(r'lo \+ \(\(hi - lo\) >> 1\)', code))
    assert nmid >= 2
    lo, hi = np.int32(1_400_000_000), np.int32(1_420_000_000)
    with np.errstate(over='ignore'):
        assert (lo + hi) >> np.int32(1) < 0                       # what the old form did
        mid = lo + ((hi - lo) >> np.int32(1))
    assert lo <= mid <= hi


def test_matrix_free_action_plans_respect_kernel_limits():
    """``BilinearForm.Apply`` turns every trial row (both facet sides) into a field slot of k_coef: the slot count and
    the register budget of the bytecode stay within MAX_FSLOTS / OCMP_MAX_REGS for every matrix-free case, and k_lin's
    entry scan sees no contraction plan (arity 1)."""
    import opencmp_b200.ngs as ngs
    from test_matrix_free import CASES as MF, matrix_free_vs_csr
    for name in MF:
        be = PlanProbe()
        old = ngs._backend
        ngs.set_backend(be)
        try:
            n0 = None
            c = MF[name]()
            c['a'].Assemble()
            n0 = len(be.seen)
            matrix_free_vs_csr(c)
        finally:
            ngs.set_backend(old)
        assert len(be.seen) > n0, name
