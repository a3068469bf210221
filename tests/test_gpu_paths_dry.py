"""Dry run of the GPU-only host code on a machine without a GPU.

``CudaBackend`` is instantiated with host tensors in place of device buffers and a *null* library in place of
``libopencmp_b200.so``: every ``ocmp_*`` call is checked against the ctypes signature the binding declares (argument
count, and that each argument converts to its declared C type) and then returns success without computing anything.
The bodies of the ``-m gpu`` tests are then executed. Their numerical assertions cannot hold (nothing is computed) —
``AssertionError`` is therefore expected and ignored — but every other exception (a renamed variable, a missing
attribute, a wrong argument list, a dtype the binding cannot convert) is a bug in the host code that would otherwise
only surface on the GPU box. This is how the broken ``Preconditioner(a, ...)`` of the 3-D workload was found."""
import ctypes as C

import numpy as np
import pytest
import torch

import opencmp_b200.backend as backend_mod
from opencmp_b200.backend import CudaBackend, load_library


class _NullFn:
    def __init__(self, name, real, log):
        self.name, self.real, self.log = name, real, log

    def __call__(self, *args):
        at = self.real.argtypes
        if at is not None:
            assert len(args) == len(at), '{} takes {} arguments, got {}'.format(self.name, len(at), len(args))
            for a, t in zip(args, at):
                if hasattr(a, '_obj') or isinstance(a, (C.Structure, C.Array, C._Pointer)):   # byref() / ctypes objects
                    continue
                try:
                    t.from_param(a)
                except (TypeError, C.ArgumentError) as exc:        # pragma: no cover - the failure path
                    raise TypeError('{}: argument {!r} does not convert to {}: {}'.format(self.name, a, t, exc))
        self.log.append(self.name)
        if self.name == 'ocmp_krylov':
            iters, resid = args[10], args[11]
            iters._obj.value, resid._obj.value = 1, 0.0
            system = args[0]._obj
            if system.apply_fn:
                # matrix-free Krylov operator: call back once the way the C driver does — y = A x with x the solution
                # vector and y the first work vector (device pointers; host addresses here)
                rc = backend_mod.APPLY_FN(system.apply_fn)(None, args[3], args[8], None)
                assert rc == 0, 'matrix-free operator callback failed on the null device'
                self.log.append('apply_fn')
        if self.name == 'ocmp_krylov_work_len':
            return self.real(*args)                      # pure host arithmetic
        if self.name == 'ocmp_last_error':
            return b''
        if self.name in ('ocmp_launch_count', 'ocmp_krylov_history', 'ocmp_halo_plan'):
            return 0
        if self.name == 'ocmp_profile_bytes':
            return 0.0
        return 0


class _NullLib:
    def __init__(self):
        self._real = load_library()
        self.calls = []

    def __getattr__(self, name):
        if not name.startswith('ocmp_'):
            raise AttributeError(name)
        return _NullFn(name, getattr(self._real, name), self.calls)


class DryCudaBackend(CudaBackend):
    """CudaBackend whose buffers live on the host and whose library does nothing."""

    def __init__(self, device=None):
        self.torch = torch
        self.lib = _NullLib()
        self.device = torch.device('cpu')
        import weakref
        self._mesh_cache, self._space_cache = {}, {}
        self._plan_cache = weakref.WeakKeyDictionary()
        self._dbuf = None
        self._abuf = None
        self._scal = torch.zeros(64, dtype=torch.float64)
        self.launches = 0
        self.last_iters = 0
        self.last_resid = 0.0
        self.chunk_bytes = 48 << 20

    def _stream(self):
        return None

    def integrate(self, program):
        super().integrate(program)             # the host path runs; nothing is summed
        return 1.0                             # keeps norms non-zero so the callers' divisions go through


@pytest.fixture
def dry(monkeypatch):
    monkeypatch.setattr(backend_mod, 'CudaBackend', DryCudaBackend)
    # nothing is assembled, so dense coarse-level operators are singular: keep the call, skip the factorisation
    monkeypatch.setattr(torch.linalg, 'inv', lambda a: a.clone())
    return DryCudaBackend


def _run_ignoring_numerics(fn, *args, **kw):
    try:
        fn(*args, **kw)
    except AssertionError:
        pass                                   # numerical comparisons are meaningless without kernels
    except (np.linalg.LinAlgError, FloatingPointError, ZeroDivisionError):
        pass                                   # ditto (host-side least squares / norms of all-zero data)


def _gpu_tests(module):
    out = []
    marks = getattr(module, 'pytestmark', None)
    module_gpu = marks is not None and 'gpu' in str(marks)
    for name in sorted(dir(module)):
        fn = getattr(module, name)
        if not (name.startswith('test_') and callable(fn)):
            continue
        own = [m for m in getattr(fn, 'pytestmark', [])]
        if not (module_gpu or any(m.name == 'gpu' for m in own)):
            continue
        params = [m for m in own if m.name == 'parametrize']
        out.append((name, fn, params))
    return out


def _invoke(fn, params, monkeypatch, ngs_fixture):
    import inspect
    import itertools
    sig = inspect.signature(fn).parameters
    names, values = [], []
    for m in params:
        names.append(m.args[0])
        values.append(list(m.args[1]))
    for combo in itertools.product(*values) if values else [()]:
        kw = {}
        for nm, val in zip(names, combo):
            keys = [k.strip() for k in nm.split(',')] if isinstance(nm, str) else list(nm)
            if len(keys) == 1:
                kw[keys[0]] = val
            else:
                kw.update(zip(keys, getattr(val, 'values', val)))      # pytest.param(...) or plain tuple
        if 'monkeypatch' in sig:
            kw['monkeypatch'] = monkeypatch
        if 'cuda_backend' in sig:
            kw['cuda_backend'] = ngs_fixture
        _run_ignoring_numerics(fn, **kw)


@pytest.mark.parametrize('modname', ['test_gpu_parity', 'test_golden_programs', 'test_golden_fixtures',
                                     'test_zz_gpu_late_additions', 'test_zzz_gpu_unmeasured_kernels',
                                     'test_zzzz_gpu_matrix_free', 'test_gpu_direct', 'test_gpu_krylov', 'test_gpu_deterministic', 'test_gpu_dimgen'])
def test_gpu_test_bodies_execute_on_a_null_device(dry, monkeypatch, modname):
    import importlib
    import opencmp_b200.ngs as ngs
    module = importlib.import_module(modname)
    tests = _gpu_tests(module)
    assert tests, modname
    old = ngs._backend
    ngs.set_backend(DryCudaBackend())
    try:
        for name, fn, params in tests:
            _invoke(fn, params, monkeypatch, ngs)
    finally:
        ngs.set_backend(old)


def test_fp32_patch_storage_host_path(dry, monkeypatch):
    """OCMP_PATCH_FP32=1 routes set-up and application to the *_f32 entry points with a float32 buffer."""
    import opencmp_b200.ngs as ngs
    import cases
    monkeypatch.setenv('OCMP_PATCH_FP32', '1')
    be = DryCudaBackend()
    old = ngs._backend
    ngs.set_backend(be)
    try:
        c = cases.stokes(cases.channel_mesh(4), 2, True)
        c['a'].Assemble()
        pre = ngs.Preconditioner(c['a'], 'asm')
        pre.Update()
        assert 'ocmp_asm_setup_f32' in be.lib.calls and 'ocmp_asm_setup' not in be.lib.calls
        st = pre.state
        assert st.inv.dtype == torch.float32 and st.bs % 4 == 0 and st.storage == 'fp32'
        sys_ = be._system(c['a'].mat, None, st)
        assert sys_.inv_storage == 1 and sys_.bs == st.bs
    finally:
        ngs.set_backend(old)


class _FakeEvent:
    def __init__(self, enable_timing=False):
        pass

    def record(self, *a):
        pass

    def elapsed_time(self, other):
        return 1.0


class _FakeStream:
    cuda_stream = 0

    def synchronize(self):
        pass


@pytest.mark.parametrize('argv', [['--workload', 'ins2d', '--N', '8', '--steps', '1', '--warmup', '1', '--no-cpu'],
                                  ['--workload', 'ins2d', '--N', '8', '--steps', '1', '--warmup', '1', '--no-cpu',
                                   '--precond-storage', 'fp64', '--full-mg-setup', '--lag-smoother'],
                                  ['--N', '4', '--steps', '1', '--warmup', '1', '--no-cpu', '--no-secondary',
                                   '--cpu-N', '4'],
                                  ['--N', '4', '--steps', '1', '--warmup', '1', '--no-cpu', '--no-secondary',
                                   '--cpu-N', '4', '--precond-storage', 'bf16']])
def test_bench_main_executes_on_a_null_device(dry, monkeypatch, capsys, argv):
    """bench.py's own arm, start to JSON line, with the CUDA runtime calls stubbed: the line must carry every key of
    the bench contract (values are meaningless here)."""
    import json
    import sys
    import bench
    real_tensor = torch.tensor
    monkeypatch.setattr(torch.cuda, 'set_device', lambda *a: None)
    monkeypatch.setattr(torch.cuda, 'synchronize', lambda *a: None)
    monkeypatch.setattr(torch.cuda, 'empty_cache', lambda *a: None)
    monkeypatch.setattr(torch.cuda, 'Event', _FakeEvent)
    monkeypatch.setattr(torch.cuda, 'current_stream', lambda *a: _FakeStream())
    monkeypatch.setattr(torch.Tensor, 'pin_memory', lambda self: self)
    monkeypatch.setattr(torch, 'tensor', lambda data, **kw: real_tensor(data, **{k: v for k, v in kw.items()
                                                                                  if k != 'device'}))
    monkeypatch.setattr(bench, 'fp64_peak_tflops', lambda t: {'burst': 40.0, 'sustained': 39.0, 'how': 'stub'})
    monkeypatch.setattr(sys, 'argv', ['bench.py'] + argv)
    for k, v in (('OCMP_PATCH_STORAGE', 'fp64'), ('OCMP_SPMV_FP32', '0')):   # bench sets them; restored afterwards
        monkeypatch.setenv(k, v)
    monkeypatch.setenv('OCMP_MG_REUSE_COARSE', '1')
    monkeypatch.setenv('OCMP_MG_LAG', '0')
    for k in ('WORLD_SIZE', 'RANK', 'LOCAL_RANK'):
        monkeypatch.delenv(k, raising=False)
    try:
        bench.main()
    except ZeroDivisionError:
        pytest.skip('all-zero fields: a norm of the workload vanished before the JSON line (numerics, not host logic)')
    line = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    for key in ('metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling',
                'vs_baseline', 'dtype', 'data', 'config', 'roofline', 'e2e', 'gpu_launches', 'clocks',
                'roofline_patch_apply', 'roofline_spmv', 'roofline_assembly', 'fp64_peak_tflops',
                'same_size_as_cpu_sample'):
        assert key in line, key
    assert line['config']['workload'] and line['dtype'] == 'f64' and line['higher_is_better'] is False
    want = 'fp64' if 'fp64' in argv else 'bf16' if 'bf16' in argv else 'fp32'
    assert line['config']['precond_storage'] == {'patch_inverses': want,
                                                 'level_matrices_in_cycle': 'fp64' if want == 'fp64' else 'fp32'}
    for key in ('bound', 'achieved', 'peak', 'unit', 'frac', 'traffic'):
        assert key in line['roofline'], key
    assert line['roofline_assembly']['flops_per_assembly'] > 0 and line['roofline_assembly']['bound'] == 'fp64'
    for key in ('value', 'unit', 'h2d_bytes_per_step', 'd2h_bytes_per_step'):
        assert key in line['e2e'], key
    assert 'error' not in line['same_size_as_cpu_sample'], line['same_size_as_cpu_sample']


def test_bench_default_line_carries_the_2d_workload(dry, monkeypatch, capsys):
    """The default (3-D) line on one GPU also reports the 2-D INS step as the object ``ins2d``."""
    import json
    import sys
    import bench
    real_tensor = torch.tensor
    monkeypatch.setattr(torch.cuda, 'set_device', lambda *a: None)
    monkeypatch.setattr(torch.cuda, 'synchronize', lambda *a: None)
    monkeypatch.setattr(torch.cuda, 'empty_cache', lambda *a: None)
    monkeypatch.setattr(torch.cuda, 'Event', _FakeEvent)
    monkeypatch.setattr(torch.cuda, 'current_stream', lambda *a: _FakeStream())
    monkeypatch.setattr(torch.Tensor, 'pin_memory', lambda self: self)
    monkeypatch.setattr(torch, 'tensor', lambda data, **kw: real_tensor(data, **{k: v for k, v in kw.items()
                                                                                  if k != 'device'}))
    monkeypatch.setattr(bench, 'fp64_peak_tflops', lambda t: {'burst': 40.0, 'sustained': 39.0, 'how': 'stub'})
    monkeypatch.setitem(bench.DEFAULTS, 'ins2d', dict(N=8, order=3))
    monkeypatch.setattr(sys, 'argv', ['bench.py', '--N', '4', '--steps', '1', '--warmup', '1', '--no-cpu', '--cpu-N', '4'])
    for k, v in (('OCMP_PATCH_STORAGE', 'fp64'), ('OCMP_SPMV_FP32', '0'), ('OCMP_MG_REUSE_COARSE', '1'),
                 ('OCMP_MG_LAG', '0')):
        monkeypatch.setenv(k, v)
    for k in ('WORLD_SIZE', 'RANK', 'LOCAL_RANK'):
        monkeypatch.delenv(k, raising=False)
    try:
        bench.main()
    except ZeroDivisionError:
        pytest.skip('all-zero fields: a norm of the workload vanished before the JSON line')
    line = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    assert 'INS-DIM 3D' in line['config']['workload']
    two = line['ins2d']
    assert 'error' not in two, two
    for key in ('config', 'value', 'e2e', 'problem', 'roofline_patch_apply', 'roofline_spmv', 'roofline_assembly',
                'kernel_time_share', 'gpu_launches'):
        assert key in two, key
    assert 'Taylor-Green 2D' in two['config']['workload']


@pytest.mark.parametrize('which', ['ins2d', 'ins3d_dim'])
@pytest.mark.parametrize('fp32', ['0', '1', '2'])
def test_partitioned_workloads_execute_on_a_null_device(dry, monkeypatch, which, fp32):
    """Set-up and one time step of the element-partitioned workloads (world size 1: no communicator) through the level
    array handed to the C driver (dist_mg._native_levels / _gmres_native)."""
    import opencmp_b200.ngs as ngs
    from opencmp_b200.dist_workload import DistributedINS, DistributedINSDIM3D
    monkeypatch.setenv('OCMP_PATCH_STORAGE', ['fp64', 'fp32', 'bf16'][int(fp32)])
    monkeypatch.setenv('OCMP_SPMV_FP32', '0' if fp32 == '0' else '1')
    be = DryCudaBackend()
    old = ngs._backend
    ngs.set_backend(be)
    try:
        d = DistributedINS(8, 1, 0) if which == 'ins2d' else DistributedINSDIM3D(4, 1, 0)
        d.step()
        top, arr = d.mg._native
        assert top.pre_kind == 3 and top.nlevels == len(d.mg.levels) >= 2
        assert top.inv_storage == int(fp32) and arr[top.nlevels - 1].sys.inv_storage == int(fp32)
        assert bool(arr[top.nlevels - 1].sys.vals32) == (fp32 != '0') and not arr[0].sys.vals32
        assert 'ocmp_krylov' in be.lib.calls
        assert ('ocmp_asm_setup_f32' in be.lib.calls) == (fp32 == '1')
        assert ('ocmp_asm_setup_bf16' in be.lib.calls) == (fp32 == '2')
        assert ('ocmp_to_f32' in be.lib.calls) == (fp32 != '0')
        assert d.mg.levels[-1].patches['bs'] % [2, 4, 8][int(fp32)] == 0
    finally:
        ngs.set_backend(old)


@pytest.mark.parametrize('fp32', ['0', '1'])
def test_single_gpu_multigrid_state_on_a_null_device(dry, monkeypatch, fp32):
    """MultigridState (ngs.Preconditioner(a, 'multigrid') on one GPU): level array incl. the FP32 options."""
    import opencmp_b200.ngs as ngs
    from opencmp_b200.workloads import INSTaylorGreen
    monkeypatch.setenv('OCMP_PATCH_FP32', fp32)
    monkeypatch.setenv('OCMP_SPMV_FP32', fp32)
    be = DryCudaBackend()
    old = ngs._backend
    ngs.set_backend(be)
    try:
        w = INSTaylorGreen(16)
        w.step()
        w.step()
        st = w.pre.state
        # coarse levels are set up once (their operators only depend on dt); the finest level on every Update()
        setup = 'ocmp_asm_setup_f32' if fp32 == '1' else 'ocmp_asm_setup'
        assert st.nlevels == 3 and st.coarse_setups == 1 and st.updates >= 2
        assert be.lib.calls.count(setup) == st.updates + (st.nlevels - 2)
        assert st.kind == 3 and st.nlevels >= 2
        top = st.levels[st.nlevels - 1].sys
        assert top.inv_storage == int(fp32) and bool(top.vals32) == (fp32 == '1') and not st.levels[0].sys.vals32
        sys_ = be._system(w.a.mat, st.fm, st)
        assert sys_.pre_kind == 3 and sys_.inv_storage == int(fp32) and bool(sys_.vals32) == (fp32 == '1')
    finally:
        ngs.set_backend(old)


def test_smoke_executes_on_a_null_device(dry):
    """__graft_entry__.smoke(): the host path runs to the comparison with the oracle (which cannot hold here)."""
    import __graft_entry__ as g
    with pytest.raises(AssertionError):
        g.smoke()


def _bench_rank(rank, world, port, argv, out):
    """One rank of ``torchrun ... bench.py --gpus 2`` on a null device: gloo instead of NCCL, CUDA runtime stubbed."""
    import json
    import os
    import sys
    import io
    import contextlib
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    backend_mod.CudaBackend = DryCudaBackend
    torch.linalg.inv = lambda a: a.clone()
    real_tensor, real_init = torch.tensor, dist.init_process_group
    torch.cuda.set_device = lambda *a: None
    torch.cuda.synchronize = lambda *a: None
    torch.cuda.empty_cache = lambda *a: None
    torch.cuda.Event = _FakeEvent
    torch.cuda.current_stream = lambda *a: _FakeStream()
    torch.Tensor.pin_memory = lambda self: self
    torch.tensor = lambda data, **kw: real_tensor(data, **{k: v for k, v in kw.items() if k != 'device'})
    dist.init_process_group = lambda backend=None, **kw: real_init('gloo', rank=rank, world_size=world)
    sys.argv = ['bench.py'] + argv
    import bench
    bench.fp64_peak_tflops = lambda t: {'burst': 40.0, 'sustained': 39.0, 'how': 'stub'}
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        bench.main()
    if rank == 0:
        out['line'] = json.loads(buf.getvalue().strip().splitlines()[-1])
    else:
        out['other'] = buf.getvalue().strip()


@pytest.mark.parametrize('argv', [['--gpus', '2', '--workload', 'ins2d', '--N', '8', '--steps', '1', '--warmup', '1'],
                                  ['--gpus', '2', '--N', '4', '--steps', '1', '--warmup', '1', '--layout', 'bricks'],
                                  ['--gpus', '2', '--N', '6', '--steps', '1', '--warmup', '1']])
def test_bench_two_ranks_execute_on_a_null_device(argv):
    """The multi-GPU arm of bench.py (element-partitioned workload, communicator set-up through the C ABI, halo plans,
    level array of the distributed C driver, max-over-ranks timing, rank 0 prints) with two gloo ranks."""
    import socket
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    out = mp.get_context('spawn').Manager().dict()          # never fork a multi-threaded pytest process
    mp.spawn(_bench_rank, args=(2, port, argv, out), nprocs=2, join=True)
    line = out['line']
    assert out.get('other', '') == ''                       # only rank 0 prints
    assert line['n_gpus'] == 2 and line['scaling'] == 'weak' and line['problem']['ranks'] == 2
    assert line['problem']['global_dofs'] > line['problem']['dofs'] * 1.2
    assert 'element-partitioned' in line['config']['parallelism']


def test_free_dof_masks_are_uploaded_once(dry):
    be = DryCudaBackend()
    a = np.array([True, False, True, True, False] * 1000)
    m1 = be._mask(None, a.copy())
    m2 = be._mask(None, a.copy())                       # a fresh host array with the same bits: no second upload
    b = a.copy()
    b[7] = not b[7]
    m3 = be._mask(None, b)
    assert m1 is m2 and m3 is not m1
    assert np.array_equal(m1.numpy(), a.astype(np.float64)) and np.array_equal(m3.numpy(), b.astype(np.float64))
    for k in range(12):                                 # the cache stays small and keeps the most recent entries
        c = a.copy()
        c[k] = not c[k]
        be._mask(None, c)
    assert len(be._mask_cache) == 8 and be._mask(None, c) is be._mask_cache[-1][2]


def test_lagged_smoother_on_a_null_device(dry, monkeypatch):
    """OCMP_MG_LAG=1 on one GPU: the krylov call reports its iteration count to the multigrid state, and the finest
    level is inverted once while the (null) solves keep reporting the same count."""
    import opencmp_b200.ngs as ngs
    from opencmp_b200.workloads import INSTaylorGreen
    monkeypatch.setenv('OCMP_MG_LAG', '1')
    be = DryCudaBackend()
    old = ngs._backend
    ngs.set_backend(be)
    try:
        w = INSTaylorGreen(8)
        w.step()
        w.step()
        st = w.pre.state
        assert st.updates >= 2 and st.lag.enabled and st.lag.fresh_setups == 1 and st.lag.last == 1
        assert be.lib.calls.count('ocmp_asm_setup') == 1
    finally:
        ngs.set_backend(old)


def test_matrix_free_krylov_callback_on_a_null_device(dry):
    """``ocmp_system.apply_fn``: the ctypes callback CudaBackend.krylov installs for a ``nonassemble`` operator resolves
    the raw pointers the C driver hands it to views of the Krylov vectors and runs the form's action on them."""
    import opencmp_b200.ngs as ngs
    from test_matrix_free import krylov_matrix_free_vs_csr
    be = DryCudaBackend()
    old = ngs._backend
    ngs.set_backend(be)
    try:
        krylov_matrix_free_vs_csr()
    finally:
        ngs.set_backend(old)
    calls = be.lib.calls
    assert calls.count('apply_fn') == 2 and calls.count('ocmp_krylov') == 4
    k = calls.index('apply_fn')
    assert 'ocmp_contract_vector' in calls[:k] and calls[k + 1:].count('ocmp_krylov') >= 1
