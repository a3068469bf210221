import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


# Order of the GPU suite (the driver runs it with -x): kernel-level parity first, then the direct solve, the
# Krylov / multigrid solver-level cases, and the reference-model golden fixtures last — so that one solver-level
# failure cannot hide the assembly-parity verdict.
_GPU_ORDER = ['test_gpu_parity.py::test_assembly', 'test_gpu_parity.py::test_integrate', 'test_golden_programs',
              'test_gpu_deterministic', 'test_zz_gpu_late_additions', 'test_zzz_gpu_unmeasured_kernels',
              'test_zzzz_gpu_matrix_free', 'test_gpu_dimgen', 'test_gpu_direct', 'test_gpu_krylov', 'test_gpu_parity.py', 'test_golden_fixtures']


def pytest_collection_modifyitems(config, items):
    def rank(item):
        if item.get_closest_marker('gpu') is None:
            return -1
        for i, key in enumerate(_GPU_ORDER):
            if key in item.nodeid:
                return i
        return len(_GPU_ORDER) - 2
    items.sort(key=rank)                       # stable: the order inside a group is unchanged


@pytest.fixture
def oracle_backend():
    """Route the NGSolve-style front end through the NumPy oracle (CPU tests only)."""
    import opencmp_b200.ngs as ngs
    from oracle.backend import OracleBackend
    old = ngs._backend
    ngs.set_backend(OracleBackend())
    yield ngs
    ngs.set_backend(old)


@pytest.fixture
def cuda_backend():
    import opencmp_b200.ngs as ngs
    from opencmp_b200.backend import CudaBackend
    old = ngs._backend
    ngs.set_backend(CudaBackend())
    yield ngs
    ngs.set_backend(old)
