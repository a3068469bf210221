import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


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
