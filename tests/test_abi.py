"""The C-ABI library loads on a machine without a GPU and exports every symbol include/opencmp_b200.h declares."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as g
    g.build()
    header = open(os.path.join(ROOT, 'include', 'opencmp_b200.h')).read()
    header = re.sub(r'/\*.*?\*/', '', header, flags=re.S)
    declared = sorted(set(re.findall(r'\b(ocmp_[a-z0-9_]+)\s*\(', header)))
    assert len(declared) >= 18
    lib = ctypes.CDLL(os.path.join(ROOT, 'opencmp_b200', 'lib', 'libopencmp_b200.so'))
    missing = [name for name in declared if not hasattr(lib, name)]
    assert not missing, missing
    lib.ocmp_version.restype = ctypes.c_int
    assert lib.ocmp_version() == 100
    from opencmp_b200.backend import EXPORTED
    assert set(EXPORTED) <= set(declared)


def test_struct_mirrors_match_header_sizes():
    """ctypes mirrors must have the C layout: compile a tiny probe against the header."""
    import subprocess
    import tempfile
    from opencmp_b200.backend import CoefPlan, ContractPlan, System, MGLevel
    src = '#include <stdio.h>\n#include "opencmp_b200.h"\nint main(){printf("%zu %zu %zu %zu\\n", ' \
          'sizeof(ocmp_coef_plan), sizeof(ocmp_contract_plan), sizeof(ocmp_system), sizeof(ocmp_mg_level));return 0;}\n'
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, 'probe.c')
        open(c, 'w').write(src)
        exe = os.path.join(d, 'probe')
        subprocess.run(['gcc', '-I', os.path.join(ROOT, 'include'), c, '-o', exe], check=True)
        out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout.split()
    sizes = [int(x) for x in out]
    assert sizes == [ctypes.sizeof(CoefPlan), ctypes.sizeof(ContractPlan), ctypes.sizeof(System),
                     ctypes.sizeof(MGLevel)]


def test_product_path_fails_loudly_without_gpu():
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    from opencmp_b200.backend import CudaBackend
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        CudaBackend()


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, 'opencmp_b200')
    for fn in os.listdir(pkg):
        if fn.endswith('.py'):
            txt = open(os.path.join(pkg, fn)).read()
            assert not re.search(r'^\s*(from|import)\s+oracle\b', txt, flags=re.M), fn
