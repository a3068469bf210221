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


def _header_structs():
    """{struct name: [field names in order]} parsed from include/opencmp_b200.h."""
    header = open(os.path.join(ROOT, 'include', 'opencmp_b200.h')).read()
    header = re.sub(r'/\*.*?\*/', '', header, flags=re.S)
    out = {}
    for body, name in re.findall(r'typedef struct[^{;]*\{(.*?)\}\s*(\w+)\s*;', header, flags=re.S):
        fields = []
        for decl in body.split(';'):
            decl = decl.strip()
            if not decl:
                continue
            # "const double* fvec[OCMP_MAX_FVEC]" / "int dim, kind, nq" / "ocmp_system sys"
            first, *rest = decl.split(',')
            names = [re.findall(r'(\w+)\s*(?:\[[^\]]*\])?\s*$', first.strip())[0]]
            names += [re.findall(r'(\w+)', r)[-1] if '[' not in r else re.findall(r'(\w+)\s*\[', r)[0] for r in rest]
            fields += names
        out[name] = fields
    return out


def test_struct_mirrors_match_header_field_by_field():
    """Every field of the four plan / system structs sits at the same offset in the ctypes mirror as in C."""
    import subprocess
    import tempfile
    from opencmp_b200.backend import CoefPlan, ContractPlan, System, MGLevel
    mirrors = {'ocmp_coef_plan': CoefPlan, 'ocmp_contract_plan': ContractPlan, 'ocmp_system': System,
               'ocmp_mg_level': MGLevel}
    structs = _header_structs()
    assert set(mirrors) <= set(structs)
    lines = []
    for sname, fields in structs.items():
        if sname in mirrors:
            assert fields == [f[0] for f in mirrors[sname]._fields_], sname
            lines += ['printf("{0}.{1} %zu\\n", offsetof({0}, {1}));'.format(sname, f) for f in fields]
    src = '#include <stdio.h>\n#include <stddef.h>\n#include "opencmp_b200.h"\nint main(){' + ''.join(lines) + \
          'return 0;}\n'
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, 'probe.c')
        open(c, 'w').write(src)
        exe = os.path.join(d, 'probe')
        subprocess.run(['gcc', '-I', os.path.join(ROOT, 'include'), c, '-o', exe], check=True)
        out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout.split('\n')
    got = dict(l.split() for l in out if l.strip())
    for sname, cls in mirrors.items():
        for f in cls._fields_:
            assert int(got['{}.{}'.format(sname, f[0])]) == getattr(cls, f[0]).offset, (sname, f[0])


def test_ctypes_signatures_match_header_arity():
    """The argtypes the binding declares have as many entries as the C prototype has parameters."""
    from opencmp_b200.backend import load_library
    lib = load_library()
    header = open(os.path.join(ROOT, 'include', 'opencmp_b200.h')).read()
    header = re.sub(r'/\*.*?\*/', '', header, flags=re.S)
    checked = 0
    for name, params in re.findall(r'\b(ocmp_[a-z0-9_]+)\s*\(([^)]*)\)\s*;', header):
        fn = getattr(lib, name)
        if fn.argtypes is None:
            continue
        nparams = 0 if params.strip() in ('', 'void') else len(params.split(','))
        assert len(fn.argtypes) == nparams, name
        checked += 1
    assert checked >= 20


def test_call_sites_pass_as_many_arguments_as_the_prototypes_take():
    """Static (ast) check of every ``<x>.lib.ocmp_*(...)`` / ``lib.ocmp_*(...)`` call in the host code: the GPU-only
    branches cannot run in the CPU suite, so an argument dropped in a refactor would only show on the GPU box."""
    import ast
    import glob
    header = open(os.path.join(ROOT, 'include', 'opencmp_b200.h')).read()
    header = re.sub(r'/\*.*?\*/', '', header, flags=re.S)
    arity = {name: (0 if params.strip() in ('', 'void') else len(params.split(',')))
             for name, params in re.findall(r'\b(ocmp_[a-z0-9_]+)\s*\(([^)]*)\)\s*;', header)}
    files = glob.glob(os.path.join(ROOT, 'opencmp_b200', '*.py')) + glob.glob(os.path.join(ROOT, 'tools', '*.py')) + \
        [os.path.join(ROOT, 'bench.py'), os.path.join(ROOT, '__graft_entry__.py')]
    seen = set()
    for path in files:
        for node in ast.walk(ast.parse(open(path).read(), path)):
            if isinstance(node, ast.Call) and isinstance(node.func, ast.Attribute) and \
                    node.func.attr.startswith('ocmp_'):
                name = node.func.attr
                assert name in arity, '{}:{} calls undeclared {}'.format(path, node.lineno, name)
                if any(isinstance(a, ast.Starred) for a in node.args):
                    continue
                assert len(node.args) == arity[name] and not node.keywords, \
                    '{}:{} {} takes {} arguments'.format(os.path.relpath(path, ROOT), node.lineno, name, arity[name])
                seen.add(name)
    assert len(seen) >= 18
