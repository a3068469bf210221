"""Static check for undefined names in every Python source of the repo (symtable walk, no third-party linter in the
image). The CUDA-only branches of the host code (CudaBackend, MultigridState, the bench legs) cannot execute in the
CPU test run, so a renamed variable there would only surface on the GPU box; this catches that class of error here."""
import builtins
import glob
import os
import symtable

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FILES = sorted(glob.glob(os.path.join(ROOT, 'opencmp_b200', '*.py')) + glob.glob(os.path.join(ROOT, 'oracle', '*.py')) +
               glob.glob(os.path.join(ROOT, 'tests', '*.py')) + glob.glob(os.path.join(ROOT, 'tools', '*.py')) +
               glob.glob(os.path.join(ROOT, 'tests', 'golden', '*.py')) +
               [os.path.join(ROOT, 'bench.py'), os.path.join(ROOT, '__graft_entry__.py')])
BUILTINS = set(dir(builtins)) | {'__file__', '__name__', '__doc__', '__builtins__', '__class__'}


def _undefined(path):
    src = open(path).read()
    top = symtable.symtable(src, path, 'exec')
    module_names = {s.get_name() for s in top.get_symbols() if s.is_assigned() or s.is_imported() or
                    s.is_namespace() or s.is_parameter()}
    star = 'import *' in src
    bad = []

    def walk(tab):
        for s in tab.get_symbols():
            name = s.get_name()
            if not s.is_referenced():
                continue
            if tab.get_type() == 'module':
                unresolved = not (s.is_assigned() or s.is_imported() or s.is_namespace())
            else:
                unresolved = s.is_global() and not s.is_declared_global() and not s.is_assigned()
                if s.is_declared_global():
                    unresolved = name not in module_names
            if unresolved and name not in module_names and name not in BUILTINS and not star:
                bad.append('{}:{}: {}'.format(os.path.relpath(path, ROOT), tab.get_lineno(), name))
        for child in tab.get_children():
            walk(child)

    walk(top)
    return bad


def test_no_undefined_names():
    assert len(FILES) > 30
    bad = []
    for f in FILES:
        bad += _undefined(f)
    assert not bad, '\n'.join(bad)
