"""Form programs: the lowered, flat representation of a weak form that the C-ABI consumes.

OpenCMP builds symbolic forms once and then only mutates leaves between ``Assemble()`` calls (Parameters t/dt,
DOF vectors of captured GridFunctions, Dirichlet data — SURVEY 3.2, reference solvers/base_solver.py:240-254,
transient_multistep.py:150-169, models/ins.py:354). A form is therefore lowered ONCE into

* integrals (cell / interior facet / boundary facet) each with
* a list of *entries*  (test_row, trial_row) -> output slot   [bilinear]   or  (test_row) -> slot   [linear],
  rows being physical operator rows of the space (value / d/dx_a of a block, side 0 = this cell, 1 = neighbour),
* a register-machine *bytecode* evaluated at every quadrature point that fills the output slots from coordinates,
  normal, mesh size, run-time parameters and field values,
* the list of *field slots* (which DOF vector, which block, which row, which side).

Per step only the parameter array and the DOF vectors change.
"""
from __future__ import annotations

import math
import weakref
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

# opcode table (shared with csrc/ocmp_kernels.cuh and oracle/interp.py) -----------------------------------------
OPS = ['CONST', 'PARAM', 'COORD', 'NORMAL', 'MESHSIZE', 'FIELD', 'ADD', 'SUB', 'MUL', 'DIV', 'NEG', 'ABS', 'SQRT',
       'SIN', 'COS', 'TAN', 'EXP', 'LOG', 'POW', 'IFPOS', 'MIN', 'MAX', 'TANH', 'ERF', 'FLOOR', 'CEIL', 'ROUND',
       'TRUNC', 'SGN', 'ATAN', 'OUT', 'MOV', 'MEASURE']
OP = {name: i for i, name in enumerate(OPS)}
UNARY = {'neg': 'NEG', 'abs': 'ABS', 'sqrt': 'SQRT', 'sin': 'SIN', 'cos': 'COS', 'tan': 'TAN', 'exp': 'EXP',
         'log': 'LOG', 'tanh': 'TANH', 'erf': 'ERF', 'floor': 'FLOOR', 'ceil': 'CEIL', 'round': 'ROUND',
         'trunc': 'TRUNC', 'sgn': 'SGN', 'atan': 'ATAN'}
BINARY = {'add': 'ADD', 'sub': 'SUB', 'mul': 'MUL', 'div': 'DIV', 'pow': 'POW', 'min': 'MIN', 'max': 'MAX'}
MAX_REGS = 48

_PY_UNARY = {'neg': lambda a: -a, 'abs': abs, 'sqrt': math.sqrt, 'sin': math.sin, 'cos': math.cos, 'tan': math.tan,
             'exp': lambda a: math.exp(a) if a < 709.0 else float('inf'), 'log': math.log, 'tanh': math.tanh, 'erf': math.erf, 'floor': math.floor,
             'ceil': math.ceil, 'round': lambda a: float(np.round(a)), 'trunc': math.trunc,
             'sgn': lambda a: float((a > 0) - (a < 0)), 'atan': math.atan}
_PY_BINARY = {'add': lambda a, b: a + b, 'sub': lambda a, b: a - b, 'mul': lambda a, b: a * b,
              'div': lambda a, b: a / b, 'pow': lambda a, b: a ** b, 'min': min, 'max': max}


class Coef:
    """Hash-consed node of a coefficient expression DAG (no trial/test functions inside)."""
    __slots__ = ('op', 'args', 'val', '_key', '__weakref__')
    _table = weakref.WeakValueDictionary()

    def __new__(cls, op: str, args: tuple = (), val=None):
        key = (op, tuple(id(a) for a in args), val)
        hit = Coef._table.get(key)
        if hit is not None:
            return hit
        self = object.__new__(cls)
        self.op, self.args, self.val, self._key = op, tuple(args), val, key
        Coef._table[key] = self
        return self

    def __reduce__(self):
        # pickling (the reference's post-processing hands GridFunctions to a multiprocessing.Pool,
        # post_processing/output_conversions.py:192-199) must go through __new__ again to keep the hash-consing
        return (Coef, (self.op, self.args, self.val))

    # constructors with constant folding -------------------------------------------------------------------------
    @staticmethod
    def const(v) -> 'Coef':
        return Coef('const', (), float(v))

    def is_const(self, v: Optional[float] = None) -> bool:
        return self.op == 'const' and (v is None or self.val == v)

    @staticmethod
    def unary(op: str, a: 'Coef') -> 'Coef':
        if a.op == 'const':
            return Coef.const(_PY_UNARY[op](a.val))
        if op == 'neg' and a.op == 'neg':
            return a.args[0]
        return Coef(op, (a,))

    @staticmethod
    def binary(op: str, a: 'Coef', b: 'Coef') -> 'Coef':
        if a.op == 'const' and b.op == 'const':
            return Coef.const(_PY_BINARY[op](a.val, b.val))
        if op == 'add':
            if a.is_const(0.0):
                return b
            if b.is_const(0.0):
                return a
        elif op == 'sub':
            if b.is_const(0.0):
                return a
            if a.is_const(0.0):
                return Coef.unary('neg', b)
        elif op == 'mul':
            if a.is_const(0.0) or b.is_const(0.0):
                return ZERO
            if a.is_const(1.0):
                return b
            if b.is_const(1.0):
                return a
            if a.is_const(-1.0):
                return Coef.unary('neg', b)
            if b.is_const(-1.0):
                return Coef.unary('neg', a)
        elif op == 'div':
            if a.is_const(0.0):
                return ZERO
            if b.is_const(1.0):
                return a
        elif op == 'pow':
            if b.is_const(1.0):
                return a
            if b.is_const(2.0):
                return Coef.binary('mul', a, a)
        return Coef(op, (a, b))

    @staticmethod
    def ifpos(c: 'Coef', a: 'Coef', b: 'Coef') -> 'Coef':
        if c.op == 'const':
            return a if c.val > 0 else b
        return Coef('ifpos', (c, a, b))

    def __repr__(self) -> str:
        if self.op == 'const':
            return repr(self.val)
        if self.op in ('param', 'coord', 'normal', 'h', 'meas', 'field', 'piecewise'):
            return '{}({})'.format(self.op, self.val if self.op != 'param' else id(self.val))
        return '{}({})'.format(self.op, ', '.join(repr(a) for a in self.args))


ZERO = Coef.const(0.0)
ONE = Coef.const(1.0)


def coef_leaves(roots: Sequence[Coef], op: str) -> List[Coef]:
    seen, out, stack = set(), [], list(roots)
    while stack:
        n = stack.pop()
        if id(n) in seen:
            continue
        seen.add(id(n))
        if n.op == op:
            out.append(n)
        if n.op == 'piecewise':
            stack.extend(n.val[1])
        stack.extend(n.args)
    return out


def resolve_piecewise(c: Coef, kind: str, rid: int, memo: Optional[dict] = None) -> Coef:
    """Replace region-wise leaves (``CoefficientFunction([per-region list])``) by the entry of region ``rid``."""
    memo = {} if memo is None else memo
    if id(c) in memo:
        return memo[id(c)]
    if c.op == 'piecewise':
        k, lst = c.val
        out = resolve_piecewise(lst[rid], kind, rid, memo) if (k == kind and rid < len(lst)) else \
            (ZERO if k == kind else c)
    elif not c.args:
        out = c
    else:
        new = tuple(resolve_piecewise(a, kind, rid, memo) for a in c.args)
        if all(n is o for n, o in zip(new, c.args)):
            out = c
        elif c.op == 'ifpos':
            out = Coef.ifpos(*new)
        elif len(new) == 1:
            out = Coef.unary(c.op, new[0])
        else:
            out = Coef.binary(c.op, new[0], new[1])
    memo[id(c)] = out
    return out


class Bytecode:
    """code (ninstr, 4) int32 rows [op | dst << 8, a, b, c]; consts float64; params / fields = leaf objects."""

    def __init__(self, outputs: Sequence[Coef]):
        self.consts: List[float] = []
        self.params: List[object] = []
        self.fields: List[tuple] = []          # (gf, block, row, side)
        const_ix: Dict[float, int] = {}
        param_ix: Dict[int, int] = {}
        field_ix: Dict[tuple, int] = {}
        # topological order
        order: List[Coef] = []
        state: Dict[int, int] = {}
        for root in outputs:
            stack = [(root, 0)]
            while stack:
                node, i = stack.pop()
                if i == 0:
                    if state.get(id(node)):
                        continue
                    state[id(node)] = 1
                if i < len(node.args):
                    stack.append((node, i + 1))
                    child = node.args[i]
                    if not state.get(id(child)):
                        stack.append((child, 0))
                else:
                    order.append(node)
        last_use: Dict[int, int] = {}
        for t, node in enumerate(order):
            for a in node.args:
                last_use[id(a)] = t
        nout_base = len(order)
        for k, root in enumerate(outputs):
            last_use[id(root)] = nout_base + k
        free: List[int] = []
        nreg = 0
        reg: Dict[int, int] = {}
        code: List[Tuple[int, int, int, int]] = []
        for t, node in enumerate(order):
            a = b = c = 0
            if node.op == 'const':
                a = const_ix.setdefault(node.val, len(const_ix))
                if a == len(self.consts):
                    self.consts.append(node.val)
                op = OP['CONST']
            elif node.op == 'param':
                a = param_ix.setdefault(id(node.val), len(param_ix))
                if a == len(self.params):
                    self.params.append(node.val)
                op = OP['PARAM']
            elif node.op == 'coord':
                a, op = int(node.val), OP['COORD']
            elif node.op == 'normal':
                a, op = int(node.val), OP['NORMAL']
            elif node.op == 'h':
                op = OP['MESHSIZE']
            elif node.op == 'meas':
                op = OP['MEASURE']
            elif node.op == 'field':
                gf, blk, row, side = node.val
                key = (id(gf), blk, row, side)
                a = field_ix.setdefault(key, len(field_ix))
                if a == len(self.fields):
                    self.fields.append((gf, blk, row, side))
                op = OP['FIELD']
            elif node.op == 'ifpos':
                a, b, c = (reg[id(x)] for x in node.args)
                op = OP['IFPOS']
            elif node.op in UNARY:
                a, op = reg[id(node.args[0])], OP[UNARY[node.op]]
            elif node.op in BINARY:
                a, b = reg[id(node.args[0])], reg[id(node.args[1])]
                op = OP[BINARY[node.op]]
            elif node.op == 'piecewise':
                raise ValueError('region-wise coefficient used where no region is known')
            else:
                raise ValueError('cannot lower coefficient op {}'.format(node.op))
            # release argument registers whose last use is this instruction
            for x in set(id(y) for y in node.args):
                if last_use.get(x) == t:
                    free.append(reg[x])
            if free:
                dst = free.pop()
            else:
                dst = nreg
                nreg += 1
            reg[id(node)] = dst
            code.append((op | (dst << 8), a, b, c))
        for k, root in enumerate(outputs):
            code.append((OP['OUT'] | (0 << 8), k, reg[id(root)], 0))
        if nreg > MAX_REGS:
            raise ValueError('coefficient program needs {} registers (max {})'.format(nreg, MAX_REGS))
        self.nreg = max(nreg, 1)
        self.nout = len(outputs)
        self.code = np.array(code, dtype=np.int32).reshape(-1, 4)
        self.consts_arr = np.array(self.consts if self.consts else [0.0], dtype=np.float64)


class Integral:
    """One lowered integral. ``kind``: 'cell', 'ifacet' (interior facets), 'bfacet' (boundary facets).

    entries: int32 (nent, 3) rows [test_row, trial_row, slot]; trial_row = -1 for linear forms; test_row = -1 too for
    plain functionals (``Integrate``). Rows >= nrows of the space address the neighbour side of an interior facet.
    items: the cells / facets integrated over (int32), None = all.
    """

    def __init__(self, kind: str, deg: int, entries: np.ndarray, prog: Bytecode, items: Optional[np.ndarray]):
        self.kind = kind
        self.deg = int(deg)
        self.entries = np.ascontiguousarray(entries, dtype=np.int32).reshape(-1, 3)
        self.prog = prog
        self.items = None if items is None else np.ascontiguousarray(items, dtype=np.int32)


class FormProgram:
    """All integrals of one bilinear form, linear form or functional over one space."""

    def __init__(self, fes, arity: int, integrals: List[Integral]):
        self.fes = fes
        self.arity = arity            # 2 bilinear, 1 linear, 0 functional
        self.integrals = integrals

    def param_values(self, integral: Integral) -> np.ndarray:
        vals = [float(p.Get()) for p in integral.prog.params]
        return np.array(vals if vals else [0.0], dtype=np.float64)
