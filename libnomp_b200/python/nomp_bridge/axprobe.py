"""Structural recogniser for the spectral-element Ax operator: does this loop nest COMPUTE  w = A u ?

The reference accepts any C loop nest (reference python/loopy_api.py:769-821) and schedules it by annotation
(reference tests/sem.py:10-36); it has no Ax of its own.  Here the tuned kernel (nompk_ax_f64, include/nompk.h) must
be reached by every spelling of the operator, not only by the canonical kernel string of families.py: reordered
statements, other temporaries (scalars, arrays of any shape, none at all), `+=` into a zeroed w, g[e][f][k][j][i]
instead of flat subscripts, loops in another order.  Matching syntax cannot do that; this module decides by what the
kernel computes:

 1. a static gate admits only loop nests whose value is a polynomial of the inputs that does not depend on the data
    path: one outer loop over the elements, `for (v = 0; v < bound; v++)` loops, assignments with = += -= *=,
    expressions of + - * and unary minus, no branches, calls, divisions or data-dependent subscripts, subscripts
    affine in the element index, integer temporaries assigned once;
 2. the admitted nest is translated to Python and EXECUTED, with the n fixed at nomp_jit time, on three elements of
    random small integers (exact arithmetic: any order of the sums gives the same bits), with instrumented arrays
    that record what is read and written;
 3. the roles of the arrays follow from their footprints (n^2 -> D, E n^3 -> u, 6 E n^3 -> g, the written one -> w)
    and the result must equal the definition (Nekbone ax_e) exactly although w starts from
    random values (an accumulation into a w that the kernel did not zero fails here);
    for the fused form the reduce clause's variable must end as u . (A u).

By the polynomial identity lemma a nest that passes on random data is the operator; anything else keeps its generic
schedule.  Cost: 0.1 s (n = 8) to 0.6 s (n = 12) of the cold nomp_jit path, once per kernel (the on-disk JIT cache
stores the outcome); the token comparison of families.match_ax stays in front of it as the free fast path.
"""
from __future__ import annotations

import random
from fractions import Fraction
from typing import Dict, List, Optional

from . import cparse as c

_PROBE_ELEMENTS = 3
_SENTINEL = 1000003          # what an uninitialised temporary holds: reading one makes the comparison fail


class _NotAx(Exception):
    pass


class _Tracked(list):
    """A flat array that remembers the range of subscripts that were read and written."""

    def __init__(self, values):
        super().__init__(values)
        self.rlo = self.wlo = 1 << 62
        self.rhi = self.whi = -1

    def __getitem__(self, i):
        if i < self.rlo:
            self.rlo = i
        if i > self.rhi:
            self.rhi = i
        return list.__getitem__(self, i)

    def __setitem__(self, i, v):
        if i < self.wlo:
            self.wlo = i
        if i > self.whi:
            self.whi = i
        list.__setitem__(self, i, v)


class _Translator:
    """C loop nest (the subset of the static gate) -> Python source."""

    def __init__(self, func: c.Function, reduce_var: Optional[str], fixed: Optional[Dict[str, object]] = None):
        self.func, self.reduce_var = func, reduce_var
        # NOMP_JIT arguments: already literals in the body, but parameter dimensions (`g[E][6][n][n][n]`) still name them
        self.fixed = {k: int(v) for k, v in (fixed or {}).items() if isinstance(v, int) or (isinstance(v, float) and v == int(v))}
        self.params = {p.name: p for p in func.params}
        self.int_names = {p.name for p in func.params if not p.is_array and not p.ctype.is_float}
        self.float_scalars = {p.name for p in func.params if not p.is_array and p.ctype.is_float}
        self.local_arrays: Dict[str, List[str]] = {}
        self.lines: List[str] = []
        self.element_var: Optional[str] = None
        self.loop_vars: List[str] = []

    # -- expressions -------------------------------------------------------------------------------------------
    def index_expr(self, e: c.Node) -> str:
        """Integer expression over loop variables, integer parameters and integer temporaries."""
        if isinstance(e, c.Num):
            if not e.is_int:
                raise _NotAx("non-integer subscript")
            return str(e.value)
        if isinstance(e, c.Name):
            if e.id in self.fixed and e.id not in self.int_names:
                return str(self.fixed[e.id])
            if e.id not in self.int_names:
                raise _NotAx(f"subscript uses {e.id}")
            return f"c_{e.id}"
        if isinstance(e, c.BinOp) and e.op in ("+", "-", "*"):
            return f"({self.index_expr(e.left)} {e.op} {self.index_expr(e.right)})"
        if isinstance(e, c.UnOp) and e.op in ("-", "+"):
            return f"({e.op}{self.index_expr(e.operand)})"
        if isinstance(e, c.Cast) and not e.ctype.is_float and not e.ctype.ptr:
            return self.index_expr(e.operand)
        raise _NotAx("subscript is not a polynomial of the loop variables")

    def degree_in_element(self, e: c.Node) -> int:
        if isinstance(e, c.Name):
            if e.id == self.element_var:
                return 1
            return self.int_degree.get(e.id, 0)
        if isinstance(e, c.BinOp):
            a, b = self.degree_in_element(e.left), self.degree_in_element(e.right)
            return a + b if e.op == "*" else max(a, b)
        if isinstance(e, (c.UnOp, c.Cast)):
            return self.degree_in_element(e.operand)
        return 0

    def flat_index(self, base: str, index: List[c.Node]) -> str:
        for ix in index:
            if self.degree_in_element(ix) > 1 or self.elements_param in _names(ix):
                raise _NotAx("subscript is not affine in the element index, or uses the number of elements")
        if base in self.local_arrays:
            dims = self.local_arrays[base]
        else:
            p = self.params[base]
            dims = [None if d is None else self.index_expr(d) for d in p.dims] if p.dims else [None]
        if len(index) != len(dims):
            raise _NotAx(f"{base} is subscripted with {len(index)} of {len(dims)} indices")
        flat = self.index_expr(index[0])
        for d, ix in zip(dims[1:], index[1:]):
            flat = f"({flat} * {d} + {self.index_expr(ix)})"
        return flat

    def value_expr(self, e: c.Node) -> str:
        if isinstance(e, c.Num):
            if e.is_int:
                return str(e.value)
            v = Fraction(e.text.rstrip("fFlL"))
            return str(v.numerator) if v.denominator == 1 else f"Fraction({v.numerator}, {v.denominator})"
        if isinstance(e, c.Name):
            if e.id in self.params and self.params[e.id].is_array:
                raise _NotAx("array used as a value")
            return f"c_{e.id}"
        if isinstance(e, c.Subscript):
            if not isinstance(e.base, c.Name) or (e.base.id not in self.local_arrays and not
                                                  (e.base.id in self.params and self.params[e.base.id].is_array)):
                raise _NotAx("subscript of something that is not an array")
            return f"c_{e.base.id}[{self.flat_index(e.base.id, e.index)}]"
        if isinstance(e, c.BinOp) and e.op in ("+", "-", "*"):
            return f"({self.value_expr(e.left)} {e.op} {self.value_expr(e.right)})"
        if isinstance(e, c.UnOp) and e.op in ("-", "+"):
            return f"({e.op}{self.value_expr(e.operand)})"
        if isinstance(e, c.Cast) and e.ctype.is_float and not e.ctype.ptr:
            return self.value_expr(e.operand)
        raise _NotAx(f"expression outside the polynomial subset: {type(e).__name__}")

    # -- statements ----------------------------------------------------------------------------------------------
    def stmts(self, nodes: List[c.Node], depth: int, top: bool = False):
        pad = "  " * depth
        if not nodes:
            self.lines.append(pad + "pass")
        for n in nodes:
            if isinstance(n, c.For):
                if n.tag is not None or n.var in self.loop_vars or n.var in self.int_names:
                    raise _NotAx("loop variable reused")
                if not (isinstance(n.lo, c.Num) and n.lo.is_int and n.lo.value == 0):
                    raise _NotAx("loop does not start at 0")
                if top:
                    if self.element_var is not None or not isinstance(n.hi, c.Name) or n.hi.id not in self.int_names:
                        raise _NotAx("expected one outer loop over the elements")
                    self.element_var, self.elements_param = n.var, n.hi.id
                    hi = f"c_{n.hi.id}"
                else:
                    if self.elements_param in _names(n.hi):
                        raise _NotAx("inner loop bound depends on the number of elements")
                    hi = self.index_expr(n.hi)
                self.loop_vars.append(n.var)
                self.int_names.add(n.var)
                self.lines.append(f"{pad}for c_{n.var} in range({hi}):")
                self.stmts(n.body, depth + 1)
                self.int_names.discard(n.var)
                self.loop_vars.pop()
            elif top:
                raise _NotAx("statement outside the element loop")
            elif isinstance(n, c.Decl):
                if n.name in self.params or n.name in self.local_arrays:
                    raise _NotAx("temporary shadows another name")
                if n.ctype.ptr:
                    raise _NotAx("pointer temporary")
                if n.dims:
                    if not n.ctype.is_float or n.init is not None:
                        raise _NotAx("unsupported array temporary")
                    dims = [self.index_expr(d) for d in n.dims]
                    self.local_arrays[n.name] = dims
                    self.lines.append(f"{pad}c_{n.name} = [{_SENTINEL}] * ({' * '.join(dims)})")
                elif n.ctype.is_float:
                    self.float_scalars.add(n.name)
                    self.lines.append(f"{pad}c_{n.name} = {self.value_expr(n.init) if n.init is not None else _SENTINEL}")
                else:   # integer temporary: assigned once, by its initialiser, from integers only
                    if n.init is None or n.name in self.int_names:
                        raise _NotAx("integer temporary without an initialiser")
                    self.int_degree[n.name] = self.degree_in_element(n.init)
                    self.lines.append(f"{pad}c_{n.name} = {self.index_expr(n.init)}")
                    self.int_names.add(n.name)
            elif isinstance(n, c.Assign):
                if n.op not in ("=", "+=", "-=", "*="):
                    raise _NotAx(f"assignment operator {n.op}")
                if isinstance(n.target, c.Name):
                    if n.target.id not in self.float_scalars or n.target.id in self.params:
                        raise _NotAx("assignment to something that is not a floating-point temporary")
                    target = f"c_{n.target.id}"
                elif isinstance(n.target, c.Subscript) and isinstance(n.target.base, c.Name):
                    b = n.target.base.id
                    if b in self.params and (not self.params[b].is_array or self.params[b].ctype.const):
                        raise _NotAx("write to a read-only argument")
                    if b == self.reduce_var and n.op != "+=":
                        raise _NotAx("the reduction variable may only be accumulated into")
                    target = self.value_expr(n.target)
                else:
                    raise _NotAx("unsupported assignment target")
                value = self.value_expr(n.value)
                if self.reduce_var is not None and self.reduce_var in _names(n.value):
                    raise _NotAx("the reduction variable is read")
                self.lines.append(f"{pad}{target} {n.op} {value}")
            else:
                raise _NotAx(f"statement outside the polynomial subset: {type(n).__name__}")

    def translate(self) -> str:
        self.int_degree: Dict[str, int] = {}
        self.elements_param = None
        names = [p.name for p in self.func.params]
        self.lines = [f"def kernel({', '.join('c_' + x for x in names)}):"]
        self.stmts(self.func.body, 1, top=True)
        if self.element_var is None:
            raise _NotAx("no element loop")
        return "\n".join(self.lines) + "\n"


def _names(e) -> set:
    out = set()

    def visit(x):
        if isinstance(x, c.Name):
            out.add(x.id)
        elif isinstance(x, c.Subscript):
            visit(x.base)
            for i in x.index:
                visit(i)
        elif isinstance(x, c.BinOp):
            visit(x.left), visit(x.right)
        elif isinstance(x, (c.UnOp, c.Cast)):
            visit(x.operand)
        elif isinstance(x, c.Ternary):
            visit(x.cond), visit(x.then), visit(x.other)
        elif isinstance(x, c.Call):
            for a in x.args:
                visit(a)
    visit(e)
    return out


def _definition(n: int, E: int, u, g, D):
    """w = A u on E elements by the definition (plain loops over exact integers; layouts of include/nompk.h)."""
    n2, n3 = n * n, n * n * n
    w = [0] * (E * n3)
    rng = range(n)
    for e in range(E):
        ub, gb = e * n3, e * 6 * n3
        wr, ws, wt = [0] * n3, [0] * n3, [0] * n3
        for k in rng:
            for j in rng:
                for i in rng:
                    p = k * n2 + j * n + i
                    r = s = t = 0
                    for l in rng:
                        r += D[i * n + l] * u[ub + k * n2 + j * n + l]
                        s += D[j * n + l] * u[ub + k * n2 + l * n + i]
                        t += D[k * n + l] * u[ub + l * n2 + j * n + i]
                    g1, g2, g3 = g[gb + p], g[gb + n3 + p], g[gb + 2 * n3 + p]
                    g4, g5, g6 = g[gb + 3 * n3 + p], g[gb + 4 * n3 + p], g[gb + 5 * n3 + p]
                    wr[p] = g1 * r + g2 * s + g3 * t
                    ws[p] = g2 * r + g4 * s + g5 * t
                    wt[p] = g3 * r + g5 * s + g6 * t
        for k in rng:
            for j in rng:
                for i in rng:
                    acc = 0
                    for l in rng:
                        acc += D[l * n + i] * wr[k * n2 + j * n + l] + D[l * n + j] * ws[k * n2 + l * n + i] + D[l * n + k] * wt[l * n2 + j * n + i]
                    w[ub + k * n2 + j * n + i] = acc
    return w


def probe_ax(func: c.Function, n: int, reduce_var: Optional[str] = None,
             fixed: Optional[Dict[str, object]] = None) -> Optional[Dict[str, str]]:
    """{role: parameter name} for roles w, u, g, D, E (and pap) if `func` -- the untransformed loop nest with the JIT-fixed
    n already substituted -- computes the Ax operator (and, with `reduce_var`, accumulates u . (A u) into it); else None."""
    arrays = [p for p in func.params if p.is_array]
    ints = [p for p in func.params if not p.is_array and not p.ctype.is_float]
    want_arrays = 4 + (1 if reduce_var else 0)
    if len(arrays) != want_arrays or len(ints) != 1 or len(func.params) != want_arrays + 1:
        return None
    if any(p.ctype.base != "double" or p.ctype.ptr != 1 for p in arrays):
        return None
    if reduce_var is not None and reduce_var not in {p.name for p in arrays}:
        return None
    try:
        source = _Translator(func, reduce_var, fixed).translate()
        scope: Dict[str, object] = {"Fraction": Fraction, "__builtins__": {"range": range}}
        exec(compile(source, "<nomp ax probe>", "exec"), scope)   # text generated above from the AST, identifiers only
        kernel = scope["kernel"]
    except (_NotAx, SyntaxError, ValueError, RecursionError):
        return None

    E, n3 = _PROBE_ELEMENTS, n * n * n
    rng = random.Random(0x5EED + n)
    size = 6 * E * n3
    data: Dict[str, _Tracked] = {}
    for p in arrays:
        if p.name == reduce_var:
            data[p.name] = _Tracked([0])
        else:                                   # every array gets values of every role's range; sized for the largest role
            data[p.name] = _Tracked([rng.randint(-3, 3) for _ in range(size)])
    before = {k: list(v) for k, v in data.items()}
    try:
        kernel(*[data[p.name] if p.is_array else E for p in func.params])
    except Exception:                           # IndexError, a sentinel that is not a number, ...
        return None

    written = [k for k, v in data.items() if v.whi >= 0 and k != reduce_var]
    if len(written) != 1:
        return None
    w = written[0]
    if data[w].wlo != 0 or data[w].whi != E * n3 - 1 or any(v.rlo < 0 or v.wlo < 0 for v in data.values() if v.rhi >= 0 or v.whi >= 0):
        return None
    roles: Dict[str, str] = {"w": w, "E": ints[0].name}
    for k, v in data.items():
        if k in (w, reduce_var):
            continue
        if v.rlo != 0:
            return None
        role = {n * n: "D", E * n3: "u", 6 * E * n3: "g"}.get(v.rhi + 1)
        if role is None or role in roles:
            return None
        roles[role] = k
    if not all(r in roles for r in ("u", "g", "D")):
        return None
    want = _definition(n, E, before[roles["u"]], before[roles["g"]], before[roles["D"]])
    got = list(data[w])
    if got[:E * n3] != want or got[E * n3:] != before[w][E * n3:]:
        return None
    if reduce_var is not None:
        u = before[roles["u"]]
        if data[reduce_var].whi != 0 or list(data[reduce_var]) != [sum(a * b for a, b in zip(u, want))]:
            return None
        roles["pap"] = reduce_var
    return roles
