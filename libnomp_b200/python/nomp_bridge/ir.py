"""Kernel object handed to user transform scripts, and the loop transformations they may apply.

The reference hands a loopy TranslationUnit to the user's Python function (reference src/loopy.c:203-236) and the
test scripts call exactly: knl.default_entrypoint.all_inames(), lp.split_iname(...), lp.tag_inames(...)
(reference tests/nomp_api_100.py:11-20, nomp_api_225.py:9-18, nomp_api_300.py:10-48, nomp_api_400.py:9-18,
tests/sem.py:10-36).  This module provides those operations on our own loop-nest IR (the C AST of cparse.py);
the `loopy` package next to this one re-exports them under loopy's names.

Transformations are functional: each returns a new Kernel and leaves its argument untouched, like loopy.
"""
from __future__ import annotations

import copy
import fnmatch
import re
from typing import Dict, Iterable, List, Optional

from . import cparse as c

VALID_TAG = re.compile(r"^(for|unr|g\.[0-2]|l\.[0-2])$")


class KernelError(Exception):
    """Raised for invalid transformations (surfaces as NOMP_PY_CALL_FAILURE in libnomp)."""


def walk(nodes: Iterable[c.Node]):
    """Yield every statement node, depth first."""
    for n in nodes:
        yield n
        if isinstance(n, (c.For, c.Bind)):
            yield from walk(n.body)
        elif isinstance(n, c.If):
            yield from walk(n.then)
            yield from walk(n.other)


def map_expr(e: Optional[c.Node], fn):
    """Rebuild expression e bottom-up, applying fn to every node."""
    if e is None:
        return None
    if isinstance(e, c.BinOp):
        e = c.BinOp(e.op, map_expr(e.left, fn), map_expr(e.right, fn))
    elif isinstance(e, c.UnOp):
        e = c.UnOp(e.op, map_expr(e.operand, fn))
    elif isinstance(e, c.Ternary):
        e = c.Ternary(map_expr(e.cond, fn), map_expr(e.then, fn), map_expr(e.other, fn))
    elif isinstance(e, c.Subscript):
        e = c.Subscript(map_expr(e.base, fn), [map_expr(i, fn) for i in e.index])
    elif isinstance(e, c.Call):
        e = c.Call(e.func, [map_expr(a, fn) for a in e.args])
    elif isinstance(e, c.Cast):
        e = c.Cast(e.ctype, map_expr(e.operand, fn))
    return fn(e)


def map_stmts(nodes: List[c.Node], efn) -> List[c.Node]:
    """Apply efn to every expression of every statement (in place on a private copy)."""
    out = []
    for n in nodes:
        if isinstance(n, c.Assign):
            out.append(c.Assign(map_expr(n.target, efn), n.op, map_expr(n.value, efn)))
        elif isinstance(n, c.Decl):
            out.append(c.Decl(n.ctype, n.name, [map_expr(d, efn) for d in n.dims], map_expr(n.init, efn)))
        elif isinstance(n, c.For):
            out.append(c.For(n.var, n.vtype, map_expr(n.lo, efn), map_expr(n.hi, efn), map_stmts(n.body, efn), n.tag))
        elif isinstance(n, c.Bind):
            out.append(c.Bind(n.var, n.vtype, map_expr(n.value, efn), map_expr(n.hi, efn), map_stmts(n.body, efn)))
        elif isinstance(n, c.If):
            out.append(c.If(map_expr(n.cond, efn), map_stmts(n.then, efn), map_stmts(n.other, efn)))
        else:
            out.append(n)
    return out


def expr_names(e: Optional[c.Node]) -> set:
    found = set()

    def visit(x):
        if isinstance(x, c.Name):
            found.add(x.id)
        return x

    map_expr(e, visit)
    return found


class _Entrypoint:
    """What `knl.default_entrypoint` returns: just enough of loopy's LoopKernel for the transform scripts."""

    def __init__(self, knl: "Kernel"):
        self._knl = knl
        self.name = knl.name

    def all_inames(self):
        return frozenset(self._knl.inames())

    @property
    def inames(self):
        return {n: None for n in self._knl.inames()}

    @property
    def args(self):
        return list(self._knl.func.params)


class Kernel:
    """A parsed nomp kernel plus its schedule annotations (tags on loops)."""

    def __init__(self, func: c.Function, source: str = "", target: str = "cuda"):
        self.func = func
        self.source = source
        self.target = target
        self.reduction = None       # (var name, op) once the reduce clause has been realised
        self.fixed = {}             # JIT-fixed parameters: name -> value
        self.annotations = {}       # key -> value from annotate clauses (kept for the family recogniser)

    # -- loopy-like surface ----------------------------------------------------------------------------
    @property
    def name(self) -> str:
        return self.func.name

    @property
    def default_entrypoint(self) -> _Entrypoint:
        return _Entrypoint(self)

    def copy(self) -> "Kernel":
        return copy.deepcopy(self)

    # -- queries ---------------------------------------------------------------------------------------
    def loops(self) -> List[c.For]:
        return [n for n in walk(self.func.body) if isinstance(n, c.For)]

    def inames(self) -> List[str]:
        seen = []
        for l in self.loops():
            if l.var not in seen:
                seen.append(l.var)
        return seen

    def tags(self) -> Dict[str, Optional[str]]:
        return {l.var: l.tag for l in self.loops()}

    def all_names(self) -> set:
        names = {p.name for p in self.func.params}
        for n in walk(self.func.body):
            if isinstance(n, (c.For, c.Bind)):
                names.add(n.var)
            elif isinstance(n, c.Decl):
                names.add(n.name)
        return names


# ----------------------------------------------------------------------------------------------------
# transformations
# ----------------------------------------------------------------------------------------------------

def _ceil_div(a: c.Node, b: int) -> c.Node:
    return c.BinOp("/", c.BinOp("+", a, c.Num(str(b - 1))), c.Num(str(b)))


def _is_zero(e: c.Node) -> bool:
    return isinstance(e, c.Num) and e.is_int and e.value == 0


def split_iname(knl: Kernel, split_iname: str, inner_length: int, *, outer_iname: Optional[str] = None,
                inner_iname: Optional[str] = None, inner_tag: Optional[str] = None,
                outer_tag: Optional[str] = None, **_ignored) -> Kernel:
    """for (i = lo; i < hi; i++) B  ->  for (i_outer ...) for (i_inner < inner_length) { i = lo + i_outer*len + i_inner; if (i < hi) B }."""
    if not isinstance(knl, Kernel):
        raise KernelError("split_iname: first argument must be a kernel")
    inner_length = int(inner_length)
    if inner_length <= 0:
        raise KernelError("split_iname: inner_length must be positive")
    if split_iname not in knl.inames():
        raise KernelError(f"split_iname: iname {split_iname!r} not found (have {sorted(knl.inames())})")
    outer = outer_iname or f"{split_iname}_outer"
    inner = inner_iname or f"{split_iname}_inner"
    taken = knl.all_names()
    for nm in (outer, inner):
        if nm in taken:
            raise KernelError(f"split_iname: name {nm!r} is already in use")
    for tag in (inner_tag, outer_tag):
        if tag is not None and not VALID_TAG.match(tag):
            raise KernelError(f"split_iname: invalid tag {tag!r}")
    new = knl.copy()

    def rewrite(nodes: List[c.Node]) -> List[c.Node]:
        out = []
        for n in nodes:
            if isinstance(n, c.For):
                body = rewrite(n.body)
                if n.var == split_iname:
                    extent = n.hi if _is_zero(n.lo) else c.BinOp("-", n.hi, n.lo)
                    flat = c.BinOp("+", c.BinOp("*", c.Name(outer), c.Num(str(inner_length))), c.Name(inner))
                    value = flat if _is_zero(n.lo) else c.BinOp("+", n.lo, flat)
                    bind = c.Bind(n.var, n.vtype, value, n.hi, body)
                    inner_loop = c.For(inner, n.vtype, c.Num("0"), c.Num(str(inner_length)), [bind], inner_tag)
                    out.append(c.For(outer, n.vtype, c.Num("0"), _ceil_div(extent, inner_length), [inner_loop], outer_tag))
                else:
                    out.append(c.For(n.var, n.vtype, n.lo, n.hi, body, n.tag))
            elif isinstance(n, c.Bind):
                out.append(c.Bind(n.var, n.vtype, n.value, n.hi, rewrite(n.body)))
            elif isinstance(n, c.If):
                out.append(c.If(n.cond, rewrite(n.then), rewrite(n.other)))
            else:
                out.append(n)
        return out

    new.func.body = rewrite(new.func.body)
    return new


def tag_inames(knl: Kernel, iname_to_tag, **_ignored) -> Kernel:
    """Accepts a dict, a list of (iname, tag) tuples or a "i:g.0,j:l.0" string; iname keys may be globs ("j*")."""
    if not isinstance(knl, Kernel):
        raise KernelError("tag_inames: first argument must be a kernel")
    if isinstance(iname_to_tag, str):
        pairs = [tuple(s.strip() for s in item.split(":")) for item in iname_to_tag.split(",") if item.strip()]
    elif isinstance(iname_to_tag, dict):
        pairs = list(iname_to_tag.items())
    else:
        pairs = [tuple(p) for p in iname_to_tag]
    new = knl.copy()
    names = new.inames()
    for pattern, tag in pairs:
        tag = str(tag)
        if not VALID_TAG.match(tag):
            raise KernelError(f"tag_inames: invalid tag {tag!r} for {pattern!r}")
        matched = [n for n in names if fnmatch.fnmatchcase(n, pattern)]
        if not matched:
            raise KernelError(f"tag_inames: no iname matches {pattern!r} (have {sorted(names)})")
        for l in new.loops():
            if l.var in matched:
                l.tag = tag
    return new


def fix_parameters(knl: Kernel, **params) -> Kernel:
    """Substitute JIT-time constants (NOMP_JIT arguments) and drop them from the signature
    (reference src/nomp.c:467-476, python/loopy_api.py:834-836)."""
    new = knl.copy()
    values = {k: v for k, v in params.items() if v is not None}

    def subst(e):
        if isinstance(e, c.Name) and e.id in values:
            v = values[e.id]
            return c.Num(repr(float(v)) if isinstance(v, float) else str(int(v)))
        return e

    new.func.body = map_stmts(new.func.body, subst)
    new.func.params = [p for p in new.func.params if p.name not in values]
    new.fixed.update(values)
    return new
