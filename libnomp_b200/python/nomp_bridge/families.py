"""Loop-family recogniser: decides which execution path a parsed kernel takes.

  native map      -> nompk_map()       hand-written sm_100a kernel in libnompk.so
  native reduce   -> nompk_reduce()
  native ax       -> nompk_ax_f64()
  map skeleton    -> NVRTC: any other single-loop elementwise kernel gets the same 128-bit grid-stride schedule
  reduce skeleton -> NVRTC: any other single-loop reduce clause gets the same single-pass reduction schedule
  generic         -> NVRTC: everything else, scheduled by the user's transform (emit_cuda.py)

The canonical spellings are the ones the reference tests use (reference tests/nomp-api-200-impl.h:36-40,
tests/nomp-api-600-impl.h:36-40, tests/nomp-api-500-impl.h:23-27, :47-51, :95-99, :180-184) plus the Ax kernel
string below (the reference has no Ax).  The result is a one-line descriptor, `//!nomp key=value ...`, that
travels to the backend as the first line of the `src` argument of knl_build() (reference
include/nomp-impl.h:213-237 keeps the vtable signatures), followed by CUDA source for the NVRTC paths.
"""
from __future__ import annotations

import re
from typing import Dict, List, Optional, Tuple

from . import cparse as c
from .emit_cuda import PRELUDE, cuda_type, expr_str, grid_expr_str, signature, const_int
from .ir import Kernel, KernelError, expr_names, map_expr, map_stmts, walk

# dtype / op codes of include/nompk.h
DTYPE_CODE = {("int", False): 0, ("int", True): 1, ("long", False): 2, ("long", True): 3,
              ("longlong", False): 2, ("longlong", True): 3, ("float", False): 4, ("double", False): 5}
MAP_OPS = {"add": 0, "sub": 1, "mul": 2, "axpy": 3, "xpay": 4, "axpby": 5, "scale": 6, "copy": 7, "fill": 8, "add3": 9}
RED_OPS = {"+": 0, "*": 1, "min": 2, "max": 3}

# ----------------------------------------------------------------------------------------------------
# The canonical Ax kernel string (layouts of include/nompk.h).  `n` is passed as NOMP_INT | NOMP_JIT.
# It is plain nomp C: the generic path can run it for any n, the oracle compiles it with gcc, and the
# recogniser maps it (up to renaming of identifiers) onto nompk_ax_f64 for the n that have a tuned kernel.
# ----------------------------------------------------------------------------------------------------
AX_KERNEL_SOURCE = """\
void nomp_ax(double *w, const double *u, const double *g, const double *D, int E, int n) {
  for (int e = 0; e < E; e++) {
    double ur[n][n][n];
    double us[n][n][n];
    double ut[n][n][n];
    for (int k = 0; k < n; k++)
      for (int j = 0; j < n; j++)
        for (int i = 0; i < n; i++) {
          double r = 0;
          double s = 0;
          double t = 0;
          for (int l = 0; l < n; l++) {
            r += D[i * n + l] * u[e * n * n * n + k * n * n + j * n + l];
            s += D[j * n + l] * u[e * n * n * n + k * n * n + l * n + i];
            t += D[k * n + l] * u[e * n * n * n + l * n * n + j * n + i];
          }
          ur[k][j][i] = g[(e * 6 + 0) * n * n * n + k * n * n + j * n + i] * r + g[(e * 6 + 1) * n * n * n + k * n * n + j * n + i] * s + g[(e * 6 + 2) * n * n * n + k * n * n + j * n + i] * t;
          us[k][j][i] = g[(e * 6 + 1) * n * n * n + k * n * n + j * n + i] * r + g[(e * 6 + 3) * n * n * n + k * n * n + j * n + i] * s + g[(e * 6 + 4) * n * n * n + k * n * n + j * n + i] * t;
          ut[k][j][i] = g[(e * 6 + 2) * n * n * n + k * n * n + j * n + i] * r + g[(e * 6 + 4) * n * n * n + k * n * n + j * n + i] * s + g[(e * 6 + 5) * n * n * n + k * n * n + j * n + i] * t;
        }
    for (int k = 0; k < n; k++)
      for (int j = 0; j < n; j++)
        for (int i = 0; i < n; i++) {
          double acc = 0;
          for (int l = 0; l < n; l++) {
            acc += D[l * n + i] * ur[k][j][l];
            acc += D[l * n + j] * us[k][l][i];
            acc += D[l * n + k] * ut[l][j][i];
          }
          w[e * n * n * n + k * n * n + j * n + i] = acc;
        }
  }
}
"""


# Ax fused with the dot product u . (A u) (p.Ap of a CG iteration): the same loop nest with one more statement and a
# reduce clause on `pap`.
AX_DOT_KERNEL_SOURCE = AX_KERNEL_SOURCE.replace(
    "void nomp_ax(double *w, const double *u, const double *g, const double *D, int E, int n) {",
    "void nomp_ax_dot(double *w, const double *u, const double *g, const double *D, int E, int n, double *pap) {").replace(
    "          w[e * n * n * n + k * n * n + j * n + i] = acc;\n",
    "          w[e * n * n * n + k * n * n + j * n + i] = acc;\n"
    "          pap[0] += u[e * n * n * n + k * n * n + j * n + i] * acc;\n")
assert "pap[0]" in AX_DOT_KERNEL_SOURCE and "nomp_ax_dot" in AX_DOT_KERNEL_SOURCE


# The first kernel of a CG iteration: the direction update p = r + beta p in front of the operator, then w = A p and p.Ap
# (nompk_ax_xpay_dot_peers_f64).  beta is a scalar argument, or -- second form -- a scalar in device memory read as beta[0].
_POINT = "e * n * n * n + k * n * n + j * n + i"
AX_XPAY_DOT_KERNEL_SOURCE = AX_DOT_KERNEL_SOURCE.replace(
    "void nomp_ax_dot(double *w, const double *u, const double *g, const double *D, int E, int n, double *pap) {",
    "void nomp_ax_xpay_dot(double *w, double *p, const double *res, const double *g, const double *D, double beta, int E, int n,"
    " double *pap) {").replace(
    "  for (int e = 0; e < E; e++) {\n",
    "  for (int e = 0; e < E; e++) {\n"
    "    for (int k = 0; k < n; k++)\n"
    "      for (int j = 0; j < n; j++)\n"
    "        for (int i = 0; i < n; i++)\n"
    f"          p[{_POINT}] = res[{_POINT}] + beta * p[{_POINT}];\n", 1).replace(" u[", " p[")
AX_XPAY_DOT_DEV_KERNEL_SOURCE = AX_XPAY_DOT_KERNEL_SOURCE.replace("double beta, int E", "const double *beta, int E").replace(
    "+ beta * p[", "+ beta[0] * p[")
assert AX_XPAY_DOT_KERNEL_SOURCE.count(" p[") == 6 and " u[" not in AX_XPAY_DOT_KERNEL_SOURCE and "beta[0]" in AX_XPAY_DOT_DEV_KERNEL_SOURCE


def _canonical_tokens(src: str) -> Tuple[List[str], List[str]]:
    """Token list with identifiers renamed v0, v1, ... in order of first appearance (keywords/types kept)."""
    keep = set(c._TYPE_WORDS) | {"for", "if", "else", "break", "continue"}
    names: List[str] = []
    out = []
    for kind, text, _ in c.tokenize(src):
        if kind == "id" and text not in keep:
            if text not in names:
                names.append(text)
            out.append(f"v{names.index(text)}")
        elif kind != "eof":
            out.append(text)
    return out, names


_AX_TOKENS, _AX_NAMES = _canonical_tokens(AX_KERNEL_SOURCE)


_AX_FUSED = None   # [(family, beta in device memory, tokens, names)] of the canonical strings that end in a p.Ap


def match_ax_dot(knl: Kernel) -> Optional[Dict[str, str]]:
    """Same for the strings fused with the dot product: Ax + dot, and (x)pay + Ax + dot in its two forms.  The roles carry
    "family" ("axdot" / "axxpaydot") and, for the latter, "beta_dev" ("0" / "1")."""
    global _AX_FUSED
    if _AX_FUSED is None:
        _AX_FUSED = [("axdot", None, *_canonical_tokens(AX_DOT_KERNEL_SOURCE)),
                     ("axxpaydot", "0", *_canonical_tokens(AX_XPAY_DOT_KERNEL_SOURCE)),
                     ("axxpaydot", "1", *_canonical_tokens(AX_XPAY_DOT_DEV_KERNEL_SOURCE))]
    try:
        toks, names = _canonical_tokens(knl.source)
    except Exception:
        return None
    for family, beta_dev, ref_toks, ref_names in _AX_FUSED:
        if toks == ref_toks and len(names) == len(ref_names):
            roles = dict(zip(ref_names, names))
            roles["family"] = family
            if beta_dev is not None:
                roles["beta_dev"] = beta_dev
                roles["u"] = roles["p"]
            return roles
    return None


def match_ax(knl: Kernel) -> Optional[Dict[str, str]]:
    """If the kernel text is the canonical Ax string up to renaming, return {role: user identifier}."""
    try:
        toks, names = _canonical_tokens(knl.source)
    except Exception:
        return None
    if toks != _AX_TOKENS or len(names) != len(_AX_NAMES):
        return None
    return dict(zip(_AX_NAMES, names))


def _ax_shaped(func: c.Function, n_arrays: int, n_ints: int) -> bool:
    """Cheap gate in front of the structural recogniser: n_arrays `double *` parameters and n_ints integer scalars,
    nothing else, and a loop nest at least four deep (elements, two point loops, a contraction)."""
    arrays = [p for p in func.params if p.is_array]
    ints = [p for p in func.params if not p.is_array and not p.ctype.is_float]
    if len(arrays) != n_arrays or len(ints) != n_ints or len(func.params) != n_arrays + n_ints:
        return False
    if any(p.ctype.base != "double" or p.ctype.ptr != 1 for p in arrays):
        return False

    def depth(nodes) -> int:
        return max([1 + depth(n.body) for n in nodes if isinstance(n, c.For)] or [0])

    return depth(func.body) >= 4


def may_be_ax_dot(func: c.Function, var: str) -> bool:
    """Before n is fixed: could this reduce clause be Ax fused with p.Ap (five double arrays, E and n)?"""
    return _ax_shaped(func, 5, 2) and var in {p.name for p in func.params}


def match_ax_structural(func: c.Function, fixed: Dict[str, object], supported, reduce_var: Optional[str]):
    """`func`: the untransformed loop nest with the NOMP_JIT arguments substituted.  (n, {role: parameter}) if it computes
    the Ax operator for one of the JIT-fixed integers n that has a tuned kernel (decided by axprobe.probe_ax); else None."""
    if not _ax_shaped(func, 5 if reduce_var else 4, 1):
        return None
    from .axprobe import probe_ax
    for value in fixed.values():
        try:
            n = int(value)
        except (TypeError, ValueError):
            continue
        if n != value or n not in supported:
            continue
        roles = probe_ax(func, n, reduce_var, fixed)
        if roles is not None:
            return n, roles
    return None


# ----------------------------------------------------------------------------------------------------
# helpers on the ORIGINAL (untransformed) function
# ----------------------------------------------------------------------------------------------------

def _dtype_code(t: c.CType) -> Optional[int]:
    return DTYPE_CODE.get((t.base, t.unsigned))


def _single_loop(func: c.Function) -> Optional[c.For]:
    body = [n for n in func.body if not (isinstance(n, c.If) and not n.then and not n.other)]
    if len(body) != 1 or not isinstance(body[0], c.For):
        return None
    loop = body[0]
    if any(isinstance(n, (c.For, c.Bind)) for n in walk(loop.body)):
        return None
    return loop


def _params(func: c.Function) -> Dict[str, c.Param]:
    return {p.name: p for p in func.params}


def _is_elem(e: c.Node, arrays: Dict[str, c.Param], var: str) -> Optional[str]:
    """`a[i]` with a a pointer parameter and i the loop variable -> 'a'."""
    if (isinstance(e, c.Subscript) and isinstance(e.base, c.Name) and e.base.id in arrays and len(e.index) == 1
            and isinstance(e.index[0], c.Name) and e.index[0].id == var):
        return e.base.id
    return None


def invariant_arrays(stmts, exprs, arrays: Dict[str, c.Param]) -> set:
    """Pointer parameters that the loop only READS, and only at constant subscripts (`alpha[0]`): scalars that live in
    device memory -- the result a reduce clause left there, a coefficient computed by an earlier kernel.  The skeletons
    keep such reads as they are (one cached load per thread) instead of giving up their vectorised schedule."""
    kinds: Dict[str, set] = {}
    written = set()

    def visit(e):
        if isinstance(e, c.Subscript) and isinstance(e.base, c.Name) and e.base.id in arrays:
            kinds.setdefault(e.base.id, set()).add(len(e.index) == 1 and const_int(e.index[0]) is not None)
        return e

    for n in walk(stmts):
        if isinstance(n, c.Assign):
            map_expr(n.target, visit)
            map_expr(n.value, visit)
            if isinstance(n.target, c.Subscript) and isinstance(n.target.base, c.Name):
                written.add(n.target.base.id)
        elif isinstance(n, c.Decl):
            map_expr(n.init, visit)
        elif isinstance(n, c.If):
            map_expr(n.cond, visit)
    for e in exprs:
        map_expr(e, visit)
    return {a for a, k in kinds.items() if k == {True} and a not in written}


def hoist_invariants(stmts: List[c.Node], exprs: List[c.Node], invariant: set, scalars: set):
    """Loop-invariant subexpressions of an elementwise loop body -- built from literals, scalar arguments and reads of
    device-resident scalars (`rr[0]`, arrays in `invariant`) -- that are EXPENSIVE (a division, a remainder, a call) are
    evaluated once per CTA instead of once per element: thread 0 computes them into shared memory, one barrier, every
    thread keeps them in registers.  `x[i] += (rr[0] / pap[0]) * p[i]` (the update of a CG iteration whose scalars stay on
    the device) otherwise pays an fp64 division per element and is no longer bandwidth-bound.  The barrier has a price
    of its own -- a CTA issues its streaming loads only after thread 0's dependent loads and division, about a
    microsecond -- so the callers give such kernels four tiles per CTA, and plain reads like `alpha[0]` (one cached load
    per thread) are left where they are.  The subtree is moved as a whole: every rounding stays where the C text has it.
    Returns (statements, expressions, CUDA prologue text) with the hoisted subtrees replaced by `nomp_inv_<k>`."""
    table: Dict[str, int] = {}

    def inv(e) -> bool:
        if isinstance(e, c.Num):
            return True
        if isinstance(e, c.Name):
            return e.id in scalars
        if isinstance(e, c.Subscript):
            return isinstance(e.base, c.Name) and e.base.id in invariant and all(const_int(i) is not None for i in e.index)
        if isinstance(e, c.BinOp):
            return inv(e.left) and inv(e.right)
        if isinstance(e, (c.UnOp, c.Cast)):
            return inv(e.operand)
        if isinstance(e, c.Ternary):
            return inv(e.cond) and inv(e.then) and inv(e.other)
        if isinstance(e, c.Call):
            return all(inv(a) for a in e.args)
        return False

    def worth(e) -> bool:
        hit = [False]

        def visit(x):
            if isinstance(x, c.Call) or (isinstance(x, c.BinOp) and x.op in ("/", "%")):
                hit[0] = True
            return x
        map_expr(e, visit)
        return hit[0]

    def rewrite(e):
        if e is None or isinstance(e, (c.Num, c.Name)):
            return e
        if inv(e) and worth(e):
            return c.Name(f"nomp_inv_{table.setdefault(expr_str(e), len(table))}")
        if isinstance(e, c.BinOp):
            return c.BinOp(e.op, rewrite(e.left), rewrite(e.right))
        if isinstance(e, c.UnOp):
            return c.UnOp(e.op, rewrite(e.operand))
        if isinstance(e, c.Cast):
            return c.Cast(e.ctype, rewrite(e.operand))
        if isinstance(e, c.Ternary):
            return c.Ternary(rewrite(e.cond), rewrite(e.then), rewrite(e.other))
        if isinstance(e, c.Call):
            return c.Call(e.func, [rewrite(a) for a in e.args])
        if isinstance(e, c.Subscript):
            return c.Subscript(e.base, [rewrite(i) for i in e.index])
        return e

    def rewrite_stmts(nodes):
        out = []
        for n in nodes:
            if isinstance(n, c.Assign):
                out.append(c.Assign(rewrite(n.target), n.op, rewrite(n.value)))
            elif isinstance(n, c.Decl):
                out.append(c.Decl(n.ctype, n.name, list(n.dims), rewrite(n.init)))
            elif isinstance(n, c.If):
                out.append(c.If(rewrite(n.cond), rewrite_stmts(n.then), rewrite_stmts(n.other)))
            else:
                out.append(n)
        return out

    new_stmts = rewrite_stmts(stmts)
    new_exprs = [rewrite(e) for e in exprs]
    if not table:
        return stmts, exprs, ""
    texts = sorted(table, key=table.get)
    lines = [f"  __shared__ decltype(+({t})) nomp_inv_s{k};" for k, t in enumerate(texts)]
    lines.append("  if (threadIdx.x == 0) { " + " ".join(f"nomp_inv_s{k} = ({t});" for k, t in enumerate(texts)) + " }")
    lines.append("  __syncthreads();")
    lines += [f"  const decltype(+({t})) nomp_inv_{k} = nomp_inv_s{k};" for k, t in enumerate(texts)]
    return new_stmts, new_exprs, "\n".join(lines) + "\n"


def _terms(e: c.Node) -> Optional[List[List[c.Node]]]:
    """Flatten e into a sum of products (no parenthesised sums inside products). None if not of that shape."""
    if isinstance(e, c.BinOp) and e.op == "+":
        a, b = _terms(e.left), _terms(e.right)
        return None if a is None or b is None else a + b
    factors: List[c.Node] = []

    def prod(x) -> bool:
        if isinstance(x, c.BinOp) and x.op == "*":
            return prod(x.left) and prod(x.right)
        if isinstance(x, (c.Name, c.Subscript, c.Num)):
            factors.append(x)
            return True
        return False

    return [factors] if prod(e) else None


def _loop_count(loop: c.For, params: Dict[str, c.Param]) -> Optional[str]:
    """Trip count as 'name' (int parameter) or a literal, for loops that start at 0."""
    if const_int(loop.lo) != 0:
        return None
    if isinstance(loop.hi, c.Name) and loop.hi.id in params and not params[loop.hi.id].is_array \
            and not params[loop.hi.id].ctype.is_float:
        return loop.hi.id
    v = const_int(loop.hi)
    return str(v) if v is not None and v >= 0 else None


# ----------------------------------------------------------------------------------------------------
# native map
# ----------------------------------------------------------------------------------------------------

def match_native_map(func: c.Function) -> Optional[Dict[str, str]]:
    loop = _single_loop(func)
    if loop is None or len(loop.body) != 1 or not isinstance(loop.body[0], c.Assign):
        return None
    params = _params(func)
    count = _loop_count(loop, params)
    if count is None:
        return None
    arrays = {k: p for k, p in params.items() if p.is_array}
    st: c.Assign = loop.body[0]
    y = _is_elem(st.target, arrays, loop.var)
    if y is None or arrays[y].ctype.const:
        return None
    dt = _dtype_code(arrays[y].ctype)
    if dt is None:
        return None
    ytype = arrays[y].ctype

    def same_type(p: c.Param) -> bool:
        return (p.ctype.base, p.ctype.unsigned) == (ytype.base, ytype.unsigned) or \
               ({p.ctype.base, ytype.base} <= {"long", "longlong"} and p.ctype.unsigned == ytype.unsigned)

    value = st.value
    if st.op in ("+=", "-=", "*="):
        value = c.BinOp(st.op[0], st.target, st.value)
    elif st.op != "=":
        return None

    def classify(f: c.Node):
        """('y',) | ('x', name) | ('s', name) for y[i], other array element, scalar parameter."""
        a = _is_elem(f, arrays, loop.var)
        if a is not None:
            if not same_type(arrays[a]):
                return None
            return ("y",) if a == y else ("x", a)
        if isinstance(f, c.Name) and f.id in params and not params[f.id].is_array and same_type(params[f.id]):
            return ("s", f.id)
        return None

    out = {"family": "map", "dtype": str(dt), "y": y, "n": count}

    # y - x
    if isinstance(value, c.BinOp) and value.op == "-":
        l, r = classify(value.left), classify(value.right)
        if l == ("y",) and r and r[0] == "x":
            return {**out, "op": "sub", "x": r[1]}
        return None
    terms = _terms(value)
    if terms is None:
        return None
    cls = []
    for t in terms:
        k = [classify(f) for f in t]
        if any(x is None for x in k):
            return None
        cls.append(sorted(k))
    cls.sort()
    n_terms = len(cls)
    if n_terms == 1:
        t = cls[0]
        if t == [("y",)]:
            return None
        if len(t) == 1 and t[0][0] == "x":
            return {**out, "op": "copy", "x": t[0][1]}
        if len(t) == 1 and t[0][0] == "s":
            return {**out, "op": "fill", "alpha": t[0][1]}
        if len(t) == 2 and t[0][0] == "s" and t[1] == ("y",):
            return {**out, "op": "scale", "alpha": t[0][1]}
        if len(t) == 2 and t[0][0] == "x" and t[1] == ("y",):
            return {**out, "op": "mul", "x": t[0][1]}
        return None
    if n_terms == 2:
        a, b = cls
        flat = (tuple(a), tuple(b))
        # y + x
        if len(a) == 1 and len(b) == 1:
            kinds = sorted([a[0], b[0]])
            if kinds[0][0] == "x" and kinds[1] == ("y",):
                return {**out, "op": "add", "x": kinds[0][1]}
            if kinds[0][0] == "x" and kinds[1][0] == "x" and kinds[0][1] != kinds[1][1]:
                # order of the two reads does not matter for a two-term sum
                return {**out, "op": "add3", "x": kinds[0][1], "z": kinds[1][1]}
            return None
        one = [t for t in (a, b) if len(t) == 1]
        two = [t for t in (a, b) if len(t) == 2]
        if len(one) == 1 and len(two) == 1:
            s_x = two[0]
            # alpha * x  (+ y)   -> axpy ;  alpha * y (+ x) -> xpay
            if s_x[0][0] == "s" and s_x[1][0] == "x" and one[0] == [("y",)]:
                return {**out, "op": "axpy", "alpha": s_x[0][1], "x": s_x[1][1]}
            if s_x[0][0] == "s" and s_x[1] == ("y",) and one[0][0][0] == "x":
                return {**out, "op": "xpay", "alpha": s_x[0][1], "x": one[0][0][1]}
            return None
        if len(two) == 2:
            # alpha * x + beta * y
            tx = [t for t in two if t[0][0] == "s" and t[1][0] == "x"]
            ty = [t for t in two if t[0][0] == "s" and t[1] == ("y",)]
            if len(tx) == 1 and len(ty) == 1:
                return {**out, "op": "axpby", "alpha": tx[0][0][1], "x": tx[0][1][1], "beta": ty[0][0][1]}
        del flat
    return None


# ----------------------------------------------------------------------------------------------------
# reductions: nomp_bridge/reduction.py (re-exported here for the callers that import them from families)
# ----------------------------------------------------------------------------------------------------
from .reduction import (ReductionInfo, analyse_reduction, emit_reduce_skeleton, match_native_reduce)  # noqa: E402,F401


def match_map_skeleton(func: c.Function) -> Optional[c.For]:
    """Single loop whose every array access is `a[i]` on a pointer parameter, no local arrays, all arrays of one
    element size: eligible for the vectorised grid-stride schedule."""
    loop = _single_loop(func)
    if loop is None:
        return None
    params = _params(func)
    arrays = {k: p for k, p in params.items() if p.is_array}
    sizes = set()
    ok = True
    written_scalars = set()
    invariant = invariant_arrays(loop.body, [], arrays)

    def check_expr(e):
        nonlocal ok
        if isinstance(e, c.Subscript):
            if isinstance(e.base, c.Name) and e.base.id in invariant:
                return e
            a = _is_elem(e, arrays, loop.var)
            if a is None:
                ok = False
            else:
                sizes.add(arrays[a].ctype.size)
        return e

    for n in walk(loop.body):
        if isinstance(n, (c.Break, c.Continue)):
            return None
        if isinstance(n, c.Decl):
            if n.dims:
                return None
            map_expr(n.init, check_expr)
        elif isinstance(n, c.Assign):
            map_expr(n.target, check_expr)
            map_expr(n.value, check_expr)
            if isinstance(n.target, c.Name):
                written_scalars.add(n.target.id)
            elif isinstance(n.target, c.Subscript) and isinstance(n.target.base, c.Name) \
                    and n.target.base.id in arrays and arrays[n.target.base.id].ctype.const:
                return None
        elif isinstance(n, c.If):
            map_expr(n.cond, check_expr)
    for b in (loop.lo, loop.hi):
        if any(isinstance(x, c.Subscript) for x in _subexprs(b)):
            return None
    if not ok or len(sizes) != 1 or (written_scalars & set(params)) or sizes.pop() not in (4, 8):
        return None
    if loop.var in written_scalars:
        return None
    return loop


def _subexprs(e):
    acc = []
    map_expr(e, lambda x: (acc.append(x), x)[1])
    return acc


def emit_map_skeleton(knl: Kernel, loop: c.For, sm_count: int) -> Tuple[str, List[str], List[str]]:
    """Vectorised elementwise kernel: 128-bit loads/stores when every operand is 16-byte aligned (checked in the
    kernel, the branch is uniform), scalar grid-stride otherwise; same schedule as libnompk map.cu."""
    func = knl.func
    params = _params(func)
    arrays = {k: p for k, p in params.items() if p.is_array}
    invariant = invariant_arrays(loop.body, [], arrays)
    arrays = {k: p for k, p in arrays.items() if k not in invariant}    # `alpha[0]` stays a plain read
    used, written = [], []

    def note(e):
        if isinstance(e, c.Subscript) and isinstance(e.base, c.Name) and e.base.id in arrays and e.base.id not in used:
            used.append(e.base.id)
        return e

    for n in walk(loop.body):
        if isinstance(n, c.Assign):
            map_expr(n.value, note)
            map_expr(n.target, note)
            if isinstance(n.target, c.Subscript) and n.target.base.id not in written:
                written.append(n.target.base.id)
        elif isinstance(n, c.Decl):
            map_expr(n.init, note)
        elif isinstance(n, c.If):
            map_expr(n.cond, note)
    esize = arrays[used[0]].ctype.size if used else 8
    lanes = 16 // esize

    def scalarise(e):
        if isinstance(e, c.Subscript) and isinstance(e.base, c.Name) and e.base.id in arrays:
            return c.Name(f"nomp_{e.base.id}_i")
        return e

    written_names = {n.target.id for n in walk(loop.body) if isinstance(n, c.Assign) and isinstance(n.target, c.Name)}
    declared = {n.name for n in walk(loop.body) if isinstance(n, c.Decl)}
    scalars = {k for k, prm in params.items() if not prm.is_array} - written_names - declared - {loop.var}
    hoisted, _, prologue = hoist_invariants(loop.body, [], invariant, scalars)
    body_nodes = map_stmts(hoisted, scalarise)
    from .emit_cuda import GenericEmitter
    ge = GenericEmitter(knl)
    ge.lines = []
    ge.stmts(body_nodes, 3, True, False)
    body_scalar = "\n".join(ge.lines)
    it = cuda_type(loop.vtype)
    lo, hi = expr_str(loop.lo), expr_str(loop.hi)
    ety = {a: cuda_type(arrays[a].ctype) for a in used}
    align = " | ".join(f"(nomp_u64_t)({a} + nomp_lo)" for a in used) or "0"
    # U vectors per thread and trip: with hoisted invariants a CTA takes four tiles (below), and the loads of all four
    # are issued before the first store -- the pointers are not __restrict__, so the compiler would not move them there
    U = 4 if prologue else 1
    vec_decl = "\n".join(f"      struct __align__(16) {{ {ety[a]} v[{lanes}]; }} nomp_{a}_v[{U}];" for a in used)
    vec_load = "\n".join(
        f"        *reinterpret_cast<int4 *>(&nomp_{a}_v[nomp_q]) = *reinterpret_cast<const int4 *>({a} + nomp_lo + nomp_v * {lanes});"
        for a in used)
    vec_store = "\n".join(
        f"        *reinterpret_cast<int4 *>({a} + nomp_e) = *reinterpret_cast<int4 *>(&nomp_{a}_v[nomp_q]);" for a in written)
    lane_in = "\n".join(f"          {ety[a]} nomp_{a}_i = nomp_{a}_v[nomp_q].v[nomp_l];" for a in used)
    lane_out = "\n".join(f"          nomp_{a}_v[nomp_q].v[nomp_l] = nomp_{a}_i;" for a in written)
    sc_in = "\n".join(f"      {ety[a]} nomp_{a}_i = {a}[nomp_e];" for a in used)
    sc_out = "\n".join(f"      {a}[nomp_e] = nomp_{a}_i;" for a in written)
    ge.lines = []
    ge.stmts(body_nodes, 5, True, False)
    body = "\n".join(ge.lines)
    src = f"""{PRELUDE}
// elementwise loop over `{loop.var}`: vectorised grid-stride schedule of libnompk map.cu with a generated body
{signature(knl)} {{
  const long long nomp_lo = (long long)({lo}), nomp_hi = (long long)({hi});
  const long long nomp_n = nomp_hi - nomp_lo;
  if (nomp_n <= 0) return;
{prologue}  const long long nomp_tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long nomp_nthreads = (long long)gridDim.x * blockDim.x;
  if ((({align}) & 15u) == 0) {{
    const long long nomp_nvec = nomp_n / {lanes};
    for (long long nomp_v0 = (long long)blockIdx.x * (blockDim.x * {U}) + threadIdx.x; nomp_v0 < nomp_nvec; nomp_v0 += nomp_nthreads * {U}) {{
{vec_decl}
#pragma unroll
      for (int nomp_q = 0; nomp_q < {U}; nomp_q++) {{
        const long long nomp_v = nomp_v0 + (long long)nomp_q * blockDim.x;
        if (nomp_v >= nomp_nvec) break;
{vec_load}
      }}
#pragma unroll
      for (int nomp_q = 0; nomp_q < {U}; nomp_q++) {{
        const long long nomp_v = nomp_v0 + (long long)nomp_q * blockDim.x;
        if (nomp_v >= nomp_nvec) break;
        const long long nomp_e = nomp_lo + nomp_v * {lanes};
#pragma unroll
        for (int nomp_l = 0; nomp_l < {lanes}; nomp_l++) {{
          const {it} {loop.var} = ({it})(nomp_e + nomp_l);
{lane_in}
{body}
{lane_out}
        }}
{vec_store}
      }}
    }}
    for (long long nomp_e = nomp_lo + nomp_nvec * {lanes} + nomp_tid; nomp_e < nomp_hi; nomp_e += nomp_nthreads) {{
      const {it} {loop.var} = ({it})nomp_e;
{sc_in}
{body_scalar}
{sc_out}
    }}
  }} else {{
    for (long long nomp_e = nomp_lo + nomp_tid; nomp_e < nomp_hi; nomp_e += nomp_nthreads) {{
      const {it} {loop.var} = ({it})nomp_e;
{sc_in}
{body_scalar}
{sc_out}
    }}
  }}
}}
"""
    int_params = {p.name for p in func.params if not p.is_array and not p.ctype.is_float}
    extent = c.BinOp("-", loop.hi, loop.lo) if const_int(loop.lo) != 0 else loop.hi
    ext = grid_expr_str(extent, int_params)
    # one vector per thread, one tile per CTA: on B200 the hardware CTA scheduler streams faster than a persistent
    # grid-stride loop (tools/exp/exp_map.cu); the loops in the kernel still stride, so any grid size is correct
    # (four tiles per CTA when the CTA starts with the barrier of hoisted invariants: hoist_invariants)
    per_block = 256 * lanes * (4 if prologue else 1)
    grid = f"max(1, ({ext} + {per_block - 1}) / {per_block})"
    return src, [grid, "1", "1"], ["256", "1", "1"]
