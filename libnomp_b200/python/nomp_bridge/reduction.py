"""The reduce clause (role of the reference's python/reduction.py).

The reference rewrites `var[0] += expr` into one partial per 512-thread work group with loopy
(reference python/reduction.py:30-134: exactly one loop and one callable :37-46, the assignment whose lhs is
`var[...]`, `expression.children = (lhs, *rhs)` with the lhs DROPPED :68, Sum or Product :71-82) and leaves the final
fold to the host (reference src/reduction.c).  Here the clause is analysed on our own loop-nest IR and dispatched to
  - libnompk's hand-written single-pass reduction (sum / dot / min / max / prod of plain arrays), or
  - a generated kernel with the same single-pass schedule for any other right-hand side, conditions and elementwise
    updates in front of the accumulation (the "reduce skeleton").
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

from . import cparse as c
from .emit_cuda import PRELUDE, const_int, cuda_type, expr_str, grid_expr_str
from .ir import Kernel, KernelError, expr_names, walk

RED_OPS = {"+": 0, "*": 1, "min": 2, "max": 3}


def _fam():
    from . import families          # late import: families re-exports this module's names
    return families


def _params(func):
    return _fam()._params(func)


def _is_elem(e, arrays, var):
    return _fam()._is_elem(e, arrays, var)


def _loop_count(loop, params):
    return _fam()._loop_count(loop, params)


def _dtype_code(t):
    return _fam()._dtype_code(t)



class ReductionInfo:
    def __init__(self, loop: c.For, var: str, vtype: c.CType, op: str, rhs: c.Node, preds: List[c.Node],
                 pre: List[c.Node]):
        self.loop, self.var, self.vtype, self.op, self.rhs, self.preds, self.pre = loop, var, vtype, op, rhs, preds, pre


def analyse_reduction(func: c.Function, var: str, op: str) -> ReductionInfo:
    """Find `var[0] op= rhs` inside the single loop (reference python/reduction.py:37-48 requires one loop and a
    subscripted accumulator; the lhs' incoming value is dropped, :68)."""
    params = _params(func)
    if var not in params or not params[var].is_array:
        raise KernelError(f"reduce: {var!r} must be a pointer parameter of the kernel")
    body = [n for n in func.body if not (isinstance(n, c.If) and not n.then and not n.other)]
    if len(body) != 1 or not isinstance(body[0], c.For):
        raise KernelError("reduce: the kernel must consist of exactly one loop")
    loop = body[0]
    if any(isinstance(n, (c.For, c.Bind)) for n in walk(loop.body)):
        raise KernelError("reduce: nested loops are not supported in a reduction kernel")
    vtype = params[var].ctype.scalar()

    def is_acc(e: c.Node) -> bool:
        return (isinstance(e, c.Subscript) and isinstance(e.base, c.Name) and e.base.id == var
                and len(e.index) == 1 and const_int(e.index[0]) == 0)

    found: List[Tuple[c.Assign, List[c.Node]]] = []
    pre: List[c.Node] = []

    def scan(nodes, preds, top):
        for n in nodes:
            if isinstance(n, c.Assign) and is_acc(n.target):
                found.append((n, list(preds)))
            elif isinstance(n, c.If):
                scan(n.then, preds + [n.cond], False)
                scan(n.other, preds + [c.UnOp("!", n.cond)], False)
            elif top:
                pre.append(n)
            else:
                raise KernelError("reduce: only the accumulation may appear under a condition")

    scan(loop.body, [], True)
    if len(found) != 1:
        raise KernelError(f"reduce: expected exactly one update of {var}[0], found {len(found)}")
    st, preds = found[0]
    uses_acc = lambda e: var in expr_names(e)  # noqa: E731
    binop = {"+": "+", "*": "*"}.get(op)
    rhs = None
    if op in ("+", "*"):
        if st.op == f"{op}=":
            rhs = st.value
        elif st.op == "=" and isinstance(st.value, c.BinOp) and st.value.op == binop:
            if is_acc(st.value.left):
                rhs = st.value.right
            elif is_acc(st.value.right):
                rhs = st.value.left
    else:  # min / max:  m[0] = (x < m[0]) ? x : m[0]   in any arrangement
        if st.op == "=" and isinstance(st.value, c.Ternary):
            cand = [b for b in (st.value.then, st.value.other) if not is_acc(b)]
            if len(cand) == 1:
                rhs = cand[0]
    if rhs is None or uses_acc(rhs):
        raise KernelError(f"reduce: the update of {var}[0] does not have the form of a '{op}' reduction")
    # Statements in front of the accumulation may update other arrays, but only elementwise (index == loop variable):
    # iteration i is then owned by exactly one thread and fusing the update with the reduction is safe
    # (e.g. the CG update  x[i] += a*p[i]; r[i] -= a*w[i]; rr[0] += r[i]*r[i];).
    arrays = {k for k, p in params.items() if p.is_array}
    for n in walk(pre):
        if isinstance(n, c.Assign):
            if isinstance(n.target, c.Name) and n.target.id in params:
                raise KernelError("reduce: a reduction kernel may not assign to its scalar arguments")
            if isinstance(n.target, c.Subscript) and isinstance(n.target.base, c.Name) and n.target.base.id in arrays:
                if _is_elem(n.target, {k: params[k] for k in arrays}, loop.var) is None:
                    raise KernelError("reduce: other arguments may only be written elementwise (a[i]) in a reduction kernel")
    return ReductionInfo(loop, var, vtype, op, rhs, preds, pre)


def match_native_reduce(func: c.Function, info: ReductionInfo) -> Optional[Dict[str, str]]:
    if info.preds or info.pre:
        return None
    params = _params(func)
    count = _loop_count(info.loop, params)
    dt = _dtype_code(info.vtype)
    if count is None or dt is None:
        return None
    arrays = {k: p for k, p in params.items() if p.is_array and k != info.var}

    def elem(e):
        a = _is_elem(e, arrays, info.loop.var)
        if a is None:
            return None
        t = arrays[a].ctype
        if (t.base, t.unsigned) != (info.vtype.base, info.vtype.unsigned):
            return None
        return a

    out = {"family": "reduce", "dtype": str(dt), "op": str(RED_OPS[info.op]), "n": count, "out": info.var}
    x = elem(info.rhs)
    if x is not None:
        return {**out, "x": x}
    if isinstance(info.rhs, c.BinOp) and info.rhs.op == "*":
        x, y = elem(info.rhs.left), elem(info.rhs.right)
        if x is not None and y is not None:
            return {**out, "x": x, "y": y}
    return None


# ---- generated single-pass reduction kernel -------------------------------------------------------------------
_RED_IDENTITY = {"+": "0", "*": "1"}


def _limits(t: c.CType, hi: bool) -> str:
    if t.is_float:
        return ("" if hi else "-") + ("__int_as_float(0x7f800000)" if t.base == "float" else "__longlong_as_double(0x7ff0000000000000LL)")
    bits = t.size * 8
    if t.unsigned:
        return f"({cuda_type(t)})~({cuda_type(t)})0" if hi else "0"
    if bits == 32:
        return "2147483647" if hi else "(-2147483647 - 1)"
    return "9223372036854775807LL" if hi else "(-9223372036854775807LL - 1)"


def emit_reduce_skeleton(knl: Kernel, info: ReductionInfo, sm_count: int) -> Tuple[str, List[str], List[str], List[str]]:
    """Single-pass reduction kernel with the same structure as libnompk's reduce.cu, for an arbitrary rhs.
    Returns (source, grid exprs, block exprs, kernel parameter names).  The trailing five parameters
    (partials, ticket, result, result_host, seq) are supplied by the backend."""
    T = cuda_type(info.vtype)
    func = knl.func
    params = [p for p in func.params if p.name != info.var]
    written = set()
    for node in walk(info.pre):
        if isinstance(node, c.Assign) and isinstance(node.target, c.Subscript) and isinstance(node.target.base, c.Name):
            written.add(node.target.base.id)
    sig_parts = []
    for prm in params:
        t = prm.ctype
        if prm.is_array:
            sig_parts.append(f"{cuda_type(t)} *{prm.name}" if prm.name in written else f"const {cuda_type(t)} *__restrict__ {prm.name}")
        else:
            sig_parts.append(f"{cuda_type(t)} {prm.name}")
    sig_parts += [f"{T} *__restrict__ nomp_partials", "unsigned int *__restrict__ nomp_ticket",
                  f"{T} *__restrict__ nomp_result", f"{T} *__restrict__ nomp_result_host", "unsigned long long nomp_seq"]
    int_params = {p.name for p in params if not p.is_array and not p.ctype.is_float}
    it = cuda_type(info.loop.vtype)
    lo, hi = expr_str(info.loop.lo), expr_str(info.loop.hi)
    if info.op in _RED_IDENTITY:
        ident = f"({T}){_RED_IDENTITY[info.op]}"
        comb = lambda a, b: f"({a}) {info.op} ({b})"  # noqa: E731
    elif info.op == "min":
        ident = _limits(info.vtype, True)
        comb = lambda a, b: f"(({b}) < ({a}) ? ({b}) : ({a}))"  # noqa: E731
    else:
        ident = _limits(info.vtype, False)
        comb = lambda a, b: f"(({b}) > ({a}) ? ({b}) : ({a}))"  # noqa: E731
    pre_lines = []
    from .emit_cuda import GenericEmitter
    ge = GenericEmitter(knl)
    ge.lines = []
    ge.stmts(info.pre, 2, True, False)
    pre_lines = ge.lines
    cond = " && ".join(expr_str(p) for p in info.preds)
    rhs = f"({T})({expr_str(info.rhs)})"
    upd = f"nomp_acc = {comb('nomp_acc', 'nomp_v')};"
    body = [f"const {T} nomp_v = {rhs};", upd]
    if cond:
        body = [f"if ({cond}) {{"] + ["  " + b for b in body] + ["}"]
    shfl = f"nomp_o = __shfl_xor_sync(0xffffffffu, nomp_acc, nomp_s); nomp_acc = {comb('nomp_acc', 'nomp_o')};"
    src = f"""{PRELUDE}
// reduce clause on `{info.var}` (op {info.op}): single-pass schedule of libnompk reduce.cu with a generated right-hand side
extern "C" __global__ void __launch_bounds__(256) {knl.name}({', '.join(sig_parts)}) {{
  {T} nomp_acc = {ident};
  {T} nomp_o;
  const long long nomp_lo = (long long)({lo}), nomp_hi = (long long)({hi});
  for (long long nomp_i = nomp_lo + (long long)blockIdx.x * 256 + threadIdx.x; nomp_i < nomp_hi;
       nomp_i += (long long)gridDim.x * 256) {{
    const {it} {info.loop.var} = ({it})nomp_i;
{chr(10).join(pre_lines)}
    {(chr(10) + '    ').join(body)}
  }}
  __shared__ {T} nomp_warp[8];
  __shared__ bool nomp_last;
  for (int nomp_s = 16; nomp_s > 0; nomp_s >>= 1) {{ {shfl} }}
  if ((threadIdx.x & 31) == 0) nomp_warp[threadIdx.x >> 5] = nomp_acc;
  __syncthreads();
  if (threadIdx.x < 32) {{
    nomp_acc = threadIdx.x < 8 ? nomp_warp[threadIdx.x] : {ident};
    for (int nomp_s = 16; nomp_s > 0; nomp_s >>= 1) {{ {shfl} }}
  }}
  if (threadIdx.x == 0) {{
    nomp_partials[blockIdx.x] = nomp_acc;
    __threadfence();
    nomp_last = (atomicAdd(nomp_ticket, 1u) == gridDim.x - 1);
  }}
  __syncthreads();
  if (!nomp_last) return;
  __threadfence();
  nomp_acc = {ident};
  for (unsigned int nomp_b = threadIdx.x; nomp_b < gridDim.x; nomp_b += 256) {{
    nomp_o = __ldcg(nomp_partials + nomp_b);
    nomp_acc = {comb('nomp_acc', 'nomp_o')};
  }}
  for (int nomp_s = 16; nomp_s > 0; nomp_s >>= 1) {{ {shfl} }}
  __syncthreads();
  if ((threadIdx.x & 31) == 0) nomp_warp[threadIdx.x >> 5] = nomp_acc;
  __syncthreads();
  if (threadIdx.x < 32) {{
    nomp_acc = threadIdx.x < 8 ? nomp_warp[threadIdx.x] : {ident};
    for (int nomp_s = 16; nomp_s > 0; nomp_s >>= 1) {{ {shfl} }}
    if (threadIdx.x == 0) {{
      *nomp_result = nomp_acc;
      if (nomp_result_host) {{
        *(volatile {T} *)nomp_result_host = nomp_acc;
        __threadfence_system();
        *(volatile unsigned long long *)((char *)nomp_result_host + 8) = nomp_seq;
      }}
      *nomp_ticket = 0u;
    }}
  }}
}}
"""
    extent = c.BinOp("-", info.loop.hi, info.loop.lo) if const_int(info.loop.lo) != 0 else info.loop.hi
    try:
        ext = grid_expr_str(extent, int_params)
        grid = f"max(1, min(({ext} + 255) / 256, {max(1, sm_count) * 8}))"
    except KernelError:
        grid = str(max(1, sm_count) * 8)  # data-dependent bounds: a full grid, the loop guards itself
    names = [p.name for p in params] + ["nomp_partials", "nomp_ticket", "nomp_result", "nomp_result_host", "nomp_seq"]
    return src, [grid, "1", "1"], ["256", "1", "1"], names


