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
    """`pre` / `post`: the top-level statements of the loop body in front of / behind the accumulation, in source order
    (the reference keeps program order: seq_dependencies=True, reference python/loopy_api.py:817)."""

    def __init__(self, loop: c.For, var: str, vtype: c.CType, op: str, rhs: c.Node, preds: List[c.Node],
                 pre: List[c.Node], post: Optional[List[c.Node]] = None):
        self.loop, self.var, self.vtype, self.op, self.rhs, self.preds, self.pre = loop, var, vtype, op, rhs, preds, pre
        self.post = post or []


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
    post: List[c.Node] = []

    def scan(nodes, preds, top):
        for n in nodes:
            if isinstance(n, c.Assign) and is_acc(n.target):
                found.append((n, list(preds)))
            elif isinstance(n, c.If):
                scan(n.then, preds + [n.cond], False)
                scan(n.other, preds + [c.UnOp("!", n.cond)], False)
            elif top:
                (post if found else pre).append(n)     # a statement behind the accumulation stays behind it
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
    for n in walk(pre + post):
        if isinstance(n, c.Assign):
            if isinstance(n.target, c.Name) and n.target.id in params:
                raise KernelError("reduce: a reduction kernel may not assign to its scalar arguments")
            if isinstance(n.target, c.Subscript) and isinstance(n.target.base, c.Name) and n.target.base.id in arrays:
                if _is_elem(n.target, {k: params[k] for k in arrays}, loop.var) is None:
                    raise KernelError("reduce: other arguments may only be written elementwise (a[i]) in a reduction kernel")
    if any(var in expr_names(e) for n in walk(pre + post) for e in _stmt_exprs(n)):
        raise KernelError(f"reduce: {var} may only appear in its own update")
    return ReductionInfo(loop, var, vtype, op, rhs, preds, pre, post)


def _stmt_exprs(n: c.Node) -> List[c.Node]:
    if isinstance(n, c.Assign):
        return [n.target, n.value]
    if isinstance(n, c.Decl):
        return [n.init] if n.init is not None else []
    if isinstance(n, c.If):
        return [n.cond]
    return []


def match_native_reduce(func: c.Function, info: ReductionInfo) -> Optional[Dict[str, str]]:
    if info.preds or info.pre or info.post:
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

# workspace layout of libnompk (csrc/kernels/nompk_gridreduce.cuh; tests compare with nompk_reduce_workspace_layout())
WS_TICKET, WS_GROUP_TICKET, WS_L2, WS_L1 = 0, 64, 64 + 4 * 2048, 64 + 4 * 2048 + 8 * 2048
RED_GROUP, RED_MAX_GROUPS = 32, 2048
RED_MAX_CTAS = RED_GROUP * RED_MAX_GROUPS
RED_SINGLE_LEVEL = 2048


def _limits(t: c.CType, hi: bool) -> str:
    if t.is_float:
        return ("" if hi else "-") + ("__int_as_float(0x7f800000)" if t.base == "float" else "__longlong_as_double(0x7ff0000000000000LL)")
    bits = t.size * 8
    if t.unsigned:
        return f"({cuda_type(t)})~({cuda_type(t)})0" if hi else "0"
    if bits == 32:
        return "2147483647" if hi else "(-2147483647 - 1)"
    return "9223372036854775807LL" if hi else "(-9223372036854775807LL - 1)"


def _elementwise_arrays(func: c.Function, info: "ReductionInfo"):
    """If every access to a pointer parameter in the loop body is `a[i]` and all those arrays have one element size
    (4 or 8 bytes), return (arrays read or written, arrays written); else None -> scalar, non-substituting schedule."""
    from .ir import map_expr
    params = _params(func)
    arrays = {k: p for k, p in params.items() if p.is_array and k != info.var}
    invariant = _fam().invariant_arrays(info.pre + info.post, list(info.preds) + [info.rhs], arrays)
    arrays = {k: p for k, p in arrays.items() if k not in invariant}    # `alpha[0]`: a scalar in device memory
    used, written, ok = [], [], [True]

    def visit(e):
        if isinstance(e, c.Subscript) and isinstance(e.base, c.Name) and e.base.id in arrays:
            if _is_elem(e, arrays, info.loop.var) is None:
                ok[0] = False
            elif e.base.id not in used:
                used.append(e.base.id)
        return e

    for n in walk(info.pre + info.post):
        if isinstance(n, c.Assign):
            map_expr(n.target, visit)
            map_expr(n.value, visit)
            if isinstance(n.target, c.Subscript) and isinstance(n.target.base, c.Name) and n.target.base.id in arrays \
                    and n.target.base.id not in written:
                written.append(n.target.base.id)
        elif isinstance(n, c.Decl):
            if n.dims:
                ok[0] = False
            map_expr(n.init, visit)
        elif isinstance(n, c.If):
            map_expr(n.cond, visit)
        elif isinstance(n, (c.Break, c.Continue)):
            ok[0] = False
    for p_ in info.preds:
        map_expr(p_, visit)
    map_expr(info.rhs, visit)
    sizes = {arrays[a].ctype.size for a in used}
    if not ok[0] or len(sizes) > 1 or (sizes and sizes.pop() not in (4, 8)):
        return None
    return used, written


def emit_reduce_skeleton(knl: Kernel, info: ReductionInfo, sm_count: int) -> Tuple[str, List[str], List[str], List[str]]:
    """Single-pass reduction kernel with the schedule of libnompk's reduce.cu for an arbitrary right-hand side, with
    optional conditions and elementwise updates in front of the accumulation.  One tile of 256 threads per CTA (up to
    65536 CTAs, then the CTAs stride), 128-bit loads/stores when every array is elementwise and 16-byte aligned, and
    the two-level deterministic ticket finish of nompk_gridreduce.cuh.
    Returns (source, grid exprs, block exprs, kernel parameter names); the trailing nine parameters (workspace,
    result, result_host, seq, and the peer description of the fused all-reduce: exchange-buffer table, rank, world,
    the collective call counter in device memory, the error word in mapped host memory) are supplied by the backend."""
    from .emit_cuda import GenericEmitter
    from .ir import map_expr, map_stmts
    T = cuda_type(info.vtype)
    func = knl.func
    params = [p for p in func.params if p.name != info.var]
    written_any = set()
    for node in walk(info.pre + info.post):
        if isinstance(node, c.Assign) and isinstance(node.target, c.Subscript) and isinstance(node.target.base, c.Name):
            written_any.add(node.target.base.id)
    sig_parts = []
    for prm in params:
        t = prm.ctype
        if prm.is_array:
            sig_parts.append(f"{cuda_type(t)} *{prm.name}" if prm.name in written_any else f"const {cuda_type(t)} *__restrict__ {prm.name}")
        else:
            sig_parts.append(f"{cuda_type(t)} {prm.name}")
    sig_parts += ["void *__restrict__ nomp_ws", f"{T} *__restrict__ nomp_result", f"{T} *__restrict__ nomp_result_host",
                  "unsigned long long nomp_seq", "void *const *__restrict__ nomp_peers", "int nomp_rank", "int nomp_world",
                  "unsigned long long *nomp_cseq_dev", "unsigned long long *nomp_err_host"]
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

    # loop-invariant subexpressions (scalars in device memory, divisions of them) are evaluated once per CTA
    all_arrays = {p.name: p for p in params if p.is_array}
    inv_arrays = _fam().invariant_arrays(info.pre + info.post, list(info.preds) + [info.rhs], all_arrays)
    assigned = {n.target.id for n in walk(info.pre + info.post) if isinstance(n, c.Assign) and isinstance(n.target, c.Name)}
    declared = {n.name for n in walk(info.pre + info.post) if isinstance(n, c.Decl)}
    inv_scalars = {p.name for p in params if not p.is_array} - assigned - declared - {info.loop.var}
    npre = len(info.pre)
    stmts, exprs, prologue = _fam().hoist_invariants(info.pre + info.post, list(info.preds) + [info.rhs], inv_arrays, inv_scalars)
    info = ReductionInfo(info.loop, info.var, info.vtype, info.op, exprs[-1], exprs[:-1], stmts[:npre], stmts[npre:])

    ew = _elementwise_arrays(func, info)
    arrays = {p.name: p for p in params if p.is_array}
    if ew is not None:    # loop-invariant reads keep their subscript
        arrays = {k: p for k, p in arrays.items() if k in ew[0]}
    if ew is not None:
        sig_parts = [(f"{cuda_type(p.ctype)} *__restrict__ {p.name}" if p.name in written_any else
                      f"const {cuda_type(p.ctype)} *__restrict__ {p.name}") if p.is_array else f"{cuda_type(p.ctype)} {p.name}"
                     for p in params] + sig_parts[len(params):]

    def body_lines(depth: int, substitute: bool) -> str:
        def scalarise(e):
            if substitute and isinstance(e, c.Subscript) and isinstance(e.base, c.Name) and e.base.id in arrays:
                return c.Name(f"nomp_{e.base.id}_i")
            return e
        ge = GenericEmitter(knl)
        ge.lines = []
        ge.stmts(map_stmts(info.pre, scalarise), depth, True, False)
        cond = " && ".join(expr_str(map_expr(p_, scalarise)) for p_ in info.preds)
        pad = "  " * depth
        acc = [f"{pad}{{ const {T} nomp_v = ({T})({expr_str(map_expr(info.rhs, scalarise))}); nomp_acc = {comb('nomp_acc', 'nomp_v')}; }}"]
        if cond:
            acc = [f"{pad}if ({cond})", "  " + acc[0]]
        front = ge.lines
        ge.lines = []
        ge.stmts(map_stmts(info.post, scalarise), depth, True, False)   # behind the accumulation, as written
        return "\n".join(front + acc + ge.lines)

    shfl = f"nomp_o = __shfl_xor_sync(0xffffffffu, nomp_acc, nomp_s); nomp_acc = {comb('nomp_acc', 'nomp_o')};"
    tree = f"for (int nomp_s = 16; nomp_s > 0; nomp_s >>= 1) {{ {shfl} }}"
    slot = 8 // info.vtype.size

    if ew is not None and ew[0]:
        used, written = ew
        esize = arrays[used[0]].ctype.size
        lanes = 16 // esize
        ety = {a: cuda_type(arrays[a].ctype) for a in used}
        align = " | ".join(f"(nomp_u64_t)({a} + nomp_lo)" for a in used)
        vec_decl = "\n".join(f"      struct __align__(16) {{ {ety[a]} v[{lanes}]; }} nomp_{a}_v;" for a in used)
        vec_load = "\n".join(f"      *reinterpret_cast<int4 *>(&nomp_{a}_v) = *reinterpret_cast<const int4 *>({a} + nomp_e);" for a in used)
        vec_store = "\n".join(f"      *reinterpret_cast<int4 *>({a} + nomp_e) = *reinterpret_cast<int4 *>(&nomp_{a}_v);" for a in written)
        lane_in = "\n".join(f"        {ety[a]} nomp_{a}_i = nomp_{a}_v.v[nomp_l];" for a in used)
        lane_out = "\n".join(f"        nomp_{a}_v.v[nomp_l] = nomp_{a}_i;" for a in written)
        sc_in = "\n".join(f"      {ety[a]} nomp_{a}_i = {a}[nomp_e];" for a in used)
        sc_out = "\n".join(f"      {a}[nomp_e] = nomp_{a}_i;" for a in written)
        U = 4   # vectors per thread and tile: a CTA owns a CONTIGUOUS tile of 256 * U vectors (DRAM locality), all loads
        #         of the tile can be in flight before the first use (every array is __restrict__ in this mode: each
        #         element is only touched by the thread-iteration that owns it)
        loop = f"""  const long long nomp_tid = (long long)blockIdx.x * 256 + threadIdx.x, nomp_nthreads = (long long)gridDim.x * 256;
  if (nomp_n > 0 && ((({align}) & 15u) == 0)) {{
    const long long nomp_nvec = nomp_n / {lanes};
    for (long long nomp_base = (long long)blockIdx.x * {256 * U}; nomp_base < nomp_nvec; nomp_base += (long long)gridDim.x * {256 * U}) {{
#pragma unroll
      for (int nomp_u = 0; nomp_u < {U}; nomp_u++) {{
      const long long nomp_v = nomp_base + nomp_u * 256 + threadIdx.x;
      if (nomp_v < nomp_nvec) {{
      const long long nomp_e = nomp_lo + nomp_v * {lanes};
{vec_decl}
{vec_load}
#pragma unroll
      for (int nomp_l = 0; nomp_l < {lanes}; nomp_l++) {{
        const {it} {info.loop.var} = ({it})(nomp_e + nomp_l);
{lane_in}
{body_lines(4, True)}
{lane_out}
      }}
{vec_store}
      }}
      }}
    }}
    for (long long nomp_e = nomp_lo + nomp_nvec * {lanes} + nomp_tid; nomp_e < nomp_hi; nomp_e += nomp_nthreads) {{
      const {it} {info.loop.var} = ({it})nomp_e;
{sc_in}
{body_lines(3, True)}
{sc_out}
    }}
  }} else {{
    for (long long nomp_e = nomp_lo + nomp_tid; nomp_e < nomp_hi; nomp_e += nomp_nthreads) {{
      const {it} {info.loop.var} = ({it})nomp_e;
{sc_in}
{body_lines(3, True)}
{sc_out}
    }}
  }}"""
        per_block = 256 * lanes * U
    else:
        loop = f"""  for (long long nomp_e = nomp_lo + (long long)blockIdx.x * 256 + threadIdx.x; nomp_e < nomp_hi;
       nomp_e += (long long)gridDim.x * 256) {{
    const {it} {info.loop.var} = ({it})nomp_e;
{body_lines(2, False)}
  }}"""
        per_block = 256

    src = f"""{PRELUDE}
// Publication of the grid's result by the first warp of the finishing CTA (text twin of finish_result in
// csrc/kernels/nompk_gridreduce.cuh).  With nomp_world > 1 the all-reduce over the ranks happens right here: lane r
// stores the value into rank r's exchange buffer ({{value, call number}} per rank and slot, mapped over NVLink), waits
// for rank r's value in this rank's buffer, and the values are folded in rank order.
__device__ __forceinline__ void nomp_finish({T} nomp_v, {T} *nomp_result, {T} *nomp_result_host, unsigned long long nomp_seq,
                                            void *const *nomp_peers, int nomp_rank, int nomp_world,
                                            unsigned long long *nomp_cseq_dev, unsigned long long *nomp_err_host) {{
  const int nomp_lane = threadIdx.x & 31;
  bool nomp_late = false;
  unsigned long long nomp_cseq = 0;
  if (nomp_world > 1) {{
    if (nomp_lane == 0) {{   // number of this collective call: a counter in device memory (nothing for a graph to freeze)
      nomp_cseq = *(volatile unsigned long long *)nomp_cseq_dev + 1;
      *(volatile unsigned long long *)nomp_cseq_dev = nomp_cseq;
    }}
    nomp_cseq = __shfl_sync(0xffffffffu, nomp_cseq, 0);
    nomp_v = __shfl_sync(0xffffffffu, nomp_v, 0);
    const size_t nomp_slot = (size_t)(nomp_cseq & 1ull) * (size_t)nomp_world;
    {T} nomp_got = {ident};
    bool nomp_ok = true;
    if (nomp_lane < nomp_world) {{
      char *nomp_dst = (char *)nomp_peers[nomp_lane] + (nomp_slot + (size_t)nomp_rank) * 16;
      *(volatile {T} *)nomp_dst = nomp_v;
      __threadfence_system();
      *(volatile unsigned long long *)(nomp_dst + 8) = nomp_cseq;
      const char *nomp_src = (const char *)nomp_peers[nomp_rank] + (nomp_slot + (size_t)nomp_lane) * 16;
      unsigned long long nomp_t0, nomp_t1;
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(nomp_t0));
      while (*(const volatile unsigned long long *)(nomp_src + 8) != nomp_cseq) {{
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(nomp_t1));
        if (nomp_t1 - nomp_t0 > 20000000000ull) {{ nomp_ok = false; break; }}   // a rank that never arrives: give up after 20 s
      }}
      __threadfence_system();
      nomp_got = *(const volatile {T} *)nomp_src;
    }}
    {T} nomp_acc = __shfl_sync(0xffffffffu, nomp_got, 0);
    for (int nomp_r = 1; nomp_r < nomp_world; nomp_r++) {{
      const {T} nomp_o = __shfl_sync(0xffffffffu, nomp_got, nomp_r);
      nomp_acc = {comb('nomp_acc', 'nomp_o')};
    }}
    nomp_late = __any_sync(0xffffffffu, !nomp_ok);
    nomp_v = nomp_acc;
  }}
  if (nomp_lane == 0) {{
    *nomp_result = nomp_v;
    if (nomp_late && nomp_err_host) *(volatile unsigned long long *)nomp_err_host = nomp_cseq;
    if (nomp_result_host) {{
      *(volatile {T} *)nomp_result_host = nomp_v;
      __threadfence_system();
      *(volatile unsigned long long *)((char *)nomp_result_host + 8) = nomp_seq;
    }}
  }}
}}

// reduce clause on `{info.var}` (op {info.op}): single-pass schedule of libnompk reduce.cu with a generated loop body
extern "C" __global__ void __launch_bounds__(256) {knl.name}({', '.join(sig_parts)}) {{
  {T} nomp_acc = {ident};
  {T} nomp_o;
  const long long nomp_lo = (long long)({lo}), nomp_hi = (long long)({hi});
  const long long nomp_n = nomp_hi - nomp_lo;
{prologue}{loop}
  // ---- block tree, then the two-level deterministic ticket finish of nompk_gridreduce.cuh -----------------------------
  __shared__ {T} nomp_warp[8];
  __shared__ int nomp_role;
  char *nomp_w = (char *)nomp_ws;
  unsigned int *nomp_ticket = (unsigned int *)(nomp_w + {WS_TICKET});
  unsigned int *nomp_gticket = (unsigned int *)(nomp_w + {WS_GROUP_TICKET});
  {T} *nomp_l2 = ({T} *)(nomp_w + {WS_L2});
  {T} *nomp_l1 = ({T} *)(nomp_w + {WS_L1});
  {tree}
  if ((threadIdx.x & 31) == 0) nomp_warp[threadIdx.x >> 5] = nomp_acc;
  __syncthreads();
  if (threadIdx.x < 32) {{
    nomp_acc = threadIdx.x < 8 ? nomp_warp[threadIdx.x] : {ident};
    {tree}
  }}
  const unsigned int nomp_b = blockIdx.x, nomp_nb = gridDim.x;
  if (nomp_nb <= {RED_SINGLE_LEVEL}) {{   // few CTAs: one ticket level (a single CTA publishes directly)
    if (nomp_nb > 1) {{
      if (threadIdx.x == 0) {{
        nomp_l1[(size_t)nomp_b * {slot}] = nomp_acc;
        __threadfence();
        nomp_role = (atomicAdd(nomp_ticket, 1u) == nomp_nb - 1) ? 2 : 0;
      }}
      __syncthreads();
      if (nomp_role != 2) return;
      __threadfence();
      nomp_acc = {ident};
      for (unsigned int nomp_i = threadIdx.x; nomp_i < nomp_nb; nomp_i += 256) {{
        nomp_o = __ldcg(nomp_l1 + (size_t)nomp_i * {slot});
        nomp_acc = {comb('nomp_acc', 'nomp_o')};
      }}
      {tree}
      if ((threadIdx.x & 31) == 0) nomp_warp[threadIdx.x >> 5] = nomp_acc;
      __syncthreads();
      if (threadIdx.x < 32) {{
        nomp_acc = threadIdx.x < 8 ? nomp_warp[threadIdx.x] : {ident};
        {tree}
      }}
    }}
    if (threadIdx.x == 0 && nomp_nb > 1) *nomp_ticket = 0u;
    if (threadIdx.x < 32) nomp_finish(nomp_acc, nomp_result, nomp_result_host, nomp_seq, nomp_peers, nomp_rank, nomp_world, nomp_cseq_dev, nomp_err_host);
    return;
  }}
  const unsigned int nomp_g = nomp_b / {RED_GROUP}, nomp_ng = (nomp_nb + {RED_GROUP - 1}) / {RED_GROUP};
  const unsigned int nomp_gs = (nomp_g == nomp_ng - 1) ? nomp_nb - nomp_g * {RED_GROUP} : {RED_GROUP};
  if (threadIdx.x == 0) {{
    nomp_l1[(size_t)nomp_b * {slot}] = nomp_acc;
    __threadfence();
    nomp_role = (atomicAdd(&nomp_gticket[nomp_g], 1u) == nomp_gs - 1) ? 1 : 0;
  }}
  __syncthreads();
  if (nomp_role == 0) return;
  if (threadIdx.x < 32) {{
    __threadfence();
    nomp_acc = threadIdx.x < nomp_gs ? __ldcg(nomp_l1 + ((size_t)nomp_g * {RED_GROUP} + threadIdx.x) * {slot}) : {ident};
    {tree}
    if (threadIdx.x == 0) {{
      nomp_l2[(size_t)nomp_g * {slot}] = nomp_acc;
      nomp_gticket[nomp_g] = 0u;
      __threadfence();
      nomp_role = (atomicAdd(nomp_ticket, 1u) == nomp_ng - 1) ? 2 : 0;
    }}
  }}
  __syncthreads();
  if (nomp_role != 2) return;
  __threadfence();
  nomp_acc = {ident};
  for (unsigned int nomp_i = threadIdx.x; nomp_i < nomp_ng; nomp_i += 256) {{
    nomp_o = __ldcg(nomp_l2 + (size_t)nomp_i * {slot});
    nomp_acc = {comb('nomp_acc', 'nomp_o')};
  }}
  {tree}
  if ((threadIdx.x & 31) == 0) nomp_warp[threadIdx.x >> 5] = nomp_acc;
  __syncthreads();
  if (threadIdx.x < 32) {{
    nomp_acc = threadIdx.x < 8 ? nomp_warp[threadIdx.x] : {ident};
    {tree}
    if (threadIdx.x == 0) *nomp_ticket = 0u;
    nomp_finish(nomp_acc, nomp_result, nomp_result_host, nomp_seq, nomp_peers, nomp_rank, nomp_world, nomp_cseq_dev, nomp_err_host);
  }}
}}
"""
    extent = c.BinOp("-", info.loop.hi, info.loop.lo) if const_int(info.loop.lo) != 0 else info.loop.hi
    try:
        ext = grid_expr_str(extent, int_params)
        # one tile per CTA for large loops (two ticket levels), <= 2048 striding CTAs (one level) below 2^26 iterations;
        # the expression language has no conditional, so: min(tiles, 2048 + (tiles / 32768) * 63488)
        tiles = f"(({ext} + {per_block - 1}) / {per_block})"
        grid = f"max(1, min(min({tiles}, {RED_SINGLE_LEVEL} + ({tiles} / {(1 << 26) // per_block}) * {RED_MAX_CTAS - RED_SINGLE_LEVEL}), {RED_MAX_CTAS}))"
    except KernelError:
        grid = str(max(1, sm_count) * 8)  # data-dependent bounds: a full grid, the loop guards itself
    names = [p.name for p in params] + ["nomp_ws", "nomp_result", "nomp_result_host", "nomp_seq", "nomp_peers", "nomp_rank",
                                        "nomp_world", "nomp_cseq_dev", "nomp_err_host"]
    return src, [grid, "1", "1"], ["256", "1", "1"], names
