"""Entry points that src/loopy.c calls (names mirror the reference's loopy_api / reduction modules)."""
from __future__ import annotations

from typing import Dict, List, Tuple

from . import cparse as c
from . import families as fam
from .emit_cuda import emit_generic
from .ir import Kernel, KernelError, fix_parameters as _fix_parameters, tag_inames

_SUPPORTED_AX_N = (6, 8, 10, 12)  # nompk_ax_supported(); the backend re-checks at build time


def c_to_loopy(src: str, backend: str = "cuda") -> Kernel:
    """C kernel string -> Kernel (reference python/loopy_api.py:769-821).  Raises SyntaxError on bad input;
    libnomp reports that as NOMP_LOOPY_CONVERSION_FAILURE."""
    text = src.strip()
    try:
        func = c.parse_kernel(text)
    except c.CSyntaxError:
        # nompcc may hand over a bare loop nest; wrap it only if it really is one
        if text.startswith("for"):
            raise
        raise
    names = [p.name for p in func.params]
    if len(set(names)) != len(names):
        raise SyntaxError("duplicate parameter name in kernel signature")
    for n in func.body:
        _reject_reserved(n)
    knl = Kernel(func, source=text, target=backend)
    knl.original = knl.copy().func  # the untransformed loop nest, for the family recogniser
    return knl


def _reject_reserved(node):
    from .ir import walk
    for n in walk([node]):
        nm = getattr(n, "name", None) or getattr(n, "var", None)
        if isinstance(nm, str) and nm.startswith(("_nomp_var", "nomp_")):
            raise SyntaxError(f"kernel variables must not start with 'nomp_' / '_nomp_var': {nm}")


def realize_reduction(knl: Kernel, var: str, op: str, context: Dict) -> Kernel:
    """Record the reduce clause (reference python/reduction.py:30-134).  The loop structure is validated here so
    that errors surface at nomp_jit() time; code generation happens in get_knl_src()."""
    op = {"+": "+", "*": "*", "min": "min", "max": "max"}.get(op)
    if op is None:
        raise KernelError("reduce: operator must be one of '+', '*', 'min', 'max'")
    new = knl.copy()
    new.original = getattr(knl, "original", knl.func)
    roles = fam.match_ax_dot(knl)
    if roles is not None and roles["pap"] == var and op == "+":
        new.reduction = (var, op)
        new.ax_dot = roles          # the one reduction over a loop NEST that has a native kernel
        return new
    if op == "+" and fam.may_be_ax_dot(new.original, var):
        # another spelling of Ax + p.Ap?  Decided by what it computes (axprobe.py) once n is known: get_knl_src()
        new.reduction = (var, op)
        new.ax_probe = True
        return new
    fam.analyse_reduction(new.original, var, op)
    new.reduction = (var, op)
    return new


def fix_parameters(knl: Kernel, params: Dict) -> Kernel:
    new = _fix_parameters(knl, **params)
    orig = Kernel(getattr(knl, "original", knl.func))
    new.original = _fix_parameters(orig, **params).func
    if hasattr(knl, "ax_dot"):
        new.ax_dot = knl.ax_dot
    if hasattr(knl, "ax_probe"):
        new.ax_probe = knl.ax_probe
    return new


def annotate_passthrough(knl: Kernel, annotations: Dict, context: Dict) -> Kernel:
    """Default annotation function when no annotations script is configured: remember the keys, change nothing."""
    new = knl.copy()
    new.original = getattr(knl, "original", knl.func)
    new.annotations.update(annotations)
    return new


def get_knl_name(knl: Kernel) -> str:
    return knl.name


def _header(func=None, _ro=(), **kv) -> str:
    """`ro`: the arrays the backend may treat as read-only (their mappings keep their version: cached staging of D, no
    ordering against copies in flight) -- the ones declared const, and for the native operators also the ones the
    operator only reads by definition (`_ro`), whatever the spelling declares."""
    if func is not None:
        names = [p.name for p in func.params if p.is_array and (p.ctype.const or p.name in _ro)]
        if names:
            kv["ro"] = ",".join(names)
    return "//!nomp " + " ".join(f"{k}={v}" for k, v in kv.items() if v is not None) + "\n"


def _plan(knl: Kernel, context: Dict) -> Tuple[str, List[str], List[str]]:
    cached = getattr(knl, "_plan", None)
    if cached is not None:
        return cached
    sm_count = int(context.get("device::multiprocessor_count", 148) or 148)
    original = getattr(knl, "original", knl.func)
    one = ["1", "1", "1"]

    roles = getattr(knl, "ax_dot", None)
    if knl.reduction is not None and roles is not None:
        n_val = knl.fixed.get(roles["n"])
        if n_val is None or int(n_val) not in _SUPPORTED_AX_N:
            raise KernelError(f"the fused Ax + dot kernel needs n as a NOMP_JIT argument, one of {_SUPPORTED_AX_N}")
        extra = {"r": roles["res"], "beta": roles["beta"], "beta_dev": roles["beta_dev"]} if roles["family"] == "axxpaydot" else {}
        ro = (roles["res"], roles["g"], roles["D"]) if roles["family"] == "axxpaydot" else (roles["u"], roles["g"], roles["D"])
        plan = (_header(original, _ro=ro, kind="native", family=roles["family"], n=int(n_val), E=roles["E"], u=roles["u"], g=roles["g"],
                        D=roles["D"], w=roles["w"], out=roles["pap"], **extra), one, one)
        knl._plan = plan
        return plan

    if knl.reduction is not None and getattr(knl, "ax_probe", False):
        found = fam.match_ax_structural(original, knl.fixed, _SUPPORTED_AX_N, knl.reduction[0])
        if found is not None:
            n_val, roles = found
            plan = (_header(original, _ro=(roles["u"], roles["g"], roles["D"]), kind="native", family="axdot", n=n_val, E=roles["E"],
                            u=roles["u"], g=roles["g"], D=roles["D"], w=roles["w"], out=roles["pap"]), one, one)
            knl._plan = plan
            return plan
        # not the operator: the ordinary reduce clause decides (and reports what it cannot do)

    if knl.reduction is not None:
        var, op = knl.reduction
        info = fam.analyse_reduction(original, var, op)
        rdt = fam._dtype_code(info.vtype)
        native = fam.match_native_reduce(original, info)
        if native is not None:
            plan = (_header(original, kind="native", family="reduce", op=native["op"], dtype=native["dtype"], x=native["x"],
                            y=native.get("y", "-"), n=native["n"], out=native["out"]), one, one)
        else:
            if rdt is None:
                raise KernelError("reduce: unsupported accumulator type")
            helper = Kernel(original)
            helper.func.name = knl.name
            src, grid, block, names = fam.emit_reduce_skeleton(helper, info, sm_count)
            plan = (_header(original, kind="nvrtc", family="reduce", reduce=1, out=var, rdtype=rdt, params=",".join(names)) + src,
                    grid, block)
        knl._plan = plan
        return plan

    roles = fam.match_ax(knl)
    if roles is not None:
        n_name = roles["n"]
        n_val = knl.fixed.get(n_name)
        if n_val is not None and int(n_val) in _SUPPORTED_AX_N:
            plan = (_header(original, _ro=(roles["u"], roles["g"], roles["D"]), kind="native", family="ax", n=int(n_val), E=roles["E"],
                            u=roles["u"], g=roles["g"], D=roles["D"], w=roles["w"]), one, one)
            knl._plan = plan
            return plan
        if n_val is not None and all(t is None for t in knl.tags().values()):
            # no tuned kernel for this n and no schedule from the user: one block per element with the points as
            # threads and ur / us / ut in shared memory (what nomp_sem.py's element_loop / dof_loop clauses give);
            # one thread per element when the points do not fit a block
            max_threads = int(context.get("device::max_threads_per_block", 1024) or 1024)
            if int(n_val) ** 3 <= max_threads:
                knl = tag_inames(knl, {roles["e"]: "g.0", roles["i"]: "l.0", roles["j"]: "l.1", roles["k"]: "l.2"})
            else:
                from .ir import split_iname
                outer = roles["e"]
                k2 = split_iname(knl, outer, 32)
                knl = tag_inames(k2, {f"{outer}_outer": "g.0", f"{outer}_inner": "l.0"})

    if knl.reduction is None and roles is None:
        found = fam.match_ax_structural(original, knl.fixed, _SUPPORTED_AX_N, None)
        if found is not None:    # any other spelling of the operator (axprobe.py)
            n_val, sroles = found
            plan = (_header(original, _ro=(sroles["u"], sroles["g"], sroles["D"]), kind="native", family="ax", n=n_val, E=sroles["E"],
                            u=sroles["u"], g=sroles["g"], D=sroles["D"], w=sroles["w"]), one, one)
            knl._plan = plan
            return plan

    if knl.reduction is None:
        native = fam.match_native_map(original)
        if native is not None:
            plan = (_header(original, kind="native", family="map", op=fam.MAP_OPS[native["op"]], dtype=native["dtype"],
                            y=native["y"], x=native.get("x", "-"), z=native.get("z", "-"),
                            alpha=native.get("alpha", "-"), beta=native.get("beta", "-"), n=native["n"]), one, one)
            knl._plan = plan
            return plan
        loop = fam.match_map_skeleton(original)
        if loop is not None:
            helper = Kernel(original)
            helper.func.name = knl.name
            src, grid, block = fam.emit_map_skeleton(helper, loop, sm_count)
            names = ",".join(p.name for p in helper.func.params)
            plan = (_header(original, kind="nvrtc", family="map", params=names) + src, grid, block)
            knl._plan = plan
            return plan

    src, grid, block = emit_generic(knl)
    names = ",".join(p.name for p in knl.func.params)
    plan = (_header(original, kind="nvrtc", family="generic", params=names) + src, grid, block)
    knl._plan = plan
    return plan


def get_knl_src(knl: Kernel, context: Dict) -> str:
    """Descriptor line + (for the NVRTC paths) CUDA source.  Replaces reference python/loopy_api.py:824-826."""
    return _plan(knl, context)[0]


def get_grid_size(knl: Kernel, context: Dict):
    """((gx, gy, gz), (bx, by, bz)) as expression strings for src/gridexpr.c (reference src/loopy.c:406-449)."""
    _, grid, block = _plan(knl, context)
    return tuple(grid), tuple(block)
