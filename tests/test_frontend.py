"""Host-side jit logic without a GPU: the C-subset parser, the loop transformations user scripts call through the
`loopy` shim, the loop-family recogniser, and the CUDA emitter (every generated kernel is compiled by NVRTC for
sm_100a and, where threads are independent, executed on the host through tests/cuda_emulation.py and compared with the
kernel string itself compiled by gcc)."""
import ctypes as C
import math
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "libnomp_b200" / "python"))

import loopy as lp  # noqa: E402  (the shim)
import nomp_bridge as nb  # noqa: E402
from nomp_bridge import cparse, families  # noqa: E402
from nomp_bridge.gridexpr_py import evaluate as grid_eval  # noqa: E402

from tests.cuda_emulation import emulate, nvrtc_compile  # noqa: E402
from tests.kernel_oracle import run_kernel  # noqa: E402

CTX = {"backend::name": "cuda", "device::max_threads_per_block": 1024, "device::multiprocessor_count": 148}
TYPES = ["int", "long", "unsigned", "unsigned long", "double", "float"]
NP = {"int": np.int32, "long": np.int64, "unsigned": np.uint32, "unsigned long": np.uint64, "double": np.float64,
      "float": np.float32}


def tile(knl, context):
    (iname,) = knl.default_entrypoint.all_inames()
    bs = min(512, context["device::max_threads_per_block"])
    knl = lp.split_iname(knl, iname, bs, inner_iname=f"{iname}_inner", outer_iname=f"{iname}_outer")
    return lp.tag_inames(knl, {f"{iname}_outer": "g.0", f"{iname}_inner": "l.0"})


def tile_outer(knl, context):
    (i, j) = sorted(knl.default_entrypoint.all_inames())
    knl = lp.split_iname(knl, i, 512, inner_iname=f"{i}_inner", outer_iname=f"{i}_outer")
    return lp.tag_inames(knl, {f"{i}_outer": "g.0", f"{i}_inner": "l.0", j: "for"})


def tile2d(knl, context):
    bs = int(math.sqrt(min(1024, context["device::max_threads_per_block"])))
    knl = lp.split_iname(knl, "i", bs)
    knl = lp.split_iname(knl, "j", bs)
    return lp.tag_inames(knl, {"i_outer": "g.0", "i_inner": "l.0", "j_outer": "g.1", "j_inner": "l.1"})


def plan(src, transform=None, reduce=None, fixed=None):
    k = nb.c_to_loopy(src, "cuda")
    if transform:
        k = transform(k, CTX)
    if reduce:
        k = nb.realize_reduction(k, reduce[0], reduce[1], CTX)
    if fixed:
        k = nb.fix_parameters(k, fixed)
    text = nb.get_knl_src(k, CTX)
    header, _, body = text.partition("\n")
    desc = dict(kv.split("=", 1) for kv in header.split()[1:])
    return desc, body, nb.get_grid_size(k, CTX), k


# ---- parser --------------------------------------------------------------------------------------------------------

def test_parser_accepts_the_reference_subset():
    f = cparse.parse_kernel("""
        void foo(const double *a, unsigned long *b, int N, float alpha) {
          for (int i = 0; i < N; i++) {
            int t = 0;
            double s[4][2];
            for (unsigned j = b[i]; j <= b[i + 1]; j++) {
              if ((!(j < 3) && (j < 5)) || j == 1) continue;
              t += (j % 2 == 0) ? 1 : 2;
              if (t > 100) break;
            }
            b[i] = ~b[i] ^ (t << 2) | 1;
          }
        }""")
    assert f.name == "foo" and [p.name for p in f.params] == ["a", "b", "N", "alpha"]
    assert f.params[0].ctype.const and f.params[0].is_array and not f.params[2].is_array
    assert f.params[1].ctype.unsigned and f.params[1].ctype.base == "long"
    loop = f.body[0]
    assert isinstance(loop, cparse.For) and loop.var == "i"
    inner = [n for n in loop.body if isinstance(n, cparse.For)][0]
    assert isinstance(inner.hi, cparse.BinOp) and inner.hi.op == "+"       # <= became < hi + 1


@pytest.mark.parametrize("bad", [
    "void foo(int *a, int N) { for (int i = 0; i < N; i++) a[i] = i }",            # missing ';' (reference test 100)
    "void foo(int *a, int N) { for (int i = 0; i < N; i--) a[i] = i; }",           # update must be ++
    "void foo(int *a, int N) { for (int i = 0; i > N; i++) a[i] = i; }",           # condition must be < or <=
    "void foo(int *a, int N) { int i; for (i = 0; i < N; i++) a[i] = i; }",        # variable must be declared in init
    "void foo(int *a, int N) { while (N) a[0] = 1; }",
    "void foo(int *a, int N) { for (int i = 0; i < N; i++) a[i] = i; } int x;",
    "int main",
    "void foo(int *a, int N) { for (int i = 0; i < N; i++) { a[i] = i; }",        # unbalanced
    "void foo(int *a, int N) { for (int i = 0; i < N; i++) a[i] = i++; }",
])
def test_parser_rejects(bad):
    with pytest.raises(SyntaxError):
        nb.c_to_loopy(bad, "cuda")


def test_reserved_prefix_is_rejected():
    with pytest.raises(SyntaxError):
        nb.c_to_loopy("void f(int *a, int N) { for (int i = 0; i < N; i++) { int _nomp_var0 = 1; a[i] = _nomp_var0; } }")


# ---- loopy shim ------------------------------------------------------------------------------------------------------

def test_split_and_tag_inames():
    k = nb.c_to_loopy("void f(double *a, int N) { for (int i = 0; i < N; i++) a[i] = i; }")
    assert k.default_entrypoint.all_inames() == frozenset({"i"})
    k2 = lp.split_iname(k, "i", 128)
    assert set(k2.default_entrypoint.all_inames()) == {"i_outer", "i_inner"}
    assert k.default_entrypoint.all_inames() == frozenset({"i"}), "transformations must not mutate their argument"
    k3 = lp.tag_inames(k2, [("i_outer", "g.0"), ("i_inner", "l.0")])
    assert k3.tags() == {"i_outer": "g.0", "i_inner": "l.0"}
    with pytest.raises(lp.LoopyError):
        lp.tag_inames(k2, {"nope": "g.0"})
    with pytest.raises(lp.LoopyError):
        lp.tag_inames(k2, {"i_outer": "g.7"})
    with pytest.raises(lp.LoopyError):
        lp.split_iname(k2, "i", 32)          # i no longer exists


def test_glob_tags_and_repeated_loop_names():
    k = nb.c_to_loopy("""void f(double *b, const double *a, int n) { for (int i = 0; i < n; i++) { double s[32];
        for (int j = 0; j < 32; j++) s[j] = a[i * 32 + j];
        for (int j = 0; j < 32; j++) b[i * 32 + j] = s[j]; } }""")
    assert sorted(k.default_entrypoint.all_inames()) == ["i", "j"]
    k = lp.tag_inames(k, {"i": "g.0", "j*": "l.0"})
    assert [l.tag for l in k.loops()] == ["g.0", "l.0", "l.0"]


def test_sem_annotation_script_semantics():
    """grid_loop -> split + g.0/l.0; element_loop -> g.0; dof_loop's result is discarded by the reference script
    (reference tests/sem.py:20-24), which must be tolerated."""
    k = nb.c_to_loopy("void f(double *a, double *b, int N) { for (int i = 0; i < N; i++) a[i] += b[i]; }")
    assert lp.translation_unit.TranslationUnit is type(k)
    bs = min(512, CTX["device::max_threads_per_block"])
    lp.split_iname(k, "i", bs, inner_tag="l.0")          # discarded, no side effect
    assert k.tags() == {"i": None}
    g = lp.tag_inames(lp.split_iname(k, "i", bs), [("i_outer", "g.0"), ("i_inner", "l.0")])
    assert g.tags() == {"i_outer": "g.0", "i_inner": "l.0"}


# ---- family recogniser -------------------------------------------------------------------------------------------------

@pytest.mark.parametrize("body,op,extra", [
    ("a[i] += b[i];", "add", {"x": "b"}),
    ("a[i] = b[i] + a[i];", "add", {"x": "b"}),
    ("a[i] -= b[i];", "sub", {"x": "b"}),
    ("a[i] *= b[i];", "mul", {"x": "b"}),
    ("a[i] = a[i] * b[i];", "mul", {"x": "b"}),
    ("a[i] += s * b[i];", "axpy", {"x": "b", "alpha": "s"}),
    ("a[i] = a[i] + b[i] * s;", "axpy", {"x": "b", "alpha": "s"}),
    ("a[i] = b[i] + s * a[i];", "xpay", {"x": "b", "alpha": "s"}),
    ("a[i] = s * b[i] + t * a[i];", "axpby", {"x": "b", "alpha": "s", "beta": "t"}),
    ("a[i] = s * a[i];", "scale", {"alpha": "s"}),
    ("a[i] = b[i];", "copy", {"x": "b"}),
    ("a[i] = s;", "fill", {"alpha": "s"}),
    ("a[i] = b[i] + c[i];", "add3", {"x": "b", "z": "c"}),
])
def test_native_map_recognition(body, op, extra):
    for T in ("double", "float", "int", "unsigned long"):
        src = f"void f({T} *a, const {T} *b, const {T} *c, {T} s, {T} t, int N) {{ for (int i = 0; i < N; i++) {body} }}"
        desc, cuda, grid, _ = plan(src, tile)
        assert desc["kind"] == "native" and desc["family"] == "map", (body, desc)
        assert int(desc["op"]) == families.MAP_OPS[op] and desc["y"] == "a" and desc["n"] == "N"
        for k, v in extra.items():
            assert desc[k] == v
        assert cuda == "" and grid == (("1", "1", "1"), ("1", "1", "1"))


@pytest.mark.parametrize("body", [
    "a[i] -= b[i] + 1;", "a[i] = a[i] * a[i] + b[i] * b[i];", "a[i] = 2 * b[i] + 1;", "a[i] = a[i] + 3 * b[i] + 2 * c[i];",
    "a[i] = i;", "a[i] = a[i] & 3;", "a[i] = ~a[i];", "a[i] = (b[i] > 0) ? b[i] : c[i];",
])
def test_other_elementwise_loops_get_the_vector_skeleton(body):
    for T in ("int", "unsigned long", "double", "float"):
        if T in ("double", "float") and ("&" in body or "~" in body):
            continue
        src = f"void f({T} *a, const {T} *b, const {T} *c, int N) {{ for (int i = 0; i < N; i++) {body} }}"
        desc, cuda, grid, _ = plan(src, tile)
        assert desc["kind"] == "nvrtc" and desc["family"] == "map" and desc["params"] == "a,b,c,N"
        ok, log = nvrtc_compile(cuda)
        assert ok, log
        assert "int4" in cuda and grid[1] == ("256", "1", "1")


def test_mixed_width_or_offset_access_is_not_a_map():
    desc, _, _, _ = plan("void f(double *a, const int *b, int N) { for (int i = 0; i < N; i++) a[i] = b[i]; }", tile)
    assert desc["family"] == "generic"
    desc, _, _, _ = plan("void f(double *a, const double *b, int N) { for (int i = 0; i < N; i++) a[i] = b[i + 1]; }", tile)
    assert desc["family"] == "generic"


def test_reduce_recognition_and_errors():
    desc, _, _, _ = plan("void f(double *a, int N, double *s) { for (int i = 0; i < N; i++) { s[0] += a[i]; } }", None, ("s", "+"))
    assert (desc["kind"], desc["family"], desc["x"], desc["y"], desc["out"], int(desc["dtype"])) == ("native", "reduce", "a", "-", "s", 5)
    desc, _, _, _ = plan("void f(long *a, long *b, int N, long *t) { for (int i = 0; i < N; i++) t[0] += a[i] * b[i]; }", None, ("t", "+"))
    assert (desc["x"], desc["y"], int(desc["dtype"]), int(desc["op"])) == ("a", "b", 2, 0)
    desc, _, _, _ = plan("void f(float *a, int N, float *m) { for (int i = 0; i < N; i++) m[0] = (a[i] < m[0]) ? a[i] : m[0]; }", None, ("m", "min"))
    assert (desc["family"], int(desc["op"])) == ("reduce", 2) and desc["kind"] == "native"
    for body in ("s[0] += 1;", "s[0] += i;", "if (a[i] > 0) s[0] += 1;", "s[0] += a[i] * a[i] + 1;"):
        desc, cuda, grid, _ = plan(f"void f(double *a, int N, double *s) {{ for (int i = 0; i < N; i++) {{ {body} }} }}", None, ("s", "+"))
        assert desc["kind"] == "nvrtc" and desc["family"] == "reduce" and desc["out"] == "s"
        assert desc["params"].endswith("nomp_ws,nomp_result,nomp_result_host,nomp_seq,nomp_peers,nomp_rank,nomp_world,nomp_cseq_dev,nomp_err_host")
        ok, log = nvrtc_compile(cuda)
        assert ok, log
    with pytest.raises(nb.KernelError):   # two loops
        plan("void f(double *a, int N, double *s) { for (int i = 0; i < N; i++) for (int j = 0; j < N; j++) s[0] += a[i]; }", None, ("s", "+"))
    with pytest.raises(nb.KernelError):   # not a parameter
        plan("void f(double *a, int N, double *s) { for (int i = 0; i < N; i++) s[0] += a[i]; }", None, ("q", "+"))
    with pytest.raises(nb.KernelError):   # bad operator
        plan("void f(double *a, int N, double *s) { for (int i = 0; i < N; i++) s[0] += a[i]; }", None, ("s", "-"))


def test_ax_recognition():
    k = nb.c_to_loopy(families.AX_KERNEL_SOURCE)
    for n in (6, 8, 10, 12):
        kn = nb.fix_parameters(k, {"n": n})
        header = nb.get_knl_src(kn, CTX)
        assert header.startswith(f"//!nomp kind=native family=ax n={n} E=E u=u g=g D=D w=w ro=u,g,D")
    import re
    renamed = families.AX_KERNEL_SOURCE
    for old, new in (("nomp_ax", "my_ax"), ("ur", "gr"), ("w", "out"), ("E", "nel")):
        renamed = re.sub(rf"\b{old}\b", new, renamed)
    kr = nb.fix_parameters(nb.c_to_loopy(renamed), {"n": 8})
    assert "family=ax" in nb.get_knl_src(kr, CTX) and "w=out" in nb.get_knl_src(kr, CTX) and "E=nel" in nb.get_knl_src(kr, CTX)
    # an n without a hand-written kernel goes through NVRTC with one block per element, the points as threads and the
    # per-element temporaries in shared memory; one thread per element when the points do not fit a block
    k9 = nb.fix_parameters(k, {"n": 9})
    text = nb.get_knl_src(k9, CTX)
    assert "kind=nvrtc family=generic" in text and "__shared__ double ur[9][9][9];" in text
    ok, log = nvrtc_compile(text)
    assert ok, log
    assert nb.get_grid_size(k9, CTX) == (("E", "1", "1"), ("9", "9", "9"))
    k11 = nb.fix_parameters(k, {"n": 11})
    assert nb.get_grid_size(k11, CTX) == (("((E + 31) / 32)", "1", "1"), ("32", "1", "1"))
    ok, log = nvrtc_compile(nb.get_knl_src(k11, CTX))
    assert ok, log
    # a changed loop body is not the Ax family
    other = nb.fix_parameters(nb.c_to_loopy(families.AX_KERNEL_SOURCE.replace("double acc = 0;", "double acc = 1;")), {"n": 8})
    assert "family=ax" not in nb.get_knl_src(other, CTX)


def _signature_order(src):
    import re
    params = re.search(r"\(([^)]*)\)", src).group(1).split(",")
    return [re.sub(r"\[.*", "", prm.strip()).split()[-1].lstrip("*") for prm in params]


def _run_variant_with_gcc(src, roles, n, E, seed):
    """The variant's own C text on exact data -> (w, pap or None) and the arrays it was given."""
    rng = np.random.default_rng(seed)
    n3 = n ** 3
    data = {"w": rng.integers(-9, 10, E * n3).astype(np.float64), "u": rng.integers(-4, 5, E * n3).astype(np.float64),
            "g": rng.integers(0, 4, 6 * E * n3).astype(np.float64), "D": rng.integers(-2, 3, n * n).astype(np.float64),
            "E": E, "n": n, "pap": np.zeros(1)}
    by_name = dict(zip(roles, ("w", "u", "g", "D", "E", "n", "pap")))
    run_kernel(src, *[data[by_name[name]] for name in _signature_order(src)])
    return data


@pytest.mark.parametrize("n", [8, 6])
def test_every_spelling_of_ax_reaches_the_native_kernel(n):
    """The structural recogniser (nomp_bridge/axprobe.py): six spellings of the operator that share no token sequence
    with the canonical string -- other names and argument order, g[e][f][k][j][i] and D[a][l], flat / merged / extra
    temporaries, loops and statements in another order, `+=` into a zeroed w -- are all routed to nompk_ax_f64 with the
    right argument roles, and so are two spellings of Ax + p.Ap under a reduce clause.  The texts themselves are the
    operator: gcc runs each against the oracle first.  Near misses keep the generic schedule."""
    from oracle import ffi
    from tests import ax_variants as V
    E = 3
    for name, src, roles in V.variants() + V.fused_variants():
        d = _run_variant_with_gcc(src, roles, 4, E, 7)
        want = ffi.ax(4, d["u"], d["g"], d["D"])
        assert np.array_equal(d["w"], want), f"the test's own variant {name} is not the operator"
        fused = len(roles) == 7
        if fused:
            assert d["pap"][0] == float(d["u"] @ want)
        k = nb.c_to_loopy(src)
        if fused:
            k = nb.realize_reduction(k, roles[6], "+", CTX)
        header = nb.get_knl_src(nb.fix_parameters(k, {roles[5]: n}), CTX).splitlines()[0]
        want_header = (f"//!nomp kind=native family={'axdot' if fused else 'ax'} n={n} E={roles[4]} u={roles[1]} g={roles[2]} "
                       f"D={roles[3]} w={roles[0]}" + (f" out={roles[6]}" if fused else ""))
        assert header.startswith(want_header), (name, header)
        # what the operator only reads is read-only for the backend whether or not the spelling says const: the mapping
        # of D keeps its version, so D is staged (and looked at for its symmetry) once, not at every launch
        ro = header.split(" ro=")[1].split()[0].split(",")
        assert {roles[1], roles[2], roles[3]} <= set(ro) and roles[0] not in ro, (name, header)
    for name, src in V.not_ax():
        header = nb.get_knl_src(nb.fix_parameters(nb.c_to_loopy(src), {"n": n}), CTX).splitlines()[0]
        assert "kind=nvrtc" in header and "family=ax" not in header, (name, header)
    # a reduce clause over a nest that is not the operator is still reported, at code generation
    bad = V.fused_variants()[0][1].replace("pap[0] += acc * u[", "pap[0] += 2 * acc * u[")
    k = nb.fix_parameters(nb.realize_reduction(nb.c_to_loopy(bad), "pap", "+", CTX), {"n": n})
    with pytest.raises(nb.KernelError):
        nb.get_knl_src(k, CTX)


def test_fused_cg_families():
    k = nb.c_to_loopy(families.AX_DOT_KERNEL_SOURCE)
    k = nb.fix_parameters(nb.realize_reduction(k, "pap", "+", CTX), {"n": 10})
    assert nb.get_knl_src(k, CTX).startswith("//!nomp kind=native family=axdot n=10 E=E u=u g=g D=D w=w out=pap ro=u,g,D")
    with pytest.raises(nb.KernelError):        # no generic fallback for a reduction over a loop nest
        nb.get_knl_src(nb.fix_parameters(nb.realize_reduction(nb.c_to_loopy(families.AX_DOT_KERNEL_SOURCE), "pap", "+", CTX), {"n": 9}), CTX)
    with pytest.raises(nb.KernelError):        # the plain Ax string with a reduce clause is not a known family
        nb.realize_reduction(nb.c_to_loopy(families.AX_KERNEL_SOURCE), "w", "+", CTX)
    upd = ("void upd(double *x, double *r, const double *p, const double *w, double alpha, int N, double *rr) {"
           " for (int i = 0; i < N; i++) { x[i] += alpha * p[i]; r[i] -= alpha * w[i]; rr[0] += r[i] * r[i]; } }")
    desc, cuda, grid, _ = plan(upd, None, ("rr", "+"))
    assert desc["kind"] == "nvrtc" and desc["family"] == "reduce" and desc["ro"] == "p,w"
    assert "double *__restrict__ x, double *__restrict__ r, const double *__restrict__ p" in cuda
    ok, log = nvrtc_compile(cuda)
    assert ok, log
    with pytest.raises(nb.KernelError):
        plan("void bad(double *x, int N, double *s) { for (int i = 0; i < N; i++) { x[0] = i; s[0] += x[i]; } }", None, ("s", "+"))


# ---- generic emitter: compile with NVRTC and execute on the host -----------------------------------------------------------

def _ptr(a):
    return C.c_void_p(a.ctypes.data)


@pytest.mark.parametrize("T", TYPES)
def test_generic_inner_loop_with_continue_break_and_ternary(T):
    src = f"""void foo({T} *a, int N) {{
      for (int i = 0; i < N; i++) {{
        int t = 0;
        for (int j = 0; j < 10; j++) {{
          if ((!(j < 3) && (j < 5)) || j == 1) continue;
          t += (j == 7) ? 3 : 1;
          if (j == 8) break;
        }}
        a[i] = t + i;
      }}
    }}"""
    desc, cuda, (grid, block), _ = plan(src, tile_outer)
    assert desc["family"] == "generic" and block[0] == "512"
    ok, log = nvrtc_compile(cuda)
    assert ok, log
    n = 700
    want = np.zeros(n, dtype=NP[T])
    run_kernel(src, want, n)
    got = np.zeros(n, dtype=NP[T])
    g = (grid_eval(grid[0], {"N": n}), 1, 1)
    emulate(cuda, "foo", g, (512, 1, 1), [f"{'unsigned long long' if T == 'unsigned long' else ('long long' if T == 'long' else T)} *", "int"],
            [_ptr(got), C.c_int(n)])
    assert np.array_equal(got, want)


def test_generic_data_dependent_bounds():
    src = """void foo(double *a, int *b, int N) { for (int i = 0; i < N; i++) { int t = 0;
        for (int j = b[i]; j < b[i + 1] + 1; j++) { t += 1; } a[i] = t; } }"""
    desc, cuda, (grid, block), _ = plan(src, tile_outer)
    ok, log = nvrtc_compile(cuda)
    assert ok, log
    n = 50
    b = (2 * np.arange(n + 1)).astype(np.int32)
    got = np.zeros(n)
    emulate(cuda, "foo", (1, 1, 1), (512, 1, 1), ["double *", "int *", "int"], [_ptr(got), _ptr(b), C.c_int(n)])
    assert np.array_equal(got, np.full(n, 3.0))


def test_generic_2d_tiling_transpose_and_matmul():
    src = """void foo(double *a, double *b, int rows, int cols) { for (int j = 0; j < rows; j++)
        for (int i = 0; i < cols; i++) a[j + i * rows] = b[i + j * cols]; }"""
    desc, cuda, (grid, block), _ = plan(src, tile2d)
    assert grid[:2] == ("((cols + 31) / 32)", "((rows + 31) / 32)") and block[:2] == ("32", "32")
    ok, log = nvrtc_compile(cuda)
    assert ok, log
    rows, cols = 40, 5
    b = np.arange(rows * cols, dtype=np.float64)
    got = np.zeros(rows * cols)
    g = (grid_eval(grid[0], dict(rows=rows, cols=cols)), grid_eval(grid[1], dict(rows=rows, cols=cols)), 1)
    emulate(cuda, "foo", g, (32, 32, 1), ["double *", "double *", "int", "int"], [_ptr(got), _ptr(b), C.c_int(rows), C.c_int(cols)])
    assert np.array_equal(got.reshape(cols, rows), b.reshape(rows, cols).T)

    mm = """void foo(long *a, long *b, long *c, int size) { for (unsigned i = 0; i < size; i++) { for (unsigned j = 0; j < size; j++) {
        double dot = 0; for (unsigned k = 0; k < size; k++) dot += a[i * size + k] * b[k * size + j]; c[i * size + j] = dot; } } }"""
    desc, cuda, (grid, block), k = plan(mm, lambda kn, ctx: lp.tag_inames(tile2d(kn, ctx), {"k": "for"}))
    ok, log = nvrtc_compile(cuda)
    assert ok, log
    n = 10
    a = np.tile(np.arange(n, dtype=np.int64), n)
    bm = np.repeat(np.arange(n, dtype=np.int64), n)
    got = np.zeros(n * n, dtype=np.int64)
    emulate(cuda, "foo", (1, 1, 1), (32, 32, 1), ["long long *"] * 3 + ["int"], [_ptr(a), _ptr(bm), _ptr(got), C.c_int(n)])
    assert np.all(got == sum(i * i for i in range(n)))


def test_block_level_arrays_become_shared_memory_with_barriers():
    src = """void foo(double *b, const double *a, int n, int m) { for (int i = 0; i < n; i++) { double s[m][m];
        for (int j = 0; j < m; j++) s[j][4] = a[i * m + j];
        for (int j = 0; j < m; j++) s[j][4] += s[j][4];
        for (int j = 0; j < m; j++) b[i * m + j] = s[j][4]; } }"""
    desc, cuda, (grid, block), _ = plan(src, lambda k, c: lp.tag_inames(k, {"i": "g.0", "j*": "l.0"}), fixed={"m": 16})
    assert desc["params"] == "b,a,n" and grid[0] == "n" and block[0] == "16"
    assert "__shared__ double s[16][16];" in cuda and cuda.count("__syncthreads();") == 3
    ok, log = nvrtc_compile(cuda)
    assert ok, log
    n, m = 16, 16
    a = np.repeat(np.arange(n, dtype=np.float64), m)
    got = np.zeros(n * m)
    emulate(cuda, "foo", (n, 1, 1), (m, 1, 1), ["double *", "const double *", "int"], [_ptr(got), _ptr(a), C.c_int(n)])
    assert np.array_equal(got, 2 * a)
    with pytest.raises(nb.KernelError):     # VLA extent must be fixed at jit time
        plan(src, lambda k, c: lp.tag_inames(k, {"i": "g.0", "j*": "l.0"}))


def test_untransformed_kernel_runs_as_a_single_thread():
    desc, cuda, (grid, block), _ = plan("void f(double *a, int *b, int N) { for (int i = 0; i < N; i++) a[b[i]] = i; }")
    assert desc["family"] == "generic" and grid == ("1", "1", "1") and block == ("1", "1", "1")
    ok, log = nvrtc_compile(cuda)
    assert ok, log


def test_map_skeleton_executes_correctly_on_the_host():
    src = "void foo(unsigned *a, const unsigned *b, int N) { for (int i = 0; i < N; i++) a[i] = (a[i] ^ b[i]) + (i << 1); }"
    desc, cuda, (grid, block), _ = plan(src, tile)
    assert desc["family"] == "map" and desc["kind"] == "nvrtc"
    for n in (0, 1, 5, 1000, 1027):
        a = np.arange(n, dtype=np.uint32) * 7
        b = np.arange(n, dtype=np.uint32) * 13 + 1
        want = a.copy()
        run_kernel(src, want, b, n)
        emulate(cuda, "foo", (grid_eval(grid[0], {"N": n}), 1, 1), (256, 1, 1), ["unsigned *", "const unsigned *", "int"],
                [_ptr(a), _ptr(b), C.c_int(n)])
        assert np.array_equal(a, want), n


# ---- launch-size expressions ---------------------------------------------------------------------------------------------

def test_grid_expression_grammar_matches_the_c_evaluator():
    """src/gridexpr.c compiled standalone and compared with the Python mirror on the expressions the emitter prints."""
    import subprocess
    import tempfile
    d = Path(tempfile.mkdtemp())
    stub = d / "stub.c"
    stub.write_text('#include <stddef.h>\nint nomp_log_(const char *f, unsigned l, int e, int t, const char *fmt, ...) { return 1; }\n'
                    'unsigned nomp_log_get_verbose(void) { return 0; }\n')
    inc = ROOT / "libnomp_b200" / "csrc" / "libnomp"
    import sysconfig
    subprocess.run(["gcc", "-shared", "-fPIC", "-o", str(d / "gx.so"), str(inc / "src" / "gridexpr.c"), str(stub),
                    "-I", str(inc / "include"), "-I", str(ROOT / "include"), "-I", sysconfig.get_paths()["include"]], check=True)
    lib = C.CDLL(str(d / "gx.so"))
    names = (C.c_char_p * 3)(b"N", b"rows", b"cols")
    values = (C.c_long * 3)(1000, 40, 5)
    env = {"N": 1000, "rows": 40, "cols": 5}
    for expr in ["1", "((N + 511) / 512)", "max(1, min((N + 1023) / 1024, 1184))", "((cols + 31) / 32)", "N", "rows * cols - 3 % 2",
                 "max(((rows + 31) / 32), 7)", "-(N) + 2000", "min(N, rows) * 2"]:
        out = C.c_long()
        assert lib.nomp_gridexpr_eval(expr.encode(), names, values, 3, C.byref(out)) == 0, expr
        assert out.value == grid_eval(expr, env), expr
    for bad in ["N +", "foo", "(N", "N / 0", "max(N)", "N N"]:
        out = C.c_long()
        assert lib.nomp_gridexpr_eval(bad.encode(), names, values, 3, C.byref(out)) != 0, bad


def test_nomp_sem_annotations_give_one_block_per_element():
    """libnomp_b200/python/nomp_sem.py honours dof_loop (the reference's tests/sem.py drops it): an operator that is
    not the canonical Ax string -- here with a mass term added -- gets one thread block per element, the dof loops as
    thread axes in clause order, and its per-element temporaries in shared memory between two barriers."""
    import nomp_sem
    src = families.AX_KERNEL_SOURCE.replace("const double *D, int E, int n)", "const double *D, const double *h, int E, int n)")
    src = src.replace("nomp_ax(", "helmholtz(")
    point = "e * n * n * n + k * n * n + j * n + i"
    assert f"w[{point}] = acc;" in src
    src = src.replace(f"w[{point}] = acc;", f"w[{point}] = acc + h[{point}] * u[{point}];")
    k = nb.c_to_loopy(src, "cuda")
    for key, loop in (("element_loop", "e"), ("dof_loop", "i"), ("dof_loop", "j"), ("dof_loop", "k"), ("dof_loop", "zz")):
        k = nomp_sem.annotate(k, {key: loop}, CTX)
    assert k.tags() == {"e": "g.0", "i": "l.0", "j": "l.1", "k": "l.2", "l": None}
    with pytest.raises(ValueError):
        nomp_sem.annotate(k, {"dof_loop": "l"}, CTX)
    k = nb.fix_parameters(k, {"n": 6})
    text = nb.get_knl_src(k, CTX)
    header, _, cuda = text.partition("\n")
    assert "family=generic" in header and "kind=nvrtc" in header
    assert nb.get_grid_size(k, CTX) == (("E", "1", "1"), ("6", "6", "6"))
    for name in ("ur", "us", "ut"):
        assert f"__shared__ double {name}[6][6][6];" in cuda
    assert cuda.count("__syncthreads();") == 2
    ok, log = nvrtc_compile(cuda)
    assert ok, log
    # grid_loop behaves like the reference script
    flat = nb.c_to_loopy("void f(double *a, double *b, int N) { for (int i = 0; i < N; i++) a[i] += b[i]; }")
    flat = nomp_sem.annotate(flat, {"grid_loop": "i"}, CTX)
    assert flat.tags() == {"i_outer": "g.0", "i_inner": "l.0"}


# ---- randomized fidelity of the parser + emitter --------------------------------------------------------------------------

def _random_expr(rng, depth, leaves, int_ops):
    if depth == 0 or rng.random() < 0.25:
        return leaves[rng.integers(len(leaves))]
    kind = rng.random()
    a = _random_expr(rng, depth - 1, leaves, int_ops)
    b = _random_expr(rng, depth - 1, leaves, int_ops)
    if kind < 0.08:
        return f"(-{a})" if not a.startswith("(-") else a
    if kind < 0.16:
        c = _random_expr(rng, depth - 1, leaves, int_ops)
        return f"(({a} < {b}) ? {c} : {b})"
    ops = ["+", "-", "*"] + (["&", "|", "^"] if int_ops else [])
    op = ops[rng.integers(len(ops))]
    return f"({a} {op} {b})" if rng.random() < 0.7 else f"{a} {op} {b}"


@pytest.mark.parametrize("T", ["double", "float", "int", "unsigned", "long"])
def test_random_expressions_survive_parsing_and_code_generation(T):
    """Twenty-four random right-hand sides per type (precedence with and without parentheses, unary minus, ternaries, bit
    operations for the integer types): the generated CUDA kernel, executed on the host thread by thread, returns the
    bits of the kernel string compiled by gcc -- through the vector skeleton (no transform) and through the generic
    emitter (user tiling)."""
    rng = np.random.default_rng({"double": 1, "float": 2, "int": 3, "unsigned": 4, "long": 5}[T])
    dt = NP[T]
    is_int = np.issubdtype(dt, np.integer)
    cuda_t = {"long": "long long", "unsigned long": "unsigned long long"}.get(T, T)
    n = 257
    for case in range(24):
        leaves = ["a[i]", "b[i]", "c[i]", "s", "t", "3", "i"] if is_int else ["a[i]", "b[i]", "c[i]", "s", "t", "0.5", "2"]
        expr = _random_expr(rng, 3, leaves, is_int)
        src = (f"void foo({T} *a, const {T} *b, const {T} *c, {T} s, {T} t, int N) "
               f"{{ for (int i = 0; i < N; i++) a[i] = {expr}; }}")
        if is_int:
            a0 = rng.integers(0, 50, n).astype(dt)
            b0, c0 = rng.integers(0, 50, n).astype(dt), rng.integers(0, 50, n).astype(dt)
            sv, tv = dt(3), dt(5)
        else:
            a0 = rng.uniform(0.5, 1.5, n).astype(dt)
            b0, c0 = rng.uniform(0.5, 1.5, n).astype(dt), rng.uniform(0.5, 1.5, n).astype(dt)
            sv, tv = dt(0.75), dt(1.25)
        ctype = {np.float64: C.c_double, np.float32: C.c_float, np.int32: C.c_int, np.uint32: C.c_uint, np.int64: C.c_longlong}[dt]
        want = a0.copy()
        run_kernel(src, want, b0, c0, ctype(sv), ctype(tv), n)
        argtypes = [f"{cuda_t} *", f"const {cuda_t} *", f"const {cuda_t} *", cuda_t, cuda_t, "int"]
        for transform in (None, tile):
            desc, cuda, (grid, block), _ = plan(src, transform)
            if desc["kind"] == "native":      # the draw happens to be one of libnompk's maps: covered on the GPU
                continue
            got = a0.copy()
            g = (grid_eval(grid[0], {"N": n}), 1, 1)
            emulate(cuda, "foo", g, (int(block[0]), 1, 1), argtypes,
                    [_ptr(got), _ptr(b0), _ptr(c0), ctype(sv), ctype(tv), C.c_int(n)])
            assert np.array_equal(got, want), (T, case, desc["family"], expr)


# ---- the accepted kernel language, feature by feature ----------------------------------------------------

def _tile_first(knl, ctx):
    i = sorted(knl.default_entrypoint.all_inames())[0]
    knl = lp.split_iname(knl, i, 64)
    return lp.tag_inames(knl, {f"{i}_outer": "g.0", f"{i}_inner": "l.0"})


LANGUAGE_CASES = {
 "char_short_bool": ("void f(char *a, short *b, unsigned char *c, int N) { for (int i = 0; i < N; i++) { a[i] = a[i] + 1; b[i] = b[i] * 2; c[i] = !c[i]; } }",
     [("char *", np.int8), ("short *", np.int16), ("unsigned char *", np.uint8)]),
 "le_loop": ("void f(int *a, int N) { for (int i = 0; i <= N - 1; i++) a[i] = i % 7 + (i >> 1) - (i / 3); }", [("int *", np.int32)]),
 "div_assign": ("void f(double *a, int N) { for (int i = 0; i < N; i++) a[i] /= 4; }", [("double *", np.float64)]),
 "nested_2d_decl": ("void f(double *a, int N) { for (int i = 0; i < N; i++) { double t[2][3]; for (int j = 0; j < 2; j++) for (int k = 0; k < 3; k++) t[j][k] = a[i] * j + k; a[i] = t[1][2] - t[0][1]; } }", [("double *", np.float64)]),
 "if_else_chain": ("void f(int *a, int N) { for (int i = 0; i < N; i++) { if (a[i] > 10) a[i] = 1; else if (a[i] > 5) a[i] = 2; else a[i] = 3; } }", [("int *", np.int32)]),
 "logical": ("void f(int *a, int N) { for (int i = 0; i < N; i++) a[i] = (a[i] > 3 && a[i] < 9) || a[i] == 12; }", [("int *", np.int32)]),
 "bitops": ("void f(unsigned *a, int N) { for (int i = 0; i < N; i++) a[i] = (~a[i] & 255) | (a[i] << 3) ^ (a[i] >> 2); }", [("unsigned *", np.uint32)]),
 "float_literals": ("void f(float *a, int N) { for (int i = 0; i < N; i++) a[i] = a[i] * 0.5f + 1.5e-1f - .25f; }", [("float *", np.float32)]),
 "casts": ("void f(double *a, int N) { for (int i = 0; i < N; i++) a[i] = (double)((int)a[i] / 2) + (float)i; }", [("double *", np.float64)]),
 "while_like_break": ("void f(int *a, int N) { for (int i = 0; i < N; i++) { int s = 0; for (int j = 0; j < 100; j++) { if (j > a[i]) break; if (j % 2) continue; s += j; } a[i] = s; } }", [("int *", np.int32)]),
 "long_unsigned_long": ("void f(long *a, unsigned long *b, int N) { for (int i = 0; i < N; i++) { a[i] = a[i] * 3 - b[i]; b[i] = b[i] + a[i]; } }", [("long long *", np.int64), ("unsigned long long *", np.uint64)]),
 "math_calls": ("void f(double *a, int N) { for (int i = 0; i < N; i++) a[i] = sqrt(a[i]) + fabs(a[i] - 3.0) + fmin(a[i], 2.0); }", [("double *", np.float64)]),
 "scalar_decl_no_init": ("void f(double *a, int N) { for (int i = 0; i < N; i++) { double t; t = a[i]; t *= t; a[i] = t; } }", [("double *", np.float64)]),
 "hoisted_bound": ("void f(int *a, int *b, int N) { for (int i = 0; i < N; i++) { int s = 0; for (int j = b[i]; j < b[i + 1]; j++) s += j; a[i] = s; } }", [("int *", np.int32), ("int *", np.int32)]),
 "array_param_syntax": ("void f(double a[], const double b[], int N) { for (int i = 0; i < N; i++) a[i] += b[i] * b[i]; }", [("double *", np.float64), ("const double *", np.float64)]),
 "uint_loop_var": ("void f(int *a, unsigned N) { for (unsigned i = 0; i < N; i++) a[i] = i * 2; }", [("int *", np.int32)]),
 "pre_increment": ("void f(int *a, int N) { for (int i = 0; i < N; ++i) a[i] = i; }", [("int *", np.int32)]),
 "step_plus_equal": ("void f(int *a, int N) { for (int i = 0; i < N; i += 1) a[i] = i; }", [("int *", np.int32)]),
 "comments": ("void f(int *a, int N) { /* block */ for (int i = 0; i < N; i++) { // line\n a[i] = i; } }", [("int *", np.int32)]),
 "modulo_neg": ("void f(int *a, int N) { for (int i = 0; i < N; i++) a[i] = (a[i] - 20) % 7 + (a[i] - 20) / 3; }", [("int *", np.int32)]),
}


@pytest.mark.parametrize("name", sorted(LANGUAGE_CASES))
def test_kernel_language_feature(name):
    """One kernel per language feature (narrow integer types, <= loops, /=, N-D temporaries, else-if chains, logical and
    bit operators, float literals, casts, break / continue, math calls, data-dependent inner bounds, array-syntax
    parameters, unsigned loop variables, ++i and i += 1, comments, signed % and /): the generated CUDA -- with and
    without a user tiling -- executed on the host returns the bits of the kernel string compiled by gcc."""
    src, arrs = LANGUAGE_CASES[name]
    n = 100
    rng = np.random.default_rng(1)
    data = []
    for _, dt in arrs:
        if np.issubdtype(dt, np.floating):
            data.append(rng.uniform(1, 9, n + 1).astype(dt))
        elif "b[i + 1]" in src and len(data) == 1:
            data.append(np.sort(rng.integers(0, 30, n + 1)).astype(dt))
        else:
            data.append(rng.integers(0, 20, n + 1).astype(dt))
    unsigned_n = "unsigned N" in src
    want = [d.copy() for d in data]
    run_kernel(src, *want, (C.c_uint(n) if unsigned_n else n))
    for transform in (None, _tile_first):
        desc, cuda, (grid, block), _ = plan(src, transform)
        assert desc["kind"] == "nvrtc", desc
        ok, log = nvrtc_compile(cuda)
        assert ok, log
        got = [d.copy() for d in data]
        emulate(cuda, "f", (grid_eval(grid[0], {"N": n}), 1, 1), (int(block[0]), 1, 1),
                [a[0] for a in arrs] + ["unsigned" if unsigned_n else "int"],
                [C.c_void_p(x.ctypes.data) for x in got] + [C.c_uint(n) if unsigned_n else C.c_int(n)])
        for a, b in zip(got, want):
            assert np.array_equal(a, b), (name, desc["family"])


def test_mutated_kernel_strings_fail_cleanly():
    """1500 random token-level mutations of four kernels (delete / insert / replace): the bridge either accepts the
    result or raises its own syntax / kernel error -- never anything else, never slowly (libnomp turns the exception
    into NOMP_LOOPY_CONVERSION_FAILURE with the message; a stray IndexError or a hang would be a bug of ours)."""
    import random
    import re
    import time
    seeds = ["void f(double *a, const double *b, int N) { for (int i = 0; i < N; i++) a[i] += b[i] * 2; }",
             "void g(int *a, int N, int *s) { for (int i = 0; i < N; i++) { if (a[i] > 3) s[0] += a[i]; else continue; } }",
             "void h(double *a, int n, int m) { for (int i = 0; i < n; i++) { double t[m]; for (int j = 0; j < m; j++) "
             "t[j] = a[i * m + j]; for (int j = 0; j < m; j++) a[i * m + j] = t[m - 1 - j]; } }",
             families.AX_KERNEL_SOURCE]
    tok = re.compile(r"[A-Za-z_]\w*|\d+\.?\d*|==|!=|<=|>=|&&|\|\||<<|>>|\+=|-=|\*=|\+\+|--|\S")
    extra = ["(", ")", "{", "}", "[", "]", ";", ",", "for", "if", "else", "int", "double", "*", "+", "-", "=", "0", "x",
             "?", ":", "<", "++"]
    rng = random.Random(5)
    allowed = (nb.KernelError, cparse.CSyntaxError)
    accepted = 0
    for _ in range(1500):
        toks = tok.findall(rng.choice(seeds))
        for _ in range(rng.randint(1, 4)):
            op, pos = rng.random(), rng.randrange(len(toks))
            if op < 0.4:
                del toks[pos]
            elif op < 0.8:
                toks.insert(pos, rng.choice(extra))
            else:
                toks[pos] = rng.choice(extra)
        text = " ".join(toks)
        t0 = time.perf_counter()
        try:
            k = nb.c_to_loopy(text, "cuda")
            if rng.random() < 0.5:
                k = nb.fix_parameters(k, {"n": 4, "m": 3})
            nb.get_knl_src(k, CTX)
            nb.get_grid_size(k, CTX)
            accepted += 1
        except allowed:
            pass
        assert time.perf_counter() - t0 < 2.0, text
    assert 0 < accepted < 300


# ---- kernels whose threads cooperate, executed on the host (tests/cuda_emulation.py: emulate_cooperative) -------------------

REDUCE_TAIL = ["void *", "{T} *", "{T} *", "unsigned long long", "void **", "int", "int", "unsigned long long *", "unsigned long long *"]


def _run_reduce_skeleton(src, var, op, T, arrays, scalars, n, grid_override=None):
    """Generated reduction kernel on the host with real barriers, shuffles and block tickets.  Returns (result,
    result as published to the 'host' block, sequence number, workspace ticket words after the run)."""
    from tests.cuda_emulation import emulate_cooperative
    desc, cuda, (grid, block), _ = plan(src, reduce=(var, op))
    assert desc["kind"] == "nvrtc" and desc["family"] == "reduce", desc
    name = src.split("(")[0].split()[-1]
    cuda_t = {"long": "long long", "unsigned long": "unsigned long long"}.get(T, T)
    dt = NP[T]
    ws = np.zeros(548928 // 8 + 8, dtype=np.uint64)      # nompk_reduce_workspace_bytes()
    res, pub = np.zeros(1, dtype=dt), np.zeros(24, dtype=np.uint8)
    params = desc["params"].split(",")[:-9]
    types, args = [], []
    for prm in params:
        if prm in arrays:
            a, ct = arrays[prm]
            types.append(ct)
            args.append(C.c_void_p(a.ctypes.data))
        else:
            ct, val = scalars[prm]
            types.append(ct)
            args.append(val)
    types += [t.format(T=cuda_t) for t in REDUCE_TAIL]
    args += [C.c_void_p(ws.ctypes.data), C.c_void_p(res.ctypes.data), C.c_void_p(pub.ctypes.data), C.c_ulonglong(41),
             C.c_void_p(0), C.c_int(0), C.c_int(1), C.c_void_p(0), C.c_void_p(0)]
    g = grid_override or grid_eval(grid[0], {"N": n})
    emulate_cooperative(cuda, name, (g, 1, 1), (256, 1, 1), types, args)
    return res[0], pub[:dt().nbytes].view(dt)[0], int(pub[8:16].view(np.uint64)[0]), ws[:2].copy()


@pytest.mark.parametrize("T", ["double", "long", "int", "float"])
@pytest.mark.parametrize("grid", [None, 1, 37, 2500])
def test_reduce_skeleton_runs_on_the_host(T, grid):
    """Sum with a condition and a non-trivial right-hand side, through the generated single-pass reduction: every
    grid shape of the finish (one CTA, one ticket level, two ticket levels) gives the serial loop's value on exact
    data, publishes it with its sequence number and leaves the tickets at zero."""
    dt = NP[T]
    n = 30011
    a = (np.arange(n) * 7 % 13).astype(dt)
    b = (np.arange(n) * 5 % 11).astype(dt)
    src = f"void red({T} *a, {T} *b, int N, {T} *s) {{ for (int i = 0; i < N; i++) if (a[i] > 2) s[0] += a[i] * b[i] + 1; }}"
    cuda_t = {"long": "long long"}.get(T, T)
    got, pub, seq, tickets = _run_reduce_skeleton(src, "s", "+", T, {"a": (a, f"{cuda_t} *"), "b": (b, f"{cuda_t} *")},
                                                  {"N": ("int", C.c_int(n))}, n, grid)
    want = np.zeros(1, dtype=dt)
    run_kernel(src, a, b, n, want)
    assert got == want[0] == pub and seq == 41 and not tickets.any()


def test_min_max_and_fused_update_skeletons_run_on_the_host():
    n = 5003
    rng = np.random.default_rng(3)
    a = rng.integers(-1000, 1000, n).astype(np.float64)
    for op, body, fn in (("min", "m[0] = (a[i] * 2 < m[0]) ? a[i] * 2 : m[0];", np.min),
                         ("max", "m[0] = (a[i] * 2 > m[0]) ? a[i] * 2 : m[0];", np.max)):
        src = f"void mm(const double *a, int N, double *m) {{ for (int i = 0; i < N; i++) {body} }}"
        got, pub, _, _ = _run_reduce_skeleton(src, "m", op, "double", {"a": (a, "const double *")}, {"N": ("int", C.c_int(n))}, n)
        assert got == fn(2 * a) == pub
    # the fused CG update: elementwise writes in front of the accumulation, vectorised path (all arrays 16-byte aligned)
    src = ("void upd(double *x, double *r, const double *p, const double *w, double alpha, int N, double *rr) {"
           " for (int i = 0; i < N; i++) { x[i] += alpha * p[i]; r[i] -= alpha * w[i]; rr[0] += r[i] * r[i]; } }")
    x, r = rng.integers(-4, 5, n).astype(np.float64), rng.integers(-4, 5, n).astype(np.float64)
    p, w = rng.integers(-4, 5, n).astype(np.float64), rng.integers(-4, 5, n).astype(np.float64)
    xw, rw, want = x.copy(), r.copy(), np.zeros(1)
    run_kernel(src, xw, rw, p, w, 0.5, n, want)
    got, _, _, _ = _run_reduce_skeleton(src, "rr", "+", "double",
                                        {"x": (x, "double *"), "r": (r, "double *"), "p": (p, "const double *"), "w": (w, "const double *")},
                                        {"alpha": ("double", C.c_double(0.5)), "N": ("int", C.c_int(n))}, n)
    assert got == want[0] and np.array_equal(x, xw) and np.array_equal(r, rw)


def test_statements_behind_the_accumulation_stay_behind_it():
    """`s[0] += a[i]; a[i] = 0;` -- the write follows the accumulation in program order (the reference keeps it:
    seq_dependencies=True, ref python/loopy_api.py:817).  Vectorised (aligned) and scalar (misaligned) schedules, and a
    statement on either side of a conditional accumulation."""
    n = 5003
    rng = np.random.default_rng(5)
    for off in (0, 1):
        base = rng.integers(-9, 10, n + 1).astype(np.float64)
        a = base[off:off + n]
        src = "void clr(double *a, int N, double *s) { for (int i = 0; i < N; i++) { s[0] += a[i]; a[i] = 0; } }"
        aw, want = a.copy(), np.zeros(1)
        run_kernel(src, aw, n, want)
        got, _, _, _ = _run_reduce_skeleton(src, "s", "+", "double", {"a": (a, "double *")}, {"N": ("int", C.c_int(n))}, n)
        assert got == want[0] != 0 and np.array_equal(a, aw) and not a.any()
    src = ("void both(double *a, double *b, int N, double *s) { for (int i = 0; i < N; i++) {"
           " b[i] = a[i] * 2; if (b[i] > 3) s[0] += a[i] + b[i]; a[i] = b[i] + 1; b[i] = a[i] * a[i]; } }")
    a, b = rng.integers(-9, 10, n).astype(np.float64), np.zeros(n)
    aw, bw, want = a.copy(), b.copy(), np.zeros(1)
    run_kernel(src, aw, bw, n, want)
    got, _, _, _ = _run_reduce_skeleton(src, "s", "+", "double", {"a": (a, "double *"), "b": (b, "double *")},
                                        {"N": ("int", C.c_int(n))}, n)
    assert got == want[0] and np.array_equal(a, aw) and np.array_equal(b, bw)
    with pytest.raises(nb.KernelError):      # the accumulator may only appear in its own update
        plan("void bad(double *a, int N, double *s) { for (int i = 0; i < N; i++) { s[0] += a[i]; a[i] = s[0]; } }", reduce=("s", "+"))


def test_annotated_element_kernel_runs_on_the_host():
    """The one-block-per-element schedule of nomp_sem.py (shared-memory temporaries between barriers) for an operator
    that is not a hand-written family, executed with one coroutine per thread: bitwise the kernel string run by gcc."""
    import nomp_sem
    from tests.cuda_emulation import emulate_cooperative
    from oracle import ffi
    point = "e * n * n * n + k * n * n + j * n + i"
    src = families.AX_KERNEL_SOURCE.replace("const double *D, int E, int n)", "const double *D, const double *h, int E, int n)")
    src = src.replace("nomp_ax(", "helmholtz(").replace(f"w[{point}] = acc;", f"w[{point}] = acc + h[{point}] * u[{point}];")
    n, E = 4, 5
    k = nb.c_to_loopy(src, "cuda")
    for key, loop in (("element_loop", "e"), ("dof_loop", "i"), ("dof_loop", "j"), ("dof_loop", "k")):
        k = nomp_sem.annotate(k, {key: loop}, CTX)
    k = nb.fix_parameters(k, {"n": n})
    header, _, cuda = nb.get_knl_src(k, CTX).partition("\n")
    u = ffi.fill_uniform_f64(E * n ** 3, 1, -1.0, 1.0)
    g = ffi.fill_uniform_f64(E * 6 * n ** 3, 2, 0.5, 1.5)
    h = ffi.fill_uniform_f64(E * n ** 3, 3, 0.0, 2.0)
    D = np.ascontiguousarray(ffi.gll_derivative(n)[0].ravel())
    want, got = np.zeros_like(u), np.zeros_like(u)
    run_kernel(src, want, u, g, D, h, E, n)
    ptr = lambda a: C.c_void_p(a.ctypes.data)  # noqa: E731
    emulate_cooperative(cuda, "helmholtz", (E, 1, 1), (n, n, n),
                        ["double *", "const double *", "const double *", "const double *", "const double *", "int"],
                        [ptr(got), ptr(u), ptr(g), ptr(D), ptr(h), C.c_int(E)])
    assert np.array_equal(got, want)
    # the Ax kernel string itself at a size without a hand-written kernel takes the same schedule
    k5 = nb.fix_parameters(nb.c_to_loopy(families.AX_KERNEL_SOURCE, "cuda"), {"n": 5})
    _, _, cuda5 = nb.get_knl_src(k5, CTX).partition("\n")
    u5 = ffi.fill_int_f64(3 * 125, 5, -4, 4)
    g5 = ffi.fill_int_f64(3 * 6 * 125, 6, 0, 3)
    D5 = ffi.fill_int_f64(25, 7, -2, 2)
    w5 = np.zeros_like(u5)
    emulate_cooperative(cuda5, "nomp_ax", (3, 1, 1), (5, 5, 5), ["double *", "const double *", "const double *", "const double *", "int"],
                        [ptr(w5), ptr(u5), ptr(g5), ptr(D5), C.c_int(3)])
    assert np.array_equal(w5, ffi.ax(5, u5, g5, D5))


def test_generated_reduction_all_reduces_between_two_host_ranks():
    """nomp_finish with nomp_world = 2: two copies of the emulated kernel run at the same time in two threads, each with
    its own workspace and exchange buffer ("peer memory" = host memory both can see).  Both end with the fold of the two
    partial sums in rank order, publish it, and the next call uses the other slot."""
    import threading
    from tests.cuda_emulation import emulate_cooperative
    src = "void red(const double *a, const double *b, int N, double *s) { for (int i = 0; i < N; i++) if (a[i] > 2) s[0] += a[i] * b[i] + 1; }"
    desc, cuda, (grid, block), _ = plan(src, reduce=("s", "+"))
    world, n = 2, 20011
    xchg = [np.zeros(2 * world * 2, dtype=np.uint64) for _ in range(world)]              # [2 slots][world]{value, seq}
    table = np.array([x.ctypes.data for x in xchg], dtype=np.uint64)
    data = [((np.arange(n + 5 * r) * (7 + r) % 13).astype(np.float64), (np.arange(n + 5 * r) * 5 % 11).astype(np.float64))
            for r in range(world)]
    parts = []
    for a, b in data:
        w = np.zeros(1)
        run_kernel(src, a, b, a.size, w)
        parts.append(w[0])
    types = ["const double *", "const double *", "int", "void *", "double *", "double *", "unsigned long long", "void **", "int",
             "int", "unsigned long long *", "unsigned long long *"]
    counters = [np.zeros(1, dtype=np.uint64) for _ in range(world)]   # the call number: a counter in each rank's "device" memory
    ws = [np.zeros(548928 // 8 + 8, dtype=np.uint64) for _ in range(world)]
    res, pub = [np.zeros(1) for _ in range(world)], [np.zeros(3, dtype=np.uint64) for _ in range(world)]
    for call in (1, 2, 3):
        errors = []

        def rank_main(r):
            try:
                a, b = data[r]
                ptr = lambda v: C.c_void_p(v.ctypes.data)  # noqa: E731
                emulate_cooperative(cuda, "red", (grid_eval(grid[0], {"N": a.size}), 1, 1), (256, 1, 1), types,
                                    [ptr(a), ptr(b), C.c_int(a.size), ptr(ws[r]), ptr(res[r]), ptr(pub[r]), C.c_ulonglong(100 + call),
                                     ptr(table), C.c_int(r), C.c_int(world), ptr(counters[r]), ptr(pub[r][2:])], instance=r)
            except Exception as exc:   # pragma: no cover
                errors.append(exc)

        threads = [threading.Thread(target=rank_main, args=(r,)) for r in range(world)]
        for t in threads:
            t.start()
        for t in threads:
            t.join(60)
        assert not errors and not any(t.is_alive() for t in threads)
        for r in range(world):
            assert res[r][0] == parts[0] + parts[1] == pub[r].view(np.float64)[0]
            assert pub[r][1] == 100 + call and pub[r][2] == 0                              # published, nobody was late
            assert counters[r][0] == call                                                  # the kernel counted the call itself
        slot = (call & 1) * world
        assert all(int(xchg[r][2 * (slot + q) + 1]) == call for r in range(world) for q in range(world))


@pytest.mark.parametrize("T,op", [("long", "+"), ("int", "+"), ("unsigned", "+"), ("double", "+"), ("int", "*"), ("long", "min"),
                                  ("double", "max")])
def test_random_reductions_run_on_the_host(T, op):
    """Ten random right-hand sides per (type, operator), some under a random condition, through the generated
    single-pass reduction executed on the host; integer-valued data, so the grid-wide fold has the serial loop's bits."""
    import zlib
    rng = np.random.default_rng(zlib.crc32(f"{T} {op}".encode()))       # a fixed seed per case (str hashes are salted)
    dt = NP[T]
    is_int = np.issubdtype(dt, np.integer)
    cuda_t = {"long": "long long"}.get(T, T)
    n = 4099
    for case in range(10):
        leaves = ["a[i]", "b[i]", "3", "i"] if is_int else ["a[i]", "b[i]", "2", "0.5"]
        expr = _random_expr(rng, 2, leaves, is_int and op == "+")
        cond = f"if ({_random_expr(rng, 1, ['a[i]', 'b[i]', '4'], False)} > 3) " if rng.random() < 0.5 else ""
        if op == "+":
            stmt = f"{cond}s[0] += {expr};"
        elif op == "*":
            expr = "(a[i] & 1) + 1"                                   # factors 1 and 2 only: 2^k stays exact for a while
            stmt = f"if (i < 25) s[0] *= {expr};"
        else:
            cmp_ = "<" if op == "min" else ">"
            stmt = f"{cond}s[0] = (({expr}) {cmp_} s[0]) ? ({expr}) : s[0];"
        src = f"void red(const {T} *a, const {T} *b, int N, {T} *s) {{ for (int i = 0; i < N; i++) {stmt} }}"
        a = rng.integers(0, 9, n).astype(dt)
        b = rng.integers(0, 9, n).astype(dt)
        want = np.array([{"+": 0, "*": 1}.get(op, 0)], dtype=dt)
        if op == "min":
            want[0] = np.iinfo(dt).max if is_int else np.inf
        if op == "max":
            want[0] = np.iinfo(dt).min if is_int else -np.inf
        run_kernel(src, a, b, n, want)
        desc = plan(src, reduce=("s", op))[0]
        if desc["kind"] == "native":
            continue
        got, pub, _, tickets = _run_reduce_skeleton(src, "s", op, T, {"a": (a, f"const {cuda_t} *"), "b": (b, f"const {cuda_t} *")},
                                                    {"N": ("int", C.c_int(n))}, n, grid_override=int(rng.integers(1, 40)))
        assert got == want[0] == pub and not tickets.any(), (T, op, case, stmt)


# ---- scalars that live in device memory ---------------------------------------------------------------------------------

def test_loop_invariant_reads_keep_the_vector_schedules():
    """`alpha[0]` -- a scalar in device memory (the result a reduce clause left there, a coefficient a previous kernel
    computed) -- read at a constant subscript of an array the loop never writes: the map and reduce skeletons keep their
    128-bit schedules and leave the read as it is.  Bitwise the kernel string compiled by gcc."""
    from tests.cuda_emulation import emulate_cooperative
    rng = np.random.default_rng(5)
    n = 4099
    # map skeleton: p = r + beta[1] * p, x += alpha[0] * q
    src = ("void dir(double *p, double *x, const double *r, const double *q, const double *coef, int N) {"
           " for (int i = 0; i < N; i++) { p[i] = r[i] + coef[1] * p[i]; x[i] += coef[0] * q[i]; } }")
    desc, cuda, (grid, block), _ = plan(src)
    assert (desc["kind"], desc["family"]) == ("nvrtc", "map"), desc
    assert "coef[1]" in cuda and "coef[0]" in cuda and "nomp_coef" not in cuda and "int4" in cuda
    p, x, r, q = (rng.integers(-4, 5, n).astype(np.float64) for _ in range(4))
    coef = np.array([0.5, -0.25])
    pw, xw = p.copy(), x.copy()
    run_kernel(src, pw, xw, r, q, coef, n)
    emulate(cuda, "dir", (grid_eval(grid[0], {"N": n}), 1, 1), (256, 1, 1),
            ["double *", "double *", "const double *", "const double *", "const double *", "int"],
            [_ptr(p), _ptr(x), _ptr(r), _ptr(q), _ptr(coef), C.c_int(n)])
    assert np.array_equal(p, pw) and np.array_equal(x, xw)
    # ... and evaluated once per CTA, not once per element (families.hoist_invariants): the coefficient of a CG update whose
    # scalars stay on the device is a division of two of them
    assert "nomp_inv" not in cuda          # a plain read is not worth a barrier
    src3 = ("void upd3(double *x, const double *p, const double *rr, const double *pap, double c, int N) {"
            " for (int i = 0; i < N; i++) x[i] += (c * rr[0] / pap[0]) * p[i] + pap[0]; }")
    desc, cuda3, (grid3, _), _ = plan(src3)
    assert desc["family"] == "map" and "nomp_inv_s0 = (((c * rr[0]) / pap[0]));" in cuda3 and "nomp_inv_s1" not in cuda3
    assert grid3[0] == "max(1, (N + 2047) / 2048)"       # four tiles per CTA behind the barrier
    assert all("/ pap[0]" not in line for line in cuda3.splitlines() if "nomp_x_i" in line), "no division per element"
    xs, ps = rng.uniform(-1, 1, n), rng.uniform(-1, 1, n)
    rr_, pap_ = np.array([3.7]), np.array([1.3])
    xs_w = xs.copy()
    run_kernel(src3, xs_w, ps, rr_, pap_, 0.7, n)
    emulate(cuda3, "upd3", (grid_eval(grid3[0], {"N": n}), 1, 1), (256, 1, 1),
            ["double *", "const double *", "const double *", "const double *", "double", "int"],
            [_ptr(xs), _ptr(ps), _ptr(rr_), _ptr(pap_), C.c_double(0.7), C.c_int(n)])
    assert np.array_equal(xs, xs_w)
    ok, log = nvrtc_compile(cuda3)
    assert ok, log
    # an array that is also written, or read at i as well, is not a scalar: no vector schedule through this rule
    for bad in ("void k(double *a, double *s, int N) { for (int i = 0; i < N; i++) { a[i] += s[0]; s[0] = a[i]; } }",
                "void k(double *a, const double *s, int N) { for (int i = 0; i < N; i++) a[i] += s[0] + s[i + 1]; }"):
        assert plan(bad)[0]["family"] == "generic"
    # reduce skeleton: the fused CG update with alpha in device memory
    src = ("void upd(double *x, double *r, const double *p, const double *w, const double *alpha, int N, double *rr) {"
           " for (int i = 0; i < N; i++) { x[i] += alpha[0] * p[i]; r[i] -= alpha[0] * w[i]; rr[0] += r[i] * r[i]; } }")
    desc, cuda, _, _ = plan(src, reduce=("rr", "+"))
    assert "alpha[0]" in cuda and "nomp_alpha" not in cuda and "int4" in cuda
    x, r, p, w = (rng.integers(-4, 5, n).astype(np.float64) for _ in range(4))
    alpha = np.array([0.5])
    xw, rw, want = x.copy(), r.copy(), np.zeros(1)
    run_kernel(src, xw, rw, p, w, alpha, n, want)
    got, _, _, _ = _run_reduce_skeleton(src, "rr", "+", "double",
                                        {"x": (x, "double *"), "r": (r, "double *"), "p": (p, "const double *"),
                                         "w": (w, "const double *"), "alpha": (alpha, "const double *")},
                                        {"N": ("int", C.c_int(n))}, n)
    assert got == want[0] and np.array_equal(x, xw) and np.array_equal(r, rw)
    ok, log = nvrtc_compile(cuda)
    assert ok, log


def test_xpay_ax_dot_recognition_and_semantics():
    """The canonical xpay + Ax + dot strings (beta by value / beta[0] in device memory), also with renamed identifiers,
    bind to the native family axxpaydot for the n that have a kernel; a different update does not.  The string itself,
    compiled by gcc, is the stand-alone sequence: p = r + beta p (oracle map), w = A p (oracle Ax), pap = p . w."""
    from oracle import ffi
    for src, dev in ((families.AX_XPAY_DOT_KERNEL_SOURCE, "0"), (families.AX_XPAY_DOT_DEV_KERNEL_SOURCE, "1")):
        desc, _, _, _ = plan(src, reduce=("pap", "+"), fixed={"n": 10})
        assert (desc["kind"], desc["family"], desc["beta_dev"]) == ("native", "axxpaydot", dev), desc
        assert (desc["u"], desc["r"], desc["beta"], desc["out"], desc["n"]) == ("p", "res", "beta", "pap", "10")
        renamed = src.replace("res", "resid").replace("beta", "b").replace(" p[", " dir[").replace("*p,", "*dir,").replace("pap", "energy")
        desc, _, _, _ = plan(renamed, reduce=("energy", "+"), fixed={"n": 8})
        assert (desc["family"], desc["u"], desc["r"], desc["beta"], desc["out"]) == ("axxpaydot", "dir", "resid", "b", "energy"), desc
    other = families.AX_XPAY_DOT_KERNEL_SOURCE.replace("+ beta * p[", "- beta * p[")
    with pytest.raises(Exception):
        plan(other, reduce=("pap", "+"), fixed={"n": 8})          # a reduction over a loop nest without a native kernel
    n, E, beta = 6, 3, 2.0
    p0 = ffi.fill_int_f64(E * n ** 3, 1, -2, 2)
    r = ffi.fill_int_f64(E * n ** 3, 2, -2, 2)
    g = ffi.fill_int_f64(E * 6 * n ** 3, 3, 0, 3)
    D = ffi.fill_int_f64(n * n, 4, -2, 2)
    want_p = r + beta * p0
    want_w = ffi.ax(n, want_p, g, D)
    for src, b in ((families.AX_XPAY_DOT_KERNEL_SOURCE, beta), (families.AX_XPAY_DOT_DEV_KERNEL_SOURCE, np.array([beta]))):
        p, w, pap = p0.copy(), np.zeros_like(p0), np.zeros(1)
        run_kernel(src, w, p, r, g, D, b, E, n, pap)
        assert np.array_equal(p, want_p) and np.array_equal(w, want_w) and pap[0] == float(want_p @ want_w)
