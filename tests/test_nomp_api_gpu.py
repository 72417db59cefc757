"""The public C API on a B200 (ctypes -> libnomp.so -> CUDA backend), compared with the kernel string compiled by gcc
(tests/kernel_oracle.py) or the oracle library, on the same inputs.  Mirrors what the reference's tests/nomp-api-*.c
pin and adds large sizes, random data and the north-star additions (min/max, Ax, extensions)."""
import ctypes as C
import os
import re
import sys
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "libnomp_b200" / "python"))

from libnomp_b200 import capi  # noqa: E402
from oracle import ffi  # noqa: E402
from tests.kernel_oracle import run_kernel  # noqa: E402

P, I, U, F, JIT = capi.NOMP_PTR, capi.NOMP_INT, capi.NOMP_UINT, capi.NOMP_FLOAT, capi.NOMP_JIT
TYPES = {"int": (np.int32, I), "long": (np.int64, I), "unsigned": (np.uint32, U), "unsigned long": (np.uint64, U),
         "double": (np.float64, F), "float": (np.float32, F)}
TR = "nomp_test_transforms"


@pytest.fixture(scope="module", autouse=True)
def runtime():
    capi.check(capi.init(backend="cuda", device=0, verbose=0, scripts_dir=ROOT / "tests" / "scripts",
                         annotations_script="nomp_test_annotations"))
    yield capi.nomp()
    assert capi.nomp().nomp_finalize_excluding_interpreter() == 0


def rand(T, n, seed):
    dt = TYPES[T][0]
    rng = np.random.default_rng(seed)
    if np.issubdtype(dt, np.floating):
        return rng.uniform(0.5, 1.5, n).astype(dt)
    return rng.integers(0, 1000, n).astype(dt)


class Mapped:
    """nomp_update(TO) on entry, FROM (for outputs) + FREE on exit."""

    def __init__(self, *arrays, out=()):
        self.arrays, self.out = arrays, out

    def __enter__(self):
        for a in self.arrays:
            capi.check(capi.update(a.ctypes.data, 0, a.size, a.itemsize, capi.NOMP_TO))
        return self

    def __exit__(self, *exc):
        for a in self.arrays:
            if any(a is o for o in self.out):
                capi.check(capi.update(a.ctypes.data, 0, a.size, a.itemsize, capi.NOMP_FROM))
            capi.check(capi.update(a.ctypes.data, 0, a.size, a.itemsize, capi.NOMP_FREE))


def jit(src, clauses, args):
    err, kid = capi.jit(src, clauses, args)
    capi.check(err)
    return kid


def family(kid):
    info = capi.nomp().nomp_b200_prog_info(kid).decode()
    d = dict(kv.split("=", 1) for kv in info.split())
    return d["kind"], d["family"]


# ---- nomp_update ----------------------------------------------------------------------------------------------------

@pytest.mark.parametrize("T", list(TYPES))
def test_update_subranges_and_errors(T):
    dt = TYPES[T][0]
    a = np.zeros(50, dtype=dt)
    for op in (capi.NOMP_FREE, capi.NOMP_FROM):
        err = capi.update(a.ctypes.data, 2, 8, a.itemsize, op)
        no, text = capi.err_info(err)
        assert no == capi.NOMP_USER_MAP_OP_IS_INVALID
        assert re.search(r"\[Error\] .*libnomp/src/nomp\.c:\d+ NOMP_FREE or NOMP_FROM can only be called on a pointer "
                         r"which is already on the device\.", text)
    for s, e in ((0, 10), (5, 10), (2, 8)):
        a[:] = 0
        a[s:e] = np.arange(s, e)
        capi.check(capi.update(a.ctypes.data, s, e, a.itemsize, capi.NOMP_TO))
        capi.check(capi.update(a.ctypes.data, s, e, a.itemsize, capi.NOMP_TO))     # twice is fine
        a[:] = 0
        capi.check(capi.update(a.ctypes.data, s, e, a.itemsize, capi.NOMP_FROM))
        capi.check(capi.update(a.ctypes.data, s, e, a.itemsize, capi.NOMP_FROM))
        want = np.zeros(50, dtype=dt)
        want[s:e] = np.arange(s, e)
        assert np.array_equal(a, want)
        capi.check(capi.update(a.ctypes.data, s, e, a.itemsize, capi.NOMP_FREE))
    # partial transfers inside a mapping: [0,20) mapped, refresh [0,5) and [15,20), read back everything
    a[:] = 0
    capi.check(capi.update(a.ctypes.data, 0, 20, a.itemsize, capi.NOMP_TO))
    a[:20] = np.arange(20)
    capi.check(capi.update(a.ctypes.data, 0, 5, a.itemsize, capi.NOMP_TO))
    capi.check(capi.update(a.ctypes.data, 15, 20, a.itemsize, capi.NOMP_TO))
    a[:] = 0
    capi.check(capi.update(a.ctypes.data, 5, 15, a.itemsize, capi.NOMP_FROM))
    assert not a.any()
    capi.check(capi.update(a.ctypes.data, 0, 20, a.itemsize, capi.NOMP_FROM))
    assert np.array_equal(a[:20], np.concatenate([np.arange(5), np.zeros(10), np.arange(15, 20)]).astype(dt))
    capi.check(capi.update(a.ctypes.data, 0, 20, a.itemsize, capi.NOMP_FREE))
    # a char buffer re-typed: the lookup is by pointer AND byte range
    raw = np.zeros(64, dtype=np.uint8)
    capi.check(capi.update(raw.ctypes.data, 0, 64, 1, capi.NOMP_ALLOC))
    typed = raw.view(dt)
    typed[:] = np.arange(typed.size)
    capi.check(capi.update(raw.ctypes.data, 0, typed.size, typed.itemsize, capi.NOMP_TO))
    typed[:] = 0
    capi.check(capi.update(raw.ctypes.data, 0, typed.size, typed.itemsize, capi.NOMP_FROM))
    assert np.array_equal(typed, np.arange(typed.size).astype(dt))
    capi.check(capi.update(raw.ctypes.data, 0, 64, 1, capi.NOMP_FREE))


def test_kernels_index_submappings_like_the_host():
    """A mapping of [lo, hi) gives the kernel the device address of host element 0, so `a[i]` means the same element
    on both sides (the reference hands over the start of the buffer)."""
    a = np.arange(40, dtype=np.float64)
    lo, hi = 7, 29
    capi.check(capi.update(a.ctypes.data, lo, hi, 8, capi.NOMP_TO))
    kid = jit("void inc(double *a, int lo, int hi) { for (int i = lo; i < hi; i++) a[i] = a[i] * 2 + 1; }",
              capi.clauses(("transform", TR, "tile")), [("a", 8, P), ("lo", 4, I), ("hi", 4, I)])
    capi.check(capi.run(kid, a.ctypes.data, C.c_int(lo), C.c_int(hi)))
    want = a.copy()
    want[lo:hi] = want[lo:hi] * 2 + 1
    a[:] = -1
    capi.check(capi.update(a.ctypes.data, lo, hi, 8, capi.NOMP_FROM))
    assert np.array_equal(a[lo:hi], want[lo:hi]) and np.all(a[:lo] == -1) and np.all(a[hi:] == -1)
    assert capi.nomp().nomp_b200_device_ptr(C.c_void_p(a.ctypes.data)) is not None
    capi.check(capi.update(a.ctypes.data, lo, hi, 8, capi.NOMP_FREE))
    assert capi.nomp().nomp_b200_device_ptr(C.c_void_p(a.ctypes.data)) is None


# ---- nomp_jit / nomp_run error paths -------------------------------------------------------------------------------------
VALID = "void foo(int *a, int N) { for (int i = 0; i < N; i++) a[i] = i; }"
ARGS = [("a", 4, P), ("N", 4, I)]


def expect(err, errno, pattern):
    no, text = capi.err_info(err)
    assert no == errno, (no, text)
    assert re.search(pattern, text), text


def test_jit_error_paths():
    err, _ = capi.jit(VALID, capi.clauses(("transform", "no_such_module", "tile")), ARGS)
    expect(err, capi.NOMP_PY_CALL_FAILURE, r'\[Error\] .*src/.*\.c:\d+ Importing Python module "no_such_module" failed\.')
    err, _ = capi.jit(VALID, capi.clauses(("transform", TR, "no_such_function")), ARGS)
    expect(err, capi.NOMP_PY_CALL_FAILURE,
           rf'\[Error\] .*src/loopy\.c:\d+ Importing Python function "no_such_function" from module "{TR}" failed\.')
    for triple in (("transform", None, "tile"), ("transform", TR, None)):
        err, _ = capi.jit(VALID, capi.clauses(triple), ARGS)
        expect(err, capi.NOMP_USER_INPUT_IS_INVALID, r"Module name and/or function name not provided\.")
    err, _ = capi.jit(VALID, capi.clauses(("invalid-clause", TR, "tile")), ARGS)
    expect(err, capi.NOMP_USER_INPUT_IS_INVALID, r'Clause "invalid-clause" passed into nomp_jit is not a valid clause\.')
    err, _ = capi.jit(VALID.replace("= i;", "= i"), capi.clauses(("transform", TR, "tile")), ARGS)
    expect(err, capi.NOMP_LOOPY_CONVERSION_FAILURE, r"Converting C source to loopy kernel failed\.")
    err, _ = capi.jit(VALID, capi.clauses(("transform", TR, "raises")), ARGS)
    expect(err, capi.NOMP_PY_CALL_FAILURE, rf'Calling Python function "raises" from module "{TR}" failed\.')
    err, _ = capi.jit(VALID, capi.clauses(("transform", TR, "returns_garbage")), ARGS)
    assert capi.err_info(err)[0] in (capi.NOMP_LOOPY_KNL_NAME_NOT_FOUND, capi.NOMP_LOOPY_CODEGEN_FAILURE, capi.NOMP_PY_CALL_FAILURE)
    err, _ = capi.jit("void f(double *a, int N, double *s) { for (int i = 0; i < N; i++) s[0] += a[i]; }",
                      capi.clauses(("reduce", "s", "/")), [("a", 8, P), ("N", 4, I), ("s", 8, F)])
    expect(err, capi.NOMP_USER_INPUT_IS_INVALID, r'Reduction operator "/" is not one of')
    # a cache hit does not even look at its arguments
    kid = C.c_int(5)
    assert capi.nomp().nomp_jit(C.byref(kid), None, None, C.c_int(0)) == 0 and kid.value == 5
    jit(VALID, capi.clauses(("transform", TR, "checks_context")), ARGS)


def test_run_error_paths():
    kid = jit("void foo(double *a, double *b, int N) { for (int i = 0; i < N; i++) a[i] = a[i] * b[i]; }",
              capi.clauses(("transform", TR, "tile")), [("a", 8, P), ("b", 8, P), ("N", 4, I)])
    a, b = np.ones(20), np.ones(20)
    expect(capi.run(-1, a.ctypes.data, b.ctypes.data, C.c_int(20)), capi.NOMP_USER_INPUT_IS_INVALID,
           r"\[Error\] .*/src/nomp\.c:\d+ Kernel id -1 passed to nomp_run is not valid\.")
    expect(capi.run(10 ** 6, a.ctypes.data, b.ctypes.data, C.c_int(20)), capi.NOMP_USER_INPUT_IS_INVALID, r"is not valid")
    with Mapped(a):
        expect(capi.run(kid, a.ctypes.data, b.ctypes.data, C.c_int(20)), capi.NOMP_USER_MAP_PTR_IS_INVALID,
               r"\[Error\] .*/src/.*\.c:\d+ Map pointer 0[xX][0-9a-fA-F]* was not found on device\.")


# ---- maps ----------------------------------------------------------------------------------------------------------------
MAP_BODIES = [
    ("a[i] += b[i];", ("native", "map")),
    ("a[i] -= b[i] + 1;", ("nvrtc", "map")),
    ("a[i] *= b[i] + 1;", ("nvrtc", "map")),
    ("a[i] = a[i] * b[i];", ("native", "map")),
    ("a[i] = a[i] * a[i] + b[i] * b[i];", ("nvrtc", "map")),
    ("a[i] = 2 * b[i] + 1;", ("nvrtc", "map")),
    ("a[i] = a[i] + b[i] + c[i];", ("nvrtc", "map")),
    ("a[i] = a[i] * b[i] + c[i];", ("nvrtc", "map")),
    ("a[i] = a[i] + 3 * b[i] + 2 * c[i];", ("nvrtc", "map")),
    ("a[i] = b[i] + c[i];", ("native", "map")),
    ("a[i] = i;", ("nvrtc", "map")),
]


@pytest.mark.parametrize("T", list(TYPES))
def test_elementwise_maps_bit_exact(T):
    dt, _ = TYPES[T]
    for body, fam in MAP_BODIES:
        src = f"void foo({T} *a, const {T} *b, const {T} *c, int N) {{ for (int i = 0; i < N; i++) {body} }}"
        kid = jit(src, capi.clauses(("transform", TR, "tile")), [("a", dt().itemsize, P), ("b", dt().itemsize, P),
                                                                 ("c", dt().itemsize, P), ("N", 4, I)])
        assert family(kid) == fam, (body, family(kid))
        for n in (10, 50, 70, 1000, 100003):       # the same kernel id serves every size (launch sizes re-evaluated)
            a, b, c = rand(T, n, 1), rand(T, n, 2), rand(T, n, 3)
            want = a.copy()
            run_kernel(src, want, b, c, n)
            with Mapped(a, b, c, out=(a,)):
                capi.check(capi.run(kid, a.ctypes.data, b.ctypes.data, c.ctypes.data, C.c_int(n)))
                capi.check(capi.nomp().nomp_sync())
            assert np.array_equal(a.view(np.uint8), want.view(np.uint8)), (T, body, n)


@pytest.mark.parametrize("T", ["double", "float", "long"])
def test_scalar_arguments_axpy_family(T):
    dt, kind = TYPES[T]
    sz = dt().itemsize
    for body, fam in (("y[i] += alpha * x[i];", ("native", "map")), ("y[i] = x[i] + alpha * y[i];", ("native", "map")),
                      ("y[i] = alpha * x[i] + beta * y[i];", ("native", "map")), ("y[i] = alpha * y[i];", ("native", "map")),
                      ("y[i] = alpha;", ("native", "map")), ("y[i] = alpha * x[i] * x[i] - beta;", ("nvrtc", "map"))):
        src = f"void k({T} *y, const {T} *x, {T} alpha, {T} beta, int N) {{ for (int i = 0; i < N; i++) {body} }}"
        kid = jit(src, capi.clauses(), [("y", sz, P), ("x", sz, P), ("alpha", sz, kind), ("beta", sz, kind), ("N", 4, I)])
        assert family(kid) == fam, body
        n = 12345
        x, y = rand(T, n, 5), rand(T, n, 6)
        alpha, beta = dt(3), dt(2)
        if T != "long":
            alpha, beta = dt(0.37), dt(-1.25)
        want = y.copy()
        ca = {"double": C.c_double, "float": C.c_float, "long": C.c_long}[T]
        run_kernel(src, want, x, ca(alpha.item()), ca(beta.item()), n)
        with Mapped(x, y, out=(y,)):
            capi.check(capi.run(kid, y.ctypes.data, x.ctypes.data, ca(alpha.item()), ca(beta.item()), C.c_int(n)))
        assert np.array_equal(y.view(np.uint8), want.view(np.uint8)), (T, body)


@pytest.mark.parametrize("T", ["int", "long", "unsigned", "unsigned long"])
def test_bitwise_operators(T):
    dt, _ = TYPES[T]
    for body in ("a[i] = a[i] & 3;", "a[i] = i | 3;", "a[i] = i ^ 3;", "a[i] = a[i] << 3;", "a[i] = a[i] >> 3;", "a[i] = ~ a[i];"):
        src = f"void foo({T} *a, int N) {{ for (int i = 0; i < N; i++) {body} }}"
        kid = jit(src, capi.clauses(("transform", TR, "tile")), [("a", dt().itemsize, P), ("N", 4, I)])
        for n in (10, 70):
            a = np.arange(n).astype(dt)
            want = a.copy()
            run_kernel(src, want, n)
            with Mapped(a, out=(a,)):
                capi.check(capi.run(kid, a.ctypes.data, C.c_int(n)))
            assert np.array_equal(a, want), (T, body)


# ---- generic loop nests ----------------------------------------------------------------------------------------------------

@pytest.mark.parametrize("T", list(TYPES))
def test_sequential_inner_loops(T):
    dt, _ = TYPES[T]
    kernels = [
        f"""void foo({T} *a, int *b, int N) {{ for (int i = 0; i < N; i++) {{ int t = 0;
              for (int j = 0; j < 10; j++) {{ if ((!(j < 3) && (j < 5)) || j == 1) continue; t += 1; }} a[i] = t; }} }}""",
        f"""void foo({T} *a, int *b, int N) {{ for (int i = 0; i < N; i++) {{ int t = 0;
              for (int j = 0; j < 10; j++) {{ t = ((!(j < 3) && (j < 5)) || j == 1) ? t : t + 1; t += (j == 4) ? 0 : 1; }} a[i] = t; }} }}""",
        f"""void foo({T} *a, int *b, int N) {{ for (int i = 0; i < N; i++) {{ int t = 0;
              for (int j = 0; j < 10; j++) {{ t += 1; if (j == 5) break; }} a[i] = t; }} }}""",
        f"""void foo({T} *a, int *b, int N) {{ for (int i = 0; i < N; i++) {{ int t = 0;
              for (int j = b[i]; j < b[i + 1] + 1; j++) {{ t += 1; }} a[i] = t; }} }}""",
    ]
    for src in kernels:
        kid = jit(src, capi.clauses(("transform", TR, "tile_outer")), [("a", dt().itemsize, P), ("b", 4, P), ("N", 4, I)])
        assert family(kid) == ("nvrtc", "generic")
        for n in (10, 50, 700):
            a = np.zeros(n, dtype=dt)
            b = (2 * np.arange(n + 1)).astype(np.int32)
            want = a.copy()
            run_kernel(src, want, b, n)
            with Mapped(a, b, out=(a,)):
                capi.check(capi.run(kid, a.ctypes.data, b.ctypes.data, C.c_int(n)))
            assert np.array_equal(a, want)


@pytest.mark.parametrize("T", list(TYPES))
def test_two_dimensional_tiling(T):
    dt, _ = TYPES[T]
    sz = dt().itemsize
    add = f"void foo({T} *a, {T} *b, int rows, int cols) {{ for (int j = 0; j < rows; j++) for (int i = 0; i < cols; i++) a[j * cols + i] = a[j * cols + i] + b[j * cols + i]; }}"
    tr = f"void foo({T} *a, {T} *b, int rows, int cols) {{ for (int j = 0; j < rows; j++) for (int i = 0; i < cols; i++) a[j + i * rows] = b[i + j * cols]; }}"
    for src in (add, tr):
        kid = jit(src, capi.clauses(("transform", TR, "tile_2d")), [("a", sz, P), ("b", sz, P), ("rows", 4, I), ("cols", 4, I)])
        for rows, cols in ((40, 5), (16, 16), (100, 37)):
            a, b = rand(T, rows * cols, 1), rand(T, rows * cols, 2)
            want = a.copy()
            run_kernel(src, want, b, rows, cols)
            with Mapped(a, b, out=(a,)):
                capi.check(capi.run(kid, a.ctypes.data, b.ctypes.data, C.c_int(rows), C.c_int(cols)))
            assert np.array_equal(a, want)
    mm = f"""void foo({T} *a, {T} *b, {T} *c, int size) {{ for (unsigned i = 0; i < size; i++) {{ for (unsigned j = 0; j < size; j++) {{
               double dot = 0; for (unsigned k = 0; k < size; k++) dot += a[i * size + k] * b[k * size + j]; c[i * size + j] = dot; }} }} }}"""
    kid = jit(mm, capi.clauses(("transform", TR, "tile_2d")), [("a", sz, P), ("b", sz, P), ("c", sz, P), ("size", 4, I)])
    for n in (10, 40):
        a = np.tile(np.arange(n), n).astype(dt)
        b = np.repeat(np.arange(n), n).astype(dt)
        c = np.zeros(n * n, dtype=dt)
        with Mapped(a, b, c, out=(c,)):
            capi.check(capi.run(kid, a.ctypes.data, b.ctypes.data, c.ctypes.data, C.c_int(n)))
        assert np.all(c == dt(sum(i * i for i in range(n))))


@pytest.mark.parametrize("T", list(TYPES))
def test_per_element_temporaries_fixed_and_jit_sized(T):
    dt, _ = TYPES[T]
    sz = dt().itemsize
    shapes = [("s[32]", "s[j]", None, 32), ("s[32][8]", "s[j][4]", None, 32), ("s[2][2][32]", "s[1][1][j]", None, 32),
              ("s[m]", "s[j]", 16, 16), ("s[m][m]", "s[j][4]", 32, 32), ("s[m][m][m]", "s[j][0][0]", 8, 8)]
    for decl, ref, m, width in shapes:
        bound = "m" if m else "32"
        src = f"""void foo({T} *b, const {T} *a, int n{', int m' if m else ''}) {{ for (int i = 0; i < n; i++) {{ {T} {decl};
              for (int j = 0; j < {bound}; j++) {ref} = a[i * {bound} + j];
              for (int j = 0; j < {bound}; j++) {ref} += {ref};
              for (int j = 0; j < {bound}; j++) b[i * {bound} + j] = {ref}; }} }}"""
        args = [("b", sz, P), ("a", sz, P), ("n", 4, I)]
        if m:
            args.append(("m", 4, I | JIT, C.c_int(m)))
        kid = jit(src, capi.clauses(("transform", TR, "element_dof")), args)
        n = 16
        a = np.repeat(np.arange(n), width).astype(dt)
        b = np.zeros(n * width, dtype=dt)
        with Mapped(a, b, out=(b,)):
            capi.check(capi.run(kid, b.ctypes.data, a.ctypes.data, C.c_int(n)))   # the JIT argument is not passed at run time
        assert np.array_equal(b, 2 * a), (T, decl)


def test_annotate_clause_goes_through_the_annotations_script():
    src = "void foo(double *a, double *b, int N) { for (int i = 0; i < N; i++) a[i] = a[i] * b[i] + i; }"
    kid = jit(src, capi.clauses(("annotate", "grid_loop", "i")), [("a", 8, P), ("b", 8, P), ("N", 4, I)])
    n = 5000
    a, b = rand("double", n, 1), rand("double", n, 2)
    want = a.copy()
    run_kernel(src, want, b, n)
    with Mapped(a, b, out=(a,)):
        capi.check(capi.run(kid, a.ctypes.data, b.ctypes.data, C.c_int(n)))
    assert np.array_equal(a, want)
    # element_loop on a kernel that is not elementwise: schedule comes from the annotation
    src2 = "void foo(double *a, const int *b, int N) { for (int i = 0; i < N; i++) a[i] = b[N - 1 - i]; }"
    kid2 = jit(src2, capi.clauses(("annotate", "element_loop", "i")), [("a", 8, P), ("b", 4, P), ("N", 4, I)])
    assert family(kid2) == ("nvrtc", "generic")
    a2, b2 = np.zeros(300), np.arange(300, dtype=np.int32)
    with Mapped(a2, b2, out=(a2,)):
        capi.check(capi.run(kid2, a2.ctypes.data, b2.ctypes.data, C.c_int(300)))
    assert np.array_equal(a2, np.arange(300)[::-1])


# ---- reductions ------------------------------------------------------------------------------------------------------------

@pytest.mark.parametrize("T", list(TYPES))
def test_reduce_clause_golden_values(T):
    """Closed forms of the reference's reduce tests; the result overwrites the output and is valid when nomp_run returns
    (no nomp_sync), the incoming value of the accumulator is ignored."""
    dt, kind = TYPES[T]
    sz = dt().itemsize
    red = capi.clauses(("reduce", "s", "+"))
    k_const = jit(f"void f({T} *s, int N) {{ for (int i = 0; i < N; i++) {{ s[0] += 1; }} }}", red, [("s", sz, kind), ("N", 4, I)])
    k_index = jit(f"void f({T} *s, int N) {{ for (int i = 0; i < N; i++) {{ s[0] += i; }} }}", red, [("s", sz, kind), ("N", 4, I)])
    k_sum = jit(f"void f({T} *a, int N, {T} *s) {{ for (int i = 0; i < N; i++) {{ s[0] += a[i]; }} }}", red,
                [("a", sz, P), ("N", 4, I), ("s", sz, kind)])
    k_dot = jit(f"void f({T} *a, {T} *b, int N, {T} *s) {{ for (int i = 0; i < N; i++) {{ s[0] += a[i] * b[i]; }} }}", red,
                [("a", 8, P), ("b", sz, P), ("N", 4, I), ("s", sz, kind)])
    assert family(k_sum) == ("native", "reduce") and family(k_dot) == ("native", "reduce") and family(k_const) == ("nvrtc", "reduce")
    for N in (10, 50):
        out = np.full(4, 77, dtype=dt)       # garbage in: must be overwritten
        capi.check(capi.run(k_const, out.ctypes.data, C.c_int(N)))
        assert out[0] == dt(N) and np.all(out[1:] == 77)
        capi.check(capi.run(k_index, out.ctypes.data, C.c_int(N)))
        assert out[0] == dt(N * (N - 1) // 2)
        a = np.arange(N).astype(dt)
        with Mapped(a):
            capi.check(capi.run(k_sum, a.ctypes.data, C.c_int(N), out.ctypes.data))
            assert out[0] == dt(N * (N - 1) // 2)
            capi.check(capi.run(k_dot, a.ctypes.data, a.ctypes.data, C.c_int(N), out.ctypes.data))
            assert out[0] == dt(N * (2 * N - 1) * (N - 1) // 6)
        for it in range(1, 5):                # the same kernel, new data, several times
            a = (it * np.arange(N)).astype(dt)
            with Mapped(a):
                capi.check(capi.run(k_sum, a.ctypes.data, C.c_int(N), out.ctypes.data))
            assert out[0] == dt((N - 1) * N * it // 2)
    capi.check(capi.run(k_const, out.ctypes.data, C.c_int(0)))      # empty loop -> identity
    assert out[0] == 0


def test_reduce_large_sizes():
    n = (1 << 24) + 11
    red = capi.clauses(("reduce", "s", "+"))
    x = ffi.fill_uniform_f64(n, 1234, 0.5, 1.5)
    y = ffi.fill_uniform_f64(n, 4321, 0.5, 1.5)
    xi = ffi.fill_i64(n, 1)
    k_sum = jit("void f(const double *a, int N, double *s) { for (int i = 0; i < N; i++) s[0] += a[i]; }", red,
                [("a", 8, P), ("N", 4, I), ("s", 8, F)])
    k_dot = jit("void f(const double *a, const double *b, int N, double *s) { for (int i = 0; i < N; i++) s[0] += a[i] * b[i]; }",
                red, [("a", 8, P), ("b", 8, P), ("N", 4, I), ("s", 8, F)])
    k_isum = jit("void f(const long *a, int N, long *s) { for (int i = 0; i < N; i++) s[0] += a[i]; }", red,
                 [("a", 8, P), ("N", 4, I), ("s", 8, I)])
    k_cond = jit("void f(const double *a, int N, double *s) { for (int i = 0; i < N; i++) { if (a[i] > 1) s[0] += 1; } }", red,
                 [("a", 8, P), ("N", 4, I), ("s", 8, F)])
    k_sq = jit("void f(const double *a, int N, double *s) { for (int i = 0; i < N; i++) s[0] += a[i] * a[i] + 1; }", red,
               [("a", 8, P), ("N", 4, I), ("s", 8, F)])
    s, si = C.c_double(), C.c_long()
    with Mapped(x, y, xi):
        capi.check(capi.run(k_sum, x.ctypes.data, C.c_int(n), s))
        ref = ffi.sum_compensated(x)
        assert abs(s.value - ref) <= 1e-12 * abs(ref)                      # fp64: reduction-order tolerance 1e-12 relative
        capi.check(capi.run(k_dot, x.ctypes.data, y.ctypes.data, C.c_int(n), s))
        ref = ffi.sum_compensated(x, y)
        assert abs(s.value - ref) <= 1e-12 * abs(ref)
        capi.check(capi.run(k_isum, xi.ctypes.data, C.c_int(n), si))
        assert si.value == ffi.reduce_(0, ffi.I64, xi)                      # int64: bit-exact
        capi.check(capi.run(k_cond, x.ctypes.data, C.c_int(n), s))
        assert s.value == float((x > 1).sum())
        capi.check(capi.run(k_sq, x.ctypes.data, C.c_int(n), s))
        ref = ffi.sum_compensated(x, x) + n
        assert abs(s.value - ref) <= 1e-12 * abs(ref)


@pytest.mark.parametrize("n", [(1 << 26) + 5, 1 << 28])
def test_reduce_clause_at_baseline_size_two_level_finish(n):
    """BASELINE configs[1]: reduce-clause sum / dot at n = 2^28, fp64 and int64, through nomp_jit / nomp_run (and the
    ragged 2^26 + 5, the smallest size on the same path: one tile per CTA, two ticket levels).  The reference cannot
    run these sizes (one partial per 512 elements into a 256 KiB scratch, ref src/reduction.c:41); what it pins are the
    closed forms of ref tests/nomp-api-500-impl.h:17-59 (sum of 1, sum of i), :87-108 (a[i] = i) and :172-196 (dot),
    evaluated here mod 2^64 for `long`.  Native kernels and the generated skeleton both take the two-level finish."""
    red = capi.clauses(("reduce", "s", "+"))
    k_sum = jit("void f(const double *a, int N, double *s) { for (int i = 0; i < N; i++) s[0] += a[i]; }", red,
                [("a", 8, P), ("N", 4, I), ("s", 8, F)])
    k_dot = jit("void f(const double *a, const double *b, int N, double *s) { for (int i = 0; i < N; i++) s[0] += a[i] * b[i]; }",
                red, [("a", 8, P), ("b", 8, P), ("N", 4, I), ("s", 8, F)])
    k_isum = jit("void f(const long *a, int N, long *s) { for (int i = 0; i < N; i++) s[0] += a[i]; }", red,
                 [("a", 8, P), ("N", 4, I), ("s", 8, I)])
    k_idot = jit("void f(const long *a, const long *b, int N, long *s) { for (int i = 0; i < N; i++) s[0] += a[i] * b[i]; }", red,
                 [("a", 8, P), ("b", 8, P), ("N", 4, I), ("s", 8, I)])
    k_one = jit("void f(long *s, int N) { for (int i = 0; i < N; i++) { s[0] += 1; } }", red, [("s", 8, I), ("N", 4, I)])
    k_idx = jit("void f(long *s, int N) { for (int i = 0; i < N; i++) { s[0] += i; } }", red, [("s", 8, I), ("N", 4, I)])
    k_sq = jit("void f(const double *a, int N, double *s) { for (int i = 0; i < N; i++) s[0] += a[i] * a[i] + 1; }", red,
               [("a", 8, P), ("N", 4, I), ("s", 8, F)])
    assert family(k_sum) == family(k_dot) == family(k_isum) == family(k_idot) == ("native", "reduce")
    assert family(k_one) == family(k_idx) == family(k_sq) == ("nvrtc", "reduce")
    s, si = C.c_double(), C.c_long()
    wrap = lambda v: (v + (1 << 63)) % (1 << 64) - (1 << 63)  # noqa: E731

    capi.check(capi.run(k_one, si, C.c_int(n)))
    assert si.value == n
    capi.check(capi.run(k_idx, si, C.c_int(n)))
    assert si.value == n * (n - 1) // 2

    x = ffi.fill_int_f64(n, 1, 0, 7)         # Set X: every order exact -> bitwise
    y = ffi.fill_int_f64(n, 2, 0, 7)
    with Mapped(x, y):
        capi.check(capi.run(k_sum, x.ctypes.data, C.c_int(n), s))
        assert s.value == ffi.reduce_(0, ffi.F64, x)
        capi.check(capi.run(k_dot, x.ctypes.data, y.ctypes.data, C.c_int(n), s))
        assert s.value == ffi.reduce_(0, ffi.F64, x, y)
        capi.check(capi.run(k_sq, x.ctypes.data, C.c_int(n), s))
        assert s.value == ffi.reduce_(0, ffi.F64, x, x) + n
    x = ffi.fill_uniform_f64(n, 1234, 0.5, 1.5)      # Set R: 1e-12 relative against the compensated oracle
    y = ffi.fill_uniform_f64(n, 4321, 0.5, 1.5)
    with Mapped(x, y):
        capi.check(capi.run(k_sum, x.ctypes.data, C.c_int(n), s))
        ref = ffi.sum_compensated(x)
        assert abs(s.value - ref) <= 1e-12 * abs(ref)
        capi.check(capi.run(k_dot, x.ctypes.data, y.ctypes.data, C.c_int(n), s))
        ref = ffi.sum_compensated(x, y)
        assert abs(s.value - ref) <= 1e-12 * abs(ref)
    del x, y
    xi = ffi.fill_i64(n, 1)                          # full-range int64: bit-exact
    yi = ffi.fill_i64(n, 2)
    with Mapped(xi, yi):
        capi.check(capi.run(k_isum, xi.ctypes.data, C.c_int(n), si))
        assert si.value == ffi.reduce_(0, ffi.I64, xi)
        capi.check(capi.run(k_idot, xi.ctypes.data, yi.ctypes.data, C.c_int(n), si))
        assert si.value == ffi.reduce_(0, ffi.I64, xi, yi)
    del yi
    xi[:] = np.arange(n, dtype=np.int64)             # the reference's own fixture a[i] = i and its closed forms
    with Mapped(xi):
        capi.check(capi.run(k_isum, xi.ctypes.data, C.c_int(n), si))
        assert si.value == n * (n - 1) // 2
        capi.check(capi.run(k_idot, xi.ctypes.data, xi.ctypes.data, C.c_int(n), si))
        assert si.value == wrap(n * (2 * n - 1) * (n - 1) // 6)


@pytest.mark.parametrize("T", ["int", "unsigned long", "double", "float"])
def test_min_max_and_product_clauses(T):
    dt, kind = TYPES[T]
    sz = dt().itemsize
    n = 100003
    a = rand(T, n, 9)
    kmin = jit(f"void f(const {T} *a, int N, {T} *m) {{ for (int i = 0; i < N; i++) m[0] = (a[i] < m[0]) ? a[i] : m[0]; }}",
               capi.clauses(("reduce", "m", "min")), [("a", sz, P), ("N", 4, I), ("m", sz, kind)])
    kmax = jit(f"void f(const {T} *a, int N, {T} *m) {{ for (int i = 0; i < N; i++) m[0] = (m[0] > a[i]) ? m[0] : a[i]; }}",
               capi.clauses(("reduce", "m", "max")), [("a", sz, P), ("N", 4, I), ("m", sz, kind)])
    kprod = jit(f"void f(const {T} *a, int N, {T} *m) {{ for (int i = 0; i < N; i++) m[0] *= a[i]; }}",
                capi.clauses(("reduce", "m", "*")), [("a", sz, P), ("N", 4, I), ("m", sz, kind)])
    assert family(kmin) == ("native", "reduce")
    out = np.zeros(1, dtype=dt)
    with Mapped(a):
        capi.check(capi.run(kmin, a.ctypes.data, C.c_int(n), out.ctypes.data))
        assert out[0] == a.min()
        capi.check(capi.run(kmax, a.ctypes.data, C.c_int(n), out.ctypes.data))
        assert out[0] == a.max()
    small = (np.arange(12) % 3 + 1).astype(dt)
    with Mapped(small):
        capi.check(capi.run(kprod, small.ctypes.data, C.c_int(12), out.ctypes.data))
    assert out[0] == dt(np.prod(small.astype(np.float64)))


# ---- Ax through the API ------------------------------------------------------------------------------------------------------

@pytest.mark.parametrize("n,expect_family", [(8, ("native", "ax")), (10, ("native", "ax")), (6, ("native", "ax")), (12, ("native", "ax")),
                                             (4, ("nvrtc", "generic")), (5, ("nvrtc", "generic")), (9, ("nvrtc", "generic")),
                                             (11, ("nvrtc", "generic"))])
def test_ax_kernel_string(n, expect_family):
    from nomp_bridge.families import AX_KERNEL_SOURCE
    E = 37
    u = ffi.fill_int_f64(E * n ** 3, 5, -4, 4)
    g = ffi.fill_int_f64(E * 6 * n ** 3, 6, 0, 3)
    D = ffi.fill_int_f64(n * n, 7, -2, 2)
    w = np.zeros_like(u)
    kid = jit(AX_KERNEL_SOURCE, capi.clauses(), [("w", 8, P), ("u", 8, P), ("g", 8, P), ("D", 8, P), ("E", 4, I),
                                                 ("n", 4, I | JIT, C.c_int(n))])
    assert family(kid) == expect_family
    with Mapped(u, g, D, w, out=(w,)):
        capi.check(capi.run(kid, w.ctypes.data, u.ctypes.data, g.ctypes.data, D.ctypes.data, C.c_int(E)))
        first = w.copy()
        # D changes on the device: the cached __constant__ copy must be refreshed
        D2 = D * 2
        D[:] = D2
        capi.check(capi.update(D.ctypes.data, 0, D.size, 8, capi.NOMP_TO))
        capi.check(capi.run(kid, w.ctypes.data, u.ctypes.data, g.ctypes.data, D.ctypes.data, C.c_int(E)))
        capi.check(capi.update(w.ctypes.data, 0, w.size, 8, capi.NOMP_FROM))
        assert np.array_equal(w, ffi.ax(n, u, g, D))
        capi.check(capi.update(first.ctypes.data, 0, 1, 8, capi.NOMP_ALLOC))
        capi.check(capi.update(first.ctypes.data, 0, 1, 8, capi.NOMP_FREE))
    assert np.array_equal(w, ffi.ax(n, u, g, D))
    assert np.array_equal(w, 4 * ffi.ax(n, u, g, D / 2))


def test_every_spelling_of_ax_runs_the_native_kernel():
    """Six spellings of the operator that share no token sequence with the canonical string (tests/ax_variants.py: other
    names and argument order, g[e][f][k][j][i], merged / flat / extra temporaries, loops and statements reordered, `+=`
    into a zeroed w) and two of Ax + p.Ap under a reduce clause: nomp_jit routes every one to the hand-written kernel
    (structural recogniser, nomp_bridge/axprobe.py) and nomp_run gives the oracle's bits; near misses run through the
    generic path and give what their own C text gives."""
    from tests import ax_variants as V
    from tests.test_frontend import _signature_order
    n, E = 8, 41
    u = ffi.fill_int_f64(E * n ** 3, 5, -4, 4)
    g = ffi.fill_int_f64(E * 6 * n ** 3, 6, 0, 3)
    D = ffi.fill_int_f64(n * n, 7, -2, 2)
    want = ffi.ax(n, u, g, D)
    for name, src, roles in V.variants() + V.fused_variants():
        fused = len(roles) == 7
        by_name = dict(zip(roles, ("w", "u", "g", "D", "E", "n", "pap")))
        order = _signature_order(src)
        spec = {"w": 8, "u": 8, "g": 8, "D": 8}
        args = []
        for prm in order:
            role = by_name[prm]
            if role in spec:
                args.append((prm, 8, P))
            elif role == "E":
                args.append((prm, 4, I))
            elif role == "n":
                args.append((prm, 4, I | JIT, C.c_int(n)))
            else:
                args.append((prm, 8, F))
        kid = jit(src, capi.clauses(("reduce", roles[6], "+")) if fused else capi.clauses(), args)
        assert family(kid) == ("native", "axdot" if fused else "ax"), name
        w = np.full_like(u, 7.0)                      # garbage in: the operator overwrites
        pap = C.c_double(-1.0)
        values = {"w": w.ctypes.data, "u": u.ctypes.data, "g": g.ctypes.data, "D": D.ctypes.data, "E": C.c_int(E), "pap": pap}
        with Mapped(u, g, D, w, out=(w,)):
            capi.check(capi.run(kid, *[values[by_name[prm]] for prm in order if by_name[prm] != "n"]))
        assert np.array_equal(w, want), name
        if fused:
            assert pap.value == float(u @ want), name
    for name, src in V.not_ax()[:4]:
        order = _signature_order(src)
        kid = jit(src, capi.clauses(), [(prm, 8, P) if prm in ("w", "u", "g", "D", "out", "in", "dm") else
                                        (prm, 4, I | JIT, C.c_int(n)) if prm == "n" else (prm, 4, I) for prm in order])
        assert family(kid)[0] == "nvrtc", name
        w0 = ffi.fill_int_f64(u.size, 9, -3, 3)
        w = w0.copy()
        from tests.test_frontend import _run_variant_with_gcc  # noqa: F401  (same helper module)
        values = {"w": w.ctypes.data, "u": u.ctypes.data, "g": g.ctypes.data, "D": D.ctypes.data, "E": C.c_int(E)}
        with Mapped(u, g, D, w, out=(w,)):
            capi.check(capi.run(kid, *[values[prm] for prm in order if prm != "n"]))
        ref = w0.copy()
        run_kernel(src, *[{"w": ref, "u": u, "g": g, "D": D, "E": E, "n": n}[prm] for prm in order])
        assert np.array_equal(w, ref) and not np.array_equal(w, want), name


@pytest.mark.parametrize("n", [6, 8, 10, 12])
def test_ax_closed_form_on_a_sheared_element(n):
    """The hand-written Ax kernels against an analytic known answer (tests/ax_closed_form.py: harmonic polynomial on a
    sheared element with a full constant metric -- boundary fluxes, zero inside), the real GLL matrix, several copies of
    the element so that every lane group of the kernel sees it; also with the fused p.Ap, whose closed form is the sum of
    u times the answer."""
    from nomp_bridge.families import AX_DOT_KERNEL_SOURCE, AX_KERNEL_SOURCE
    from tests.ax_closed_form import sheared_element
    Dm, x = ffi.gll_derivative(n)
    u1, g1, want1, interior1 = sheared_element(n, x)
    E = 23
    u, g, want, interior = np.tile(u1, E), np.tile(g1, E), np.tile(want1, E), np.tile(interior1, E)
    D = np.ascontiguousarray(Dm.ravel())
    w = np.full_like(u, np.nan)
    pap = C.c_double(-1.0)
    args = [("w", 8, P), ("u", 8, P), ("g", 8, P), ("D", 8, P), ("E", 4, I), ("n", 4, I | JIT, C.c_int(n))]
    kid = jit(AX_KERNEL_SOURCE, capi.clauses(), args)
    kdot = jit(AX_DOT_KERNEL_SOURCE, capi.clauses(("reduce", "pap", "+")), args + [("pap", 8, F)])
    assert family(kid) == ("native", "ax") and family(kdot) == ("native", "axdot")
    scale = np.abs(want).max()
    with Mapped(u, g, D, w, out=(w,)):
        capi.check(capi.run(kid, w.ctypes.data, u.ctypes.data, g.ctypes.data, D.ctypes.data, C.c_int(E)))
        capi.check(capi.update(w.ctypes.data, 0, w.size, 8, capi.NOMP_FROM))
        assert np.abs(w - want).max() <= 1e-11 * scale
        assert np.abs(w[interior]).max() <= 1e-11 * scale
        w[:] = np.nan
        capi.check(capi.update(w.ctypes.data, 0, w.size, 8, capi.NOMP_TO))
        capi.check(capi.run(kdot, w.ctypes.data, u.ctypes.data, g.ctypes.data, D.ctypes.data, C.c_int(E), pap))
    assert np.abs(w - want).max() <= 1e-11 * scale
    assert abs(pap.value - float(u @ want)) <= 1e-10 * abs(float(u @ want))


@pytest.mark.parametrize("n", [8, 10, 12])
def test_ax_takes_the_even_odd_path_for_an_antisymmetric_D(n):
    """The backend reads D from the device once per version: a centro-antisymmetric D (every GLL matrix) runs the even-odd
    kernels -- the bits of nompk_ax_f64(..., NOMPK_AX_D_ANTISYMMETRIC) --, any other D the general ones; the decision
    follows nomp_update(D, TO)."""
    import torch
    from nomp_bridge.families import AX_KERNEL_SOURCE
    lib = capi.nompk()
    E = 61
    u = ffi.fill_uniform_f64(E * n ** 3, 7, 0.5, 1.5)
    g = ffi.fill_uniform_f64(E * 6 * n ** 3, 8, 0.5, 1.5)
    Dgll = np.ascontiguousarray(ffi.gll_derivative(n)[0].ravel())
    Dany = Dgll.copy()
    Dany[1] += 0.25
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def direct(D, flags):
        tu, tg, tD = (torch.from_numpy(a).cuda() for a in (u, g, D))
        tw = torch.empty_like(tu)
        capi.nompk_check(lib.nompk_ax_f64(n, E, tu.data_ptr(), tg.data_ptr(), tD.data_ptr(), tw.data_ptr(), flags, st))
        torch.cuda.synchronize()
        return tw.cpu().numpy()

    eo, general, other = direct(Dgll, 2), direct(Dgll, 0), direct(Dany, 0)
    assert not np.array_equal(eo, general)
    D = Dgll.copy()
    w = np.zeros_like(u)
    kid = jit(AX_KERNEL_SOURCE, capi.clauses(), [("w", 8, P), ("u", 8, P), ("g", 8, P), ("D", 8, P), ("E", 4, I),
                                                 ("n", 4, I | JIT, C.c_int(n))])
    with Mapped(u, g, D, w, out=(w,)):
        for want, Dnew in ((eo, None), (other, Dany), (eo, Dgll)):
            if Dnew is not None:
                D[:] = Dnew
                capi.check(capi.update(D.ctypes.data, 0, D.size, 8, capi.NOMP_TO))
            for _ in range(2):            # the second run uses the cached decision and the staged tables
                capi.check(capi.run(kid, w.ctypes.data, u.ctypes.data, g.ctypes.data, D.ctypes.data, C.c_int(E)))
            capi.check(capi.update(w.ctypes.data, 0, w.size, 8, capi.NOMP_FROM))
            assert np.array_equal(w, want)


def test_fused_cg_kernels():
    """Row (f): Ax fused with p.Ap through the canonical Ax+dot kernel string, and the fused CG update
    x += a p; r -= a w; rr = r.r through the reduce skeleton (elementwise writes in front of the accumulation)."""
    from nomp_bridge.families import AX_DOT_KERNEL_SOURCE
    n, E = 8, 53
    u = ffi.fill_int_f64(E * n ** 3, 5, -4, 4)
    g = ffi.fill_int_f64(E * 6 * n ** 3, 6, 0, 3)
    D = ffi.fill_int_f64(n * n, 7, -2, 2)
    w = np.zeros_like(u)
    kid = jit(AX_DOT_KERNEL_SOURCE, capi.clauses(("reduce", "pap", "+")),
              [("w", 8, P), ("u", 8, P), ("g", 8, P), ("D", 8, P), ("E", 4, I), ("n", 4, I | JIT, C.c_int(n)), ("pap", 8, F)])
    assert family(kid) == ("native", "axdot")
    pap = C.c_double(-1)
    with Mapped(u, g, D, w, out=(w,)):
        for _ in range(2):
            capi.check(capi.run(kid, w.ctypes.data, u.ctypes.data, g.ctypes.data, D.ctypes.data, C.c_int(E), pap))
    want = ffi.ax(n, u, g, D)
    assert np.array_equal(w, want) and pap.value == float(u @ want)
    err, _ = capi.jit(AX_DOT_KERNEL_SOURCE, capi.clauses(("reduce", "pap", "+")),
                      [("w", 8, P), ("u", 8, P), ("g", 8, P), ("D", 8, P), ("E", 4, I), ("n", 4, I | JIT, C.c_int(9)), ("pap", 8, F)])
    assert capi.err_info(err)[0] in (capi.NOMP_LOOPY_CODEGEN_FAILURE, capi.NOMP_PY_CALL_FAILURE)

    src = """void upd(double *x, double *r, const double *p, const double *w, double alpha, int N, double *rr) {
               for (int i = 0; i < N; i++) { x[i] += alpha * p[i]; r[i] -= alpha * w[i]; rr[0] += r[i] * r[i]; } }"""
    kid = jit(src, capi.clauses(("reduce", "rr", "+")), [("x", 8, P), ("r", 8, P), ("p", 8, P), ("w", 8, P), ("alpha", 8, F),
                                                        ("N", 4, I), ("rr", 8, F)])
    assert family(kid) == ("nvrtc", "reduce")
    N = 100003
    x, r, p, ww = (rand("double", N, s) for s in (1, 2, 3, 4))
    xr, rr_ = x.copy(), r.copy()
    rr_ref = np.zeros(1)
    run_kernel(src, xr, rr_, p, ww, 0.37, N, rr_ref)
    rr = C.c_double()
    with Mapped(x, r, p, ww, out=(x, r)):
        capi.check(capi.run(kid, x.ctypes.data, r.ctypes.data, p.ctypes.data, ww.ctypes.data, C.c_double(0.37), C.c_int(N), rr))
    assert np.array_equal(x, xr) and np.array_equal(r, rr_)
    assert abs(rr.value - ffi.sum_compensated(rr_, rr_)) <= 1e-12 * rr.value
    err, _ = capi.jit("void bad(double *x, int N, double *s) { for (int i = 0; i < N; i++) { x[0] = i; s[0] += x[i]; } }",
                      capi.clauses(("reduce", "s", "+")), [("x", 8, P), ("N", 4, I), ("s", 8, F)])
    assert capi.err_info(err)[0] == capi.NOMP_PY_CALL_FAILURE       # non-elementwise write in a reduction kernel


def test_async_update_pipeline():
    """nomp_b200_update_async: H2D and D2H on their own streams, ordered against the kernels; data valid after nomp_sync."""
    n, nblk = 1 << 20, 4
    pin = (lambda t: t) if os.environ.get("NOMP_HOSTDEV_ACTIVE") == "1" else (lambda t: t.pin_memory())   # tests/hostdev has no driver
    x = pin(torch.arange(n * nblk, dtype=torch.float64))
    y = pin(torch.zeros(n * nblk, dtype=torch.float64))
    kid = jit("void k(double *y, const double *x, int N) { for (int i = 0; i < N; i++) y[i] = 2 * x[i] + 1; }", capi.clauses(),
              [("y", 8, P), ("x", 8, P), ("N", 4, I)])
    blocks = [(x.data_ptr() + b * n * 8, y.data_ptr() + b * n * 8) for b in range(nblk)]
    err = capi.update_async(blocks[0][0], 0, n, 8, capi.NOMP_TO)
    assert capi.err_info(err)[0] == capi.NOMP_USER_MAP_OP_IS_INVALID          # not mapped yet
    for xb, yb in blocks:
        capi.check(capi.update(xb, 0, n, 8, capi.NOMP_ALLOC))
        capi.check(capi.update(yb, 0, n, 8, capi.NOMP_ALLOC))
    for rep in range(3):
        x.add_(1.0)
        for xb, yb in blocks:
            capi.check(capi.update_async(xb, 0, n, 8, capi.NOMP_TO))
            capi.check(capi.run(kid, yb, xb, C.c_int(n)))
            capi.check(capi.update_async(yb, 0, n, 8, capi.NOMP_FROM))
        capi.check(capi.nomp().nomp_sync())
        assert torch.equal(y, 2 * x + 1)
    assert capi.err_info(capi.update_async(blocks[0][0], 0, n, 8, capi.NOMP_FREE))[0] == capi.NOMP_USER_MAP_OP_IS_INVALID
    for xb, yb in blocks:
        capi.check(capi.update(xb, 0, n, 8, capi.NOMP_FREE))
        capi.check(capi.update(yb, 0, n, 8, capi.NOMP_FREE))


def test_async_copy_out_is_ordered_before_later_writers():
    """nomp_b200_update_async(FROM) returns at once; whatever writes the mapping afterwards -- a kernel with a non-const
    pointer to it, a blocking nomp_update(TO) -- must wait for the copy, which would otherwise deliver new or torn data."""
    n = 1 << 23                                            # 64 MiB: the copy takes milliseconds, a kernel microseconds
    pin = (lambda t: t) if os.environ.get("NOMP_HOSTDEV_ACTIVE") == "1" else (lambda t: t.pin_memory())
    y = pin(torch.zeros(n, dtype=torch.float64))
    fresh = np.full(n, 7.0)
    inc = jit("void inc(double *y, int N) { for (int i = 0; i < N; i++) y[i] += 1; }", capi.clauses(), [("y", 8, P), ("N", 4, I)])
    capi.check(capi.update(y.data_ptr(), 0, n, 8, capi.NOMP_TO))
    for rep in range(3):
        capi.check(capi.run(inc, y.data_ptr(), C.c_int(n)))                     # device: rep + 1
        capi.check(capi.update_async(y.data_ptr(), 0, n, 8, capi.NOMP_FROM))   # host gets rep + 1 ...
        capi.check(capi.run(inc, y.data_ptr(), C.c_int(n)))                     # ... although the device moves on at once
        capi.check(capi.run(inc, y.data_ptr(), C.c_int(n)))
        capi.check(capi.nomp().nomp_sync())
        assert torch.all(y == 3.0 * rep + 1.0), (rep, y[:4], y[-4:])
    capi.check(capi.update_async(y.data_ptr(), 0, n, 8, capi.NOMP_FROM))       # host gets 9 ...
    capi.check(capi.nomp().nomp_sync())
    assert torch.all(y == 9.0)
    capi.check(capi.update(y.data_ptr(), 0, n, 8, capi.NOMP_FREE))              # FREE with copies possibly in flight
    capi.check(capi.update(fresh.ctypes.data, 0, n, 8, capi.NOMP_TO))
    capi.check(capi.update(fresh.ctypes.data, 0, n, 8, capi.NOMP_FREE))


def test_extensions_and_launch_counter():
    lib = capi.nomp()
    assert lib.nomp_b200_stream() is not None
    assert lib.nomp_b200_comm_size() == 1 and lib.nomp_b200_comm_rank() == 0
    before = lib.nomp_b200_launch_count()
    a, b = np.ones(1000), np.ones(1000)
    kid = jit("void foo(double *a, double *b, int N) { for (int i = 0; i < N; i++) a[i] += b[i]; }", capi.clauses(),
              [("a", 8, P), ("b", 8, P), ("N", 4, I)])
    with Mapped(a, b, out=(a,)):
        for _ in range(3):
            capi.check(capi.run(kid, a.ctypes.data, b.ctypes.data, C.c_int(1000)))
    assert lib.nomp_b200_launch_count() - before == 3 and np.all(a == 4)


def test_repeated_updates_pin_the_host_range():
    """A mapping that is copied a second time gets its host range page-locked (cudaHostRegister) so that the copies run
    at PCIe speed instead of through the driver's staging buffers; NOMP_FREE releases it.  Small ranges are left alone."""
    import time
    cudart = C.CDLL("libcudart.so.12")
    cudart.cudaHostGetFlags.argtypes = [C.POINTER(C.c_uint), C.c_void_p]

    def pinned(arr):
        flags = C.c_uint()
        rc = cudart.cudaHostGetFlags(C.byref(flags), arr.ctypes.data)
        cudart.cudaGetLastError()
        return rc == 0

    a = np.random.default_rng(1).random(1 << 25)        # 256 MiB of pageable memory
    want = a.copy()
    times = []
    for i in range(5):
        t0 = time.perf_counter()
        capi.check(capi.update(a.ctypes.data, 0, a.size, 8, capi.NOMP_TO))
        times.append(time.perf_counter() - t0)
        assert pinned(a) == (i >= 1), i                   # from the second copy on
    a[:] = 0
    capi.check(capi.update(a.ctypes.data, 0, a.size, 8, capi.NOMP_FROM))
    assert np.array_equal(a, want)
    capi.check(capi.update(a.ctypes.data, 0, a.size, 8, capi.NOMP_FREE))
    assert not pinned(a)
    gbs = [a.nbytes / t / 1e9 for t in times]
    print("H2D GB/s per call (pageable, registering, pinned...):", [round(g, 1) for g in gbs])
    assert max(gbs[2:]) > 0.9 * gbs[0]
    small = np.arange(1000.0)
    for _ in range(3):
        capi.check(capi.update(small.ctypes.data, 0, small.size, 8, capi.NOMP_TO))
    assert not pinned(small)
    capi.check(capi.update(small.ctypes.data, 0, small.size, 8, capi.NOMP_FREE))
