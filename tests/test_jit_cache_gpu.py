"""On-disk JIT cache (libnomp_b200/csrc/libnomp/src/jitcache.c): a second process start
serves nomp_jit() from disk -- the bridge's output (.knl) and NVRTC's CUBIN (.cubin) -- with identical results, and
anything that can change a kernel (script text, clauses, NOMP_JIT values) changes the key."""
import ctypes as C
import os
import shutil
import sys
import time
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "libnomp_b200" / "python"))

from libnomp_b200 import capi  # noqa: E402
from nomp_bridge.families import AX_KERNEL_SOURCE  # noqa: E402
from oracle import ffi  # noqa: E402

P, I, F = capi.NOMP_PTR, capi.NOMP_INT, capi.NOMP_FLOAT
ADD = "void add(double *a, const double *b, int N) { for (int i = 0; i < N; i++) a[i] += b[i]; }"
SQ = "void sq(double *a, const double *b, int N) { for (int i = 0; i < N; i++) a[i] = a[i] * a[i] + b[i] * b[i]; }"
DOT = "void dot(const double *x, const double *y, int N, double *s) { for (int i = 0; i < N; i++) s[0] += x[i] * y[i]; }"
TR = None  # set by the fixture
COND = ("void cnt(const double *x, int N, double *s) { for (int i = 0; i < N; i++) if (x[i] > 3) s[0] += x[i]; }")


@pytest.fixture()
def cache(tmp_path, monkeypatch):
    scripts = tmp_path / "scripts"
    shutil.copytree(ROOT / "tests" / "scripts", scripts)
    # a module name of its own per test: the interpreter outlives the test and keeps imported scripts in sys.modules
    global TR
    TR = f"cache_tr_{abs(hash(str(tmp_path))) % 10 ** 8}"
    (scripts / f"{TR}.py").write_text((scripts / "nomp_test_transforms.py").read_text())
    monkeypatch.setenv("NOMP_JIT_CACHE", "1")
    monkeypatch.setenv("NOMP_JIT_CACHE_DIR", str(tmp_path / "jit"))
    return tmp_path


def session(cache, body):
    """One nomp_init .. nomp_finalize cycle; returns (body's result, change of the cache counters)."""
    before = capi.jit_cache_stats()
    capi.check(capi.init(backend="cuda", device=0, verbose=0, scripts_dir=cache / "scripts",
                         annotations_script="nomp_test_annotations"))
    try:
        out = body()
    finally:
        assert capi.nomp().nomp_finalize_excluding_interpreter() == 0
    after = capi.jit_cache_stats()
    return out, {k: after[k] - before[k] for k in after}


def workload():
    n = 20011
    a = ffi.fill_uniform_f64(n, 1, 0.5, 1.5)
    b = ffi.fill_uniform_f64(n, 2, 0.5, 1.5)
    x = ffi.fill_int_f64(n, 3, 0, 7)
    for arr in (a, b, x):
        capi.check(capi.update(arr.ctypes.data, 0, n, 8, capi.NOMP_TO))
    times = {}

    def jit(label, src, clauses, args):
        t0 = time.perf_counter()
        err, kid = capi.jit(src, clauses, args)
        times[label] = time.perf_counter() - t0
        capi.check(err)
        return kid

    vec = [("a", 8, P), ("b", 8, P), ("N", 4, I)]
    k_add = jit("add", ADD, capi.clauses(), vec)                                        # native family
    k_sq = jit("sq", SQ, capi.clauses(("transform", TR, "tile")), vec)          # user schedule, NVRTC
    k_dot = jit("dot", DOT, capi.clauses(("reduce", "s", "+")),
                [("x", 8, P), ("y", 8, P), ("N", 4, I), ("s", 8, F)])                   # native reduction
    k_cnt = jit("cnt", COND, capi.clauses(("reduce", "s", "+")), [("x", 8, P), ("N", 4, I), ("s", 8, F)])  # NVRTC reduce
    infos = [capi.nomp().nomp_b200_prog_info(k).decode() for k in (k_add, k_sq, k_dot, k_cnt)]
    capi.check(capi.run(k_add, a.ctypes.data, b.ctypes.data, C.c_int(n)))
    capi.check(capi.run(k_sq, a.ctypes.data, b.ctypes.data, C.c_int(n)))
    s, c = C.c_double(-1), C.c_double(-1)
    capi.check(capi.run(k_dot, x.ctypes.data, x.ctypes.data, C.c_int(n), s))
    capi.check(capi.run(k_cnt, x.ctypes.data, C.c_int(n), c))
    capi.check(capi.update(a.ctypes.data, 0, n, 8, capi.NOMP_FROM))
    for arr in (a, b, x):
        capi.check(capi.update(arr.ctypes.data, 0, n, 8, capi.NOMP_FREE))
    return a, s.value, c.value, infos, times


def test_second_start_is_served_from_disk(cache):
    (a1, s1, c1, info1, t1), d1 = session(cache, workload)
    assert d1 == {"knl_hits": 0, "knl_misses": 4, "cubin_hits": 0, "cubin_misses": 2}
    files = sorted(p.suffix for p in (cache / "jit").iterdir())
    assert files == [".cubin"] * 2 + [".knl"] * 4
    (a2, s2, c2, info2, t2), d2 = session(cache, workload)
    assert d2 == {"knl_hits": 4, "knl_misses": 0, "cubin_hits": 2, "cubin_misses": 0}
    assert np.array_equal(a1, a2) and s1 == s2 and c1 == c2 and info1 == info2
    x = ffi.fill_int_f64(20011, 3, 0, 7)
    assert s1 == float((x * x).sum()) and c1 == float(x[x > 3].sum())
    # a warm start does not enter the interpreter or NVRTC: well under the cold cost of every kernel
    assert all(t2[k] < 0.5 * t1[k] for k in t1), (t1, t2)
    print("cold jit (ms):", {k: round(1e3 * v, 2) for k, v in t1.items()},
          "warm:", {k: round(1e3 * v, 3) for k, v in t2.items()})


def test_key_covers_script_text_clauses_and_jit_values(cache):
    vec = [("a", 8, P), ("b", 8, P), ("N", 4, I)]

    def one(clauses, src=SQ):
        def body():
            err, kid = capi.jit(src, clauses, vec)
            capi.check(err)
            return capi.nomp().nomp_b200_prog_info(kid).decode()
        return session(cache, body)

    _, d = one(capi.clauses(("transform", TR, "tile")))
    assert d["knl_misses"] == 1
    _, d = one(capi.clauses(("transform", TR, "tile")))
    assert d["knl_hits"] == 1
    # another function of the same script, another script text, another clause set: all new programs
    _, d = one(capi.clauses(("transform", TR, "tile_small")))
    assert d["knl_misses"] == 1 and d["knl_hits"] == 0
    with open(cache / "scripts" / f"{TR}.py", "a") as fp:
        fp.write("\n# edited\n")
    _, d = one(capi.clauses(("transform", TR, "tile")))
    assert d["knl_misses"] == 1 and d["knl_hits"] == 0
    # ... but the generated source is the same, so NVRTC's output is reused
    assert d["cubin_hits"] == 1 and d["cubin_misses"] == 0
    _, d = one(capi.clauses(("annotate", "grid_loop", "i")))
    assert d["knl_misses"] == 1 and d["knl_hits"] == 0

    # NOMP_JIT values are part of the key: n = 8 and n = 10 of the Ax kernel are different programs
    def ax(n):
        def body():
            err, kid = capi.jit(AX_KERNEL_SOURCE, capi.clauses(),
                                [("w", 8, P), ("u", 8, P), ("g", 8, P), ("D", 8, P), ("E", 4, I),
                                 ("n", 4, I | capi.NOMP_JIT, C.c_int(n))])
            capi.check(err)
            return capi.nomp().nomp_b200_prog_info(kid).decode()
        return session(cache, body)

    i8, d = ax(8)
    assert d["knl_misses"] == 1 and "n=8" in i8
    i10, d = ax(10)
    assert d["knl_misses"] == 1 and "n=10" in i10
    i8b, d = ax(8)
    assert d["knl_hits"] == 1 and i8b == i8


def test_damaged_entries_are_rebuilt_and_failures_are_not_cached(cache):
    vec = [("a", 8, P), ("b", 8, P), ("N", 4, I)]

    def body():
        err, kid = capi.jit(SQ, capi.clauses(("transform", TR, "tile")), vec)
        capi.check(err)
        n = 1000
        a, b = np.full(n, 2.0), np.full(n, 3.0)
        capi.check(capi.update(a.ctypes.data, 0, n, 8, capi.NOMP_TO))
        capi.check(capi.update(b.ctypes.data, 0, n, 8, capi.NOMP_TO))
        capi.check(capi.run(kid, a.ctypes.data, b.ctypes.data, C.c_int(n)))
        capi.check(capi.update(a.ctypes.data, 0, n, 8, capi.NOMP_FROM))
        capi.check(capi.update(a.ctypes.data, 0, n, 8, capi.NOMP_FREE))
        capi.check(capi.update(b.ctypes.data, 0, n, 8, capi.NOMP_FREE))
        return a

    a, d = session(cache, body)
    assert np.all(a == 13.0) and d["knl_misses"] == 1 and d["cubin_misses"] == 1
    for p in (cache / "jit").iterdir():
        data = p.read_bytes()
        if p.suffix == ".cubin":     # truncated, and (second variant below) one flipped byte in the middle
            p.write_bytes(data[: len(data) // 2])
        else:
            p.write_bytes(b"NOMPJIT1\ngarbage")
    a, d = session(cache, body)
    assert np.all(a == 13.0) and d == {"knl_hits": 0, "knl_misses": 1, "cubin_hits": 0, "cubin_misses": 1}
    a, d = session(cache, body)
    assert np.all(a == 13.0) and d["knl_hits"] == 1 and d["cubin_hits"] == 1
    for p in (cache / "jit").iterdir():            # same length, one byte changed: the checksum trailer catches it
        data = bytearray(p.read_bytes())
        data[len(data) // 3] ^= 0x5A
        p.write_bytes(bytes(data))
    a, d = session(cache, body)
    assert np.all(a == 13.0) and d == {"knl_hits": 0, "knl_misses": 1, "cubin_hits": 0, "cubin_misses": 1}

    def failing():
        err, _ = capi.jit(SQ, capi.clauses(("transform", TR, "raises")), vec)
        return err

    n_files = len(list((cache / "jit").iterdir()))
    for _ in range(2):
        err, d = session(cache, failing)
        assert err > 0 and d == {"knl_hits": 0, "knl_misses": 0, "cubin_hits": 0, "cubin_misses": 0}
    assert len(list((cache / "jit").iterdir())) == n_files


def test_cache_can_be_turned_off(cache, monkeypatch):
    monkeypatch.setenv("NOMP_JIT_CACHE", "0")
    _, d = session(cache, workload)
    assert d == {"knl_hits": 0, "knl_misses": 0, "cubin_hits": 0, "cubin_misses": 0}
    assert not (cache / "jit").exists()


def test_cache_directory_must_be_private(cache):
    """Entries are loaded as device code: a directory that others can write to (or that belongs to somebody else) is not
    used at all, and a directory the runtime creates itself is created for this user only."""
    _, d = session(cache, workload)
    assert d["knl_misses"] == 4 and (os.stat(cache / "jit").st_mode & 0o077) == 0
    os.chmod(cache / "jit", 0o777)
    _, d = session(cache, workload)
    assert d == {"knl_hits": 0, "knl_misses": 0, "cubin_hits": 0, "cubin_misses": 0}, d     # cache off, kernels still built
    os.chmod(cache / "jit", 0o700)
    _, d = session(cache, workload)
    assert d == {"knl_hits": 4, "knl_misses": 0, "cubin_hits": 2, "cubin_misses": 0}
