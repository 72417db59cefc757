"""The RUNTIME of libnomp_b200 on the CPU tier: the reference's own nomp-api test programs (compiled unmodified against
our libnomp.so, oracle/_ref/tests) run against a TEST DOUBLE of the CUDA runtime (tests/hostdev: device memory = host
memory, "NVRTC" = g++ + the cooperative emulator, native-family calls answered by the oracle), LD_PRELOADed into the
test programs only.  This is the part pocl plays in the reference's CI (reference .github/workflows/ci.yml:60-77).

What it covers without a GPU: nomp_init / finalize cycles and argument parsing, the mapping table (sub-ranges, re-typing,
error paths), the jit bridge inside the embedded interpreter (user transform scripts, clauses, error strings from the
right files), argument marshalling by name, launch-size expressions, the generic emitter's kernels for every language
feature the reference tests use, the reduce skeleton with its publication protocol, the reduce finish, the on-disk JIT
cache.  The product is unchanged and has no CPU path: without the preload the same program fails in nomp_init().
"""
import concurrent.futures as cf
import os
import shutil
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
REF_TESTS = ROOT / "oracle" / "_ref" / "tests"
HOSTDEV = ROOT / "tests" / "hostdev"
CUDA_HOME = Path(os.environ.get("CUDA_HOME", "/usr/local/cuda"))

PROGRAMS = ["nomp-api-000", "nomp-api-020", "nomp-api-021", "nomp-api-050", "nomp-api-100", "nomp-api-150", "nomp-api-200",
            "nomp-api-205", "nomp-api-220", "nomp-api-225", "nomp-api-240", "nomp-api-300", "nomp-api-350", "nomp-api-400",
            "nomp-api-500", "nomp-api-600"]


@pytest.fixture(scope="module")
def double(tmp_path_factory):
    if not (REF_TESTS / "nomp-api-000").exists():
        pytest.skip("reference test programs were not built (needs /root/reference at build time)")
    if not (CUDA_HOME / "include" / "cuda_runtime.h").exists():
        pytest.skip("CUDA headers not found")
    from libnomp_b200 import build as b
    from oracle import ffi
    b.build_kernels(), b.build_libnomp()
    ffi.lib()
    out = HOSTDEV / "_build"
    out.mkdir(exist_ok=True)
    so = out / "libhostdev.so"
    srcs = [HOSTDEV / "fake_cuda.c", HOSTDEV / "fake_nompk.c"]
    if not so.exists() or any(s.stat().st_mtime > so.stat().st_mtime for s in srcs):
        subprocess.run(["gcc", "-O1", "-g", "-Wall", "-fPIC", "-shared", "-fvisibility=hidden", "-I", str(CUDA_HOME / "include"),
                        "-I", str(ROOT / "include"), "-o", str(so), *map(str, srcs), "-ldl", "-lrt"], check=True)
    # multi-rank pieces: an NCCL double (found by the dlopen of src/comm.c through LD_LIBRARY_PATH) and the product's
    # rank-exchange device code compiled for the host (tests/hostdev/build_devicecode.py)
    nccl = out / "libnccl.so.2"
    if not nccl.exists() or (HOSTDEV / "fake_nccl.c").stat().st_mtime > nccl.stat().st_mtime:
        subprocess.run(["gcc", "-O1", "-g", "-Wall", "-fPIC", "-shared", "-fvisibility=hidden", "-I", str(CUDA_HOME / "include"),
                        "-o", str(nccl), str(HOSTDEV / "fake_nccl.c")], check=True)
    dev = out / "libhostdev_devicecode.so"
    dev_deps = [HOSTDEV / "build_devicecode.py", ROOT / "tests" / "cuda_emulation.py",
                ROOT / "libnomp_b200" / "csrc" / "kernels" / "nompk_gridreduce.cuh"]
    if not dev.exists() or any(d.stat().st_mtime > dev.stat().st_mtime for d in dev_deps):
        subprocess.run([sys.executable, str(HOSTDEV / "build_devicecode.py"), str(dev)], check=True)
    for f in REF_TESTS.glob("*.pyc.bin"):       # the reference's transform scripts travel byte-compiled (oracle/Makefile)
        shutil.copyfile(f, f.with_suffix(""))
    work = tmp_path_factory.mktemp("hostdev")
    env = dict(os.environ, NOMP_INSTALL_DIR=str(ROOT / "libnomp_b200"), NOMP_JIT_CACHE="0", NOMP_HOSTDEV_DIR=str(work),
               NOMP_HOSTDEV_PYTHON=sys.executable, NOMP_HOSTDEV_COMPILER=str(HOSTDEV / "compile_kernel.py"),
               NOMP_HOSTDEV_ORACLE=str(ROOT / "oracle" / "libnomp_oracle.so"), NOMP_HOSTDEV_DEVICECODE=str(dev))
    for k in ("NOMP_BACKEND", "NOMP_DEVICE", "NOMP_PLATFORM", "NOMP_VERBOSE", "NOMP_COMM_SIZE", "NOMP_COMM_RANK"):
        env.pop(k, None)
    return so, env


def run_program(name, env, preload=None, timeout=900):
    e = dict(env)
    if preload is not None:
        e["LD_PRELOAD"] = str(preload)
    args = [str(REF_TESTS / name), "--nomp-backend", "cuda", "--nomp-device", "0", "--nomp-platform", "0", "--nomp-install-dir",
            e["NOMP_INSTALL_DIR"], "--nomp-verbose", "1", "--nomp-annotations-script", "sem"]      # reference scripts/lnrun:120-130
    return subprocess.run(args, cwd=REF_TESTS, env=e, capture_output=True, text=True, timeout=timeout)


def sanitized_runtime(so, env):
    """libnomp.so rebuilt with -fsanitize=address,undefined (tests/hostdev/_build/asan; the test programs find it through
    LD_LIBRARY_PATH) -> (environment, LD_PRELOAD string), or None where gcc has no sanitizer runtimes.  The reference
    keeps an ASAN option for its library and tests (reference CMakeLists.txt:98-110).  Leak checking is off: the embedded
    CPython keeps its arenas."""
    import sysconfig
    from libnomp_b200 import build as b
    asan, ubsan = (subprocess.run(["gcc", f"-print-file-name={n}"], capture_output=True, text=True).stdout.strip()
                   for n in ("libasan.so", "libubsan.so"))
    if not (Path(asan).is_absolute() and Path(ubsan).is_absolute()):
        return None
    out = HOSTDEV / "_build" / "asan"
    out.mkdir(parents=True, exist_ok=True)
    lib = out / "libnomp.so"
    src_dir = ROOT / "libnomp_b200" / "csrc" / "libnomp"
    srcs = [src_dir / s for s in b.LIBNOMP_SRCS]
    deps = srcs + list((src_dir / "include").glob("*.h")) + list((ROOT / "include").glob("*.h"))
    if not lib.exists() or any(d.stat().st_mtime > lib.stat().st_mtime for d in deps):
        pyver = f"python{sys.version_info.major}.{sys.version_info.minor}"
        subprocess.run(["gcc", "-O1", "-g", "-std=gnu11", "-fPIC", "-shared", "-fsanitize=address,undefined", "-fno-omit-frame-pointer",
                        "-fvisibility=hidden", "-I", str(ROOT / "include"), "-I", str(src_dir / "include"),
                        "-I", sysconfig.get_paths()["include"], "-I", str(CUDA_HOME / "include"),
                        f'-DNOMP_DEFAULT_INSTALL_DIR="{ROOT / "libnomp_b200"}"', "-o", str(lib), *map(str, srcs),
                        "-L", str(b.LIB), "-lnompk", "-L", str(CUDA_HOME / "lib64"), "-lcudart", "-lnvrtc",
                        "-L", sysconfig.get_config_var("LIBDIR") or "/usr/lib/x86_64-linux-gnu", f"-l{pyver}", "-ldl", "-lm",
                        "-lpthread", f"-Wl,-rpath,{b.LIB}", f"-Wl,-rpath,{CUDA_HOME}/lib64"], check=True)
    env = dict(env, LD_LIBRARY_PATH=str(out), ASAN_OPTIONS="detect_leaks=0", UBSAN_OPTIONS="print_stacktrace=1")
    r = subprocess.run(["ldd", str(REF_TESTS / "nomp-api-000")], env=env, capture_output=True, text=True)
    assert str(lib) in r.stdout, "the test programs do not pick up the sanitized library"
    return env, f"{asan} {ubsan} {so}"


def test_reference_suite_passes_on_the_cuda_test_double(double):
    """All 16 programs of the reference's suite, unmodified, through our runtime: every one exits 0 and prints no
    failed case -- ten of them with the runtime built under the address and undefined-behaviour sanitizers, neither of
    which may report anything."""
    so, env = double
    sanitized = sanitized_runtime(so, env)
    # the programs that launch hundreds of emulated grids stay on the plain library: the emulator allocates a stack per
    # thread and launch, which ASAN's allocator makes several times slower; their runtime calls are the same as the others'
    light = {"nomp-api-000", "nomp-api-020", "nomp-api-021", "nomp-api-050", "nomp-api-100", "nomp-api-150", "nomp-api-225",
             "nomp-api-350", "nomp-api-500", "nomp-api-600"}

    def one(prog):
        if sanitized and prog in light:
            return run_program(prog, sanitized[0], sanitized[1])
        return run_program(prog, env, so)

    with cf.ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 2)) as ex:
        results = dict(zip(PROGRAMS, ex.map(one, PROGRAMS)))
    bad = {p: (r.returncode, (r.stdout + r.stderr)[-1500:]) for p, r in results.items() if r.returncode != 0 or "Failed" in r.stdout}
    assert not bad, bad
    for p, r in results.items():
        text = r.stdout + r.stderr
        assert "AddressSanitizer" not in text and "runtime error" not in text, (p, text[-3000:])
    assert sum(r.stdout.count("Passed") for r in results.values()) >= 66       # test functions (each covers six types)


def test_without_the_double_there_is_no_cpu_path(double):
    """The same program without the preload, on a machine without a GPU: nomp_init() reports a CUDA failure."""
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("this machine has a GPU")
    except ImportError:
        pass
    _, env = double
    r = run_program("nomp-api-000", env)
    assert r.returncode != 0
    assert "CUDA runtime failure" in r.stderr + r.stdout


def test_jit_cache_serves_a_second_process(double, tmp_path):
    """On-disk JIT cache end to end (src/jitcache.c): the second process finds every program of the first one on disk --
    bridge output and "CUBIN" -- and gives the same results."""
    so, env = double
    env = dict(env, NOMP_JIT_CACHE="1", NOMP_JIT_CACHE_DIR=str(tmp_path / "cache"))
    first = run_program("nomp-api-225", env, so)
    assert first.returncode == 0, first.stdout + first.stderr
    entries = sorted(p.name for p in (tmp_path / "cache").iterdir())
    assert any(n.endswith(".knl") for n in entries) and any(n.endswith(".cubin") for n in entries)
    kernels_before = len(list(Path(env["NOMP_HOSTDEV_DIR"]).glob("*.cu")))
    second = run_program("nomp-api-225", env, so)
    assert second.returncode == 0 and "Failed" not in second.stdout, second.stdout + second.stderr
    assert len(list(Path(env["NOMP_HOSTDEV_DIR"]).glob("*.cu"))) == kernels_before      # nothing was compiled again
    assert sorted(p.name for p in (tmp_path / "cache").iterdir()) == entries


SYMMETRY_SCRIPT = r"""
import ctypes as C, sys
import numpy as np
sys.path.insert(0, "%(root)s")
sys.path.insert(0, "%(root)s/libnomp_b200/python")
from libnomp_b200 import capi
from nomp_bridge.families import AX_KERNEL_SOURCE, AX_DOT_KERNEL_SOURCE
from oracle import ffi
fake = C.CDLL(None)
fake.nomp_hostdev_last_ax_flags.restype = C.c_uint
P, I, F, JIT = capi.NOMP_PTR, capi.NOMP_INT, capi.NOMP_FLOAT, capi.NOMP_JIT
capi.check(capi.init())
n, E = 8, 3
u, g = np.ones(E * n ** 3), np.ones(E * 6 * n ** 3)
D = np.ascontiguousarray(ffi.gll_derivative(n)[0].ravel())
w = np.zeros_like(u)
args = [("w", 8, P), ("u", 8, P), ("g", 8, P), ("D", 8, P), ("E", 4, I), ("n", 4, I | JIT, C.c_int(n))]
err, kid = capi.jit(AX_KERNEL_SOURCE, capi.clauses(), args)
capi.check(err)
err, kdot = capi.jit(AX_DOT_KERNEL_SOURCE, capi.clauses(("reduce", "pap", "+")), args + [("pap", 8, F)])
capi.check(err)
for a in (u, g, D, w):
    capi.check(capi.update(a.ctypes.data, 0, a.size, 8, capi.NOMP_TO))
seen = []
def run(k=kid):
    extra = (C.c_double(0.0),) if k == kdot else ()
    capi.check(capi.run(k, w.ctypes.data, u.ctypes.data, g.ctypes.data, D.ctypes.data, C.c_int(E), *extra))
    seen.append(fake.nomp_hostdev_last_ax_flags())
run(); run(); run(kdot)                       # GLL matrix: antisymmetric (2), staged once (1 = cached from the second run on)
D[1] += 0.25
capi.check(capi.update(D.ctypes.data, 0, D.size, 8, capi.NOMP_TO))
run(); run()                                  # any other matrix: the general kernels
D[1] -= 0.25
capi.check(capi.update(D.ctypes.data, 0, D.size, 8, capi.NOMP_TO))
run(kdot)                                     # the decision follows the contents
print("FLAGS", seen)
capi.nomp().nomp_finalize_excluding_interpreter()
"""


def test_backend_finds_out_whether_D_is_antisymmetric(double, monkeypatch):
    """backends/cuda.c: ax_D_antisymmetric reads the device image of D once per version and hands
    NOMPK_AX_D_ANTISYMMETRIC (2) to the kernel library for a GLL matrix only; NOMPK_AX_D_CACHED (1) from the second
    launch with the same D.  The test double records the flags of the native calls."""
    so, env = double
    env = dict(env, LD_PRELOAD=str(so), NOMP_HOSTDEV_ACTIVE="1")
    r = subprocess.run([sys.executable, "-c", SYMMETRY_SCRIPT % {"root": str(ROOT)}], cwd=ROOT, env=env, capture_output=True,
                       text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("FLAGS")][-1]
    assert line == "FLAGS [2, 3, 3, 0, 1, 2]", line
    r = subprocess.run([sys.executable, "-c", SYMMETRY_SCRIPT % {"root": str(ROOT)}], cwd=ROOT,
                       env=dict(env, NOMP_B200_NO_EVEN_ODD="1"), capture_output=True, text=True, timeout=600)
    assert [ln for ln in r.stdout.splitlines() if ln.startswith("FLAGS")][-1] == "FLAGS [0, 1, 1, 0, 1, 0]", r.stdout[-500:]


# API-level tests of the GPU tier whose sizes the emulator finishes in seconds; the rest need the real device
# (2^25-element reductions, 256 MiB transfers, throughput floors, direct libnompk calls on torch tensors, NCCL / IPC).
API_TESTS = ["tests/test_nomp_api_gpu.py", "tests/test_jit_cache_gpu.py", "tests/test_sem_annotations_gpu.py",
             "tests/test_system_gpu.py::test_cg_example_matches_host_cg",      # examples/cg_poisson.c against a host CG
             "tests/test_system_gpu.py::test_smoke_entry_point",                # the call sequence of __graft_entry__.smoke()
             "tests/test_device_scalars_gpu.py"]                                # reduce results and scalars that stay on the device
TOO_BIG = ["tests/test_nomp_api_gpu.py::test_reduce_large_sizes", "tests/test_nomp_api_gpu.py::test_reduce_clause_at_baseline_size_two_level_finish",
           "tests/test_nomp_api_gpu.py::test_repeated_updates_pin_the_host_range",
           "tests/test_sem_annotations_gpu.py::test_annotated_operator_throughput",
           "tests/test_nomp_api_gpu.py::test_ax_takes_the_even_odd_path_for_an_antisymmetric_D"]   # compares with the real kernels


def test_api_level_gpu_tests_run_on_the_cuda_test_double(double):
    """The ctypes tests of the public API (`-m gpu` tier) in a child pytest whose libnomp.so sees the test double: update
    semantics and errors, sub-range mappings, jit and run error paths, every kernel-language feature against the kernel
    string compiled by gcc, reduce clauses of six types with +, *, min, max, the Ax and fused CG kernel strings (native
    calls answered by the oracle: marshalling only), asynchronous updates, the JIT cache, SEM annotations, and the CG example program against a host CG."""
    so, env = double
    env = dict(env, LD_PRELOAD=str(so), NOMP_HOSTDEV_ACTIVE="1")
    cmd = [sys.executable, "-m", "pytest", *API_TESTS, "-m", "gpu", "-q", "-p", "no:cacheprovider", "--timeout", "600",
           "-n", str(min(6, os.cpu_count() or 1))]
    for t in TOO_BIG:
        cmd += ["--deselect", t]
    r = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=3000)
    tail = "\n".join(r.stdout.splitlines()[-40:])
    assert r.returncode == 0, tail + r.stderr[-2000:]
    import re
    m = re.search(r"(\d+) passed", r.stdout)
    assert m and int(m.group(1)) >= 84 and "failed" not in r.stdout and "skipped" not in r.stdout, tail


# ---- several ranks: one process per rank, shared memory standing in for NVLink peer memory -------------------------------

_RUN_IDS = __import__("itertools").count()      # (next() on a count is atomic in CPython: the cases below run in threads)


def run_ranks(world, args, env, so, extra=None, timeout=600):
    """`world` processes of examples/cg_poisson.c on the test double: NCCL double through LD_LIBRARY_PATH, "device" memory
    in POSIX shared memory so that CUDA IPC handles work between the processes.  Returns the JSON lines of every rank."""
    import json
    exe = ROOT / "libnomp_b200" / "build" / "cg_poisson"
    if not exe.exists():
        pytest.skip("examples/cg_poisson was not built")
    work = Path(env["NOMP_HOSTDEV_DIR"])
    idfile = work / f"id-{os.getpid()}-{next(_RUN_IDS)}"
    procs = []
    for r in range(world):
        e = dict(env, LD_PRELOAD=str(so), LD_LIBRARY_PATH=str(so.parent), NOMP_HOSTDEV_SHARED="1", NOMP_HOSTDEV_DEVICES=str(world),
                 NOMP_COMM_SIZE=str(world), NOMP_COMM_RANK=str(r), NOMP_COMM_ID_FILE=str(idfile), **(extra or {}))
        procs.append(subprocess.Popen([str(exe), *map(str, args), "--nomp-backend", "cuda", "--nomp-device", str(r), "--nomp-verbose", "1"],
                                      env=e, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=timeout)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode in (0, 2), o[-3000:]         # 2 = not converged within max_iter, not an error here
    return [[json.loads(line) for line in o.splitlines() if line.startswith("{")] for o in outs]


TWO_RANK_CASES = [("fused", "host"), ("fused", "device"), ("fused", "device3"), ("standalone", "host"), ("standalone", "device"),
                  ("nccl", "host"), ("nccl", "device3"), ("fused", "fused"), ("fused", "device_fused"), ("fused", "graph"),
                  ("standalone", "graph")]


def test_two_ranks_of_the_cg_example(double):
    """Eleven combinations of how a reduce clause is all-reduced and where the scalars live (see _two_rank_case), four at
    a time; every failure is reported with its case."""
    if not (ROOT / "libnomp_b200" / "build" / "cg_poisson").exists():
        pytest.skip("examples/cg_poisson was not built")
    with cf.ThreadPoolExecutor(max_workers=4) as ex:
        futures = {case: ex.submit(_two_rank_case, double, *case) for case in TWO_RANK_CASES}
    failed = {case: repr(f.exception())[:2000] for case, f in futures.items() if f.exception() is not None}
    assert not failed, failed
    assert not list(Path("/dev/shm").glob("nomp-hostdev-*")), "a rank left shared-memory objects behind"


def _two_rank_case(double, path, scalars):
    """src/comm.c end to end without GPUs -- file rendezvous of the NCCL id, CUDA-IPC exchange of the ranks' buffers, the
    agreement all-reduce -- and the three ways a reduce clause is all-reduced (fused into the reduction kernel,
    stand-alone kernel after it, ncclAllReduce), each with the scalars on the host and in device memory
    (nomp_b200_device_reductions).  The exchange itself is the product's finish_result on the emulator (native families)
    and the generated nomp_finish text (reduce skeleton), talking to each other across two processes.  Against a host CG
    on the 2E-element mesh; the ranks must see the same bits."""
    from tests.test_system_gpu import _cg_reference
    so, env = double
    extra = {"fused": {}, "standalone": {"NOMP_COMM_FUSED": "0"}, "nccl": {"NOMP_COMM_ALLREDUCE": "nccl"}}[path]
    E, n = 3, 8
    iters = 13 if scalars == "graph" else 12          # a replayed graph holds two iterations (after one launched normally)
    per_rank = run_ranks(2, [E, n, iters, "1e-30", scalars, 4], env, so, extra)
    ref = _cg_reference(2 * E, n, 5)
    ref_final = _cg_reference(2 * E, n, iters)[-1]["rr"]
    for lines in per_rank:
        assert abs(lines[0]["rr0"] - ref[0]["rr0"]) <= 1e-12 * ref[0]["rr0"]
        for it in range(5 if scalars not in ("device3", "device_fused", "graph") else 0):        # those print no per-iteration trace
            for key in ("pAp", "alpha", "rr"):
                assert abs(lines[1 + it][key] - ref[1 + it][key]) <= 1e-10 * abs(ref[1 + it][key]), (it, key)
        assert lines[-1]["iterations"] == iters and lines[-1]["scalars"] == scalars
        assert abs(lines[-1]["rr_final"] - ref_final) <= 1e-9 * ref_final
    assert per_rank[0][1:] and [{k: v for k, v in d.items() if k not in ("seconds", "ms_per_iter", "GDOF_per_s_per_rank")} for d in per_rank[0]] == \
        [{k: v for k, v in d.items() if k not in ("seconds", "ms_per_iter", "GDOF_per_s_per_rank")} for d in per_rank[1]]


def test_four_ranks_mixing_host_and_device_results(double):
    """Four ranks, fused path, scalars in device memory with the residual fetched every third iteration."""
    from tests.test_system_gpu import _cg_reference
    so, env = double
    E, n = 2, 8
    per_rank = run_ranks(4, [E, n, 9, "1e-30", "device", 3], env, so)
    ref = _cg_reference(4 * E, n, 5)
    for lines in per_rank:
        for it in range(5):
            for key in ("pAp", "alpha", "rr"):
                assert abs(lines[1 + it][key] - ref[1 + it][key]) <= 1e-10 * abs(ref[1 + it][key]), (it, key)
    assert len({json_line["rr_final"] for json_line in (lines[-1] for lines in per_rank)}) == 1


def test_graph_extension_from_c_under_the_sanitizers(double):
    """tests/hostdev/graph_smoke.c: capture / replay / refusals / free of nomp_b200_graph_* from C, with the runtime built
    under the address and undefined-behaviour sanitizers where gcc has them."""
    so, env = double
    exe = HOSTDEV / "_build" / "graph_smoke"
    src = HOSTDEV / "graph_smoke.c"
    from libnomp_b200 import build as b
    if not exe.exists() or src.stat().st_mtime > exe.stat().st_mtime:
        subprocess.run(["gcc", "-O1", "-g", "-Wall", "-I", str(ROOT / "include"), str(src), "-o", str(exe), "-L", str(b.LIB), "-lnomp",
                        "-lm", f"-Wl,-rpath,{b.LIB}"], check=True)
    sanitized = sanitized_runtime(so, env)
    run_env, preload = sanitized if sanitized else (env, str(so))
    r = subprocess.run([str(exe), "--nomp-backend", "cuda", "--nomp-verbose", "1"], env=dict(run_env, LD_PRELOAD=str(preload)),
                       capture_output=True, text=True, timeout=600)
    text = r.stdout + r.stderr
    assert r.returncode == 0 and "graph smoke: ok" in r.stdout, text[-3000:]
    assert "AddressSanitizer" not in text and "runtime error" not in text, text[-3000:]
