"""Build every native artefact of the repo in-tree (no JIT caches outside the repo).

    python -m libnomp_b200.build [--force] [--only kernels|libnomp|oracle|ref-tests]

Artefacts
  libnomp_b200/lib/libnompk.so   hand-written sm_100a kernels + C ABI (include/nompk.h); nvcc, sm_100a only
  libnomp_b200/lib/libnomp.so    the libnomp runtime (public API include/nomp.h) with the CUDA-only backend; gcc
  oracle/libnomp_oracle.so       CPU oracle (test infrastructure)
  oracle/_ref/tests/*            the reference's nomp-api test programs linked against our libnomp.so
                                 (only where /root/reference exists, i.e. in the build container)
The built files are git-ignored but travel to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import argparse
import concurrent.futures as cf
import os
import shutil
import subprocess
import sys
import sysconfig
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
PKG = ROOT / "libnomp_b200"
LIB = PKG / "lib"
OBJ = PKG / "build"
CUDA_HOME = Path(os.environ.get("CUDA_HOME", "/usr/local/cuda"))
NVCC = str(CUDA_HOME / "bin" / "nvcc")
REFERENCE = Path(os.environ.get("NOMP_REFERENCE_DIR", "/root/reference"))

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-I", str(ROOT / "include"),
]

# (source, object name, extra flags).  ax.cu is compiled without a macro (its C ABI) and, per supported n, in three parts
# (-DNOMPK_AX_N=<n> -DNOMPK_AX_PART=<0|1|2>: the production and fused kernels, and two halves of the shapes kept for
# profiling) -- in parallel: a single unit with every n and shape takes most of an hour, one unit per n ten minutes.
# Longest first.
KERNEL_UNITS = [("ax.cu", f"ax_n{n}" + (f"_p{part}" if part else "") + ".cu.o", [f"-DNOMPK_AX_N={n}", f"-DNOMPK_AX_PART={part}"])
                for n in (12, 10, 6, 8) for part in (2, 1, 0)] + \
               [("gs.cu", "gs.cu.o", []), ("reduce.cu", "reduce.cu.o", []), ("map.cu", "map.cu.o", []),
                ("nompk.cu", "nompk.cu.o", []), ("ax.cu", "ax.cu.o", [])]
LIBNOMP_SRCS = ["src/nomp.c", "src/log.c", "src/aux.c", "src/loopy.c", "src/reduction.c", "src/gridexpr.c", "src/comm.c", "src/jitcache.c", "src/gs.c",
                "backends/cuda.c"]


def _run(cmd, cwd=None):
    r = subprocess.run(cmd, cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(map(str, cmd)) + "\n" + r.stdout + "\n")
        raise RuntimeError(f"build step failed: {cmd[0]} ... {cmd[-1]}")
    return r.stdout


def _stale(target: Path, deps) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(Path(d).stat().st_mtime > t for d in deps if Path(d).exists())


def build_kernels(force=False):
    src_dir = PKG / "csrc" / "kernels"
    LIB.mkdir(exist_ok=True)
    OBJ.mkdir(exist_ok=True)
    headers = [*src_dir.glob("*.cuh"), ROOT / "include" / "nompk.h"]
    out = LIB / "libnompk.so"
    jobs = []
    for s, oname, extra in KERNEL_UNITS:
        o = OBJ / oname
        if force or _stale(o, [src_dir / s, *headers]):
            jobs.append([NVCC, *NVCC_FLAGS, *extra, "-c", str(src_dir / s), "-o", str(o)])
    if jobs:
        with cf.ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
            list(ex.map(_run, jobs))   # KERNEL_UNITS is ordered longest first
    objs = [str(OBJ / oname) for _, oname, _ in KERNEL_UNITS]
    if force or jobs or _stale(out, objs):
        _run([NVCC, "-shared", "-cudart", "shared", "-gencode", "arch=compute_100a,code=sm_100a",
              "-Xlinker", f"-rpath,{CUDA_HOME}/lib64", *objs, "-o", str(out)])
    return out


def build_libnomp(force=False):
    src_dir = PKG / "csrc" / "libnomp"
    LIB.mkdir(exist_ok=True)
    OBJ.mkdir(exist_ok=True)
    out = LIB / "libnomp.so"
    srcs = [src_dir / s for s in LIBNOMP_SRCS]
    if not all(s.exists() for s in srcs):
        missing = [str(s) for s in srcs if not s.exists()]
        raise RuntimeError(f"libnomp sources missing: {missing}")
    headers = list((src_dir / "include").glob("*.h")) + list((ROOT / "include").glob("*.h"))
    pyinc = sysconfig.get_paths()["include"]
    pyver = f"python{sys.version_info.major}.{sys.version_info.minor}"
    pylibdir = sysconfig.get_config_var("LIBDIR") or "/usr/lib/x86_64-linux-gnu"
    cflags = ["-O2", "-g", "-std=gnu11", "-fPIC", "-Wall", "-Wno-unused-function", "-fvisibility=hidden",
              "-I", str(ROOT / "include"), "-I", str(src_dir / "include"), "-I", pyinc,
              "-I", str(CUDA_HOME / "include"),
              f'-DNOMP_DEFAULT_INSTALL_DIR="{PKG}"']
    jobs, objs = [], []
    for s in srcs:
        o = OBJ / ("libnomp_" + s.name + ".o")
        objs.append(str(o))
        if force or _stale(o, [s, *headers]):
            # the path handed to the compiler ends in libnomp/src/<file>.c: error strings embed __FILE__ and the
            # reference tests match "libnomp/src/nomp.c" and "src/loopy.c" (reference tests/nomp-api-000.c:20)
            jobs.append(["gcc", *cflags, "-c", str(s), "-o", str(o)])
    if jobs:
        with cf.ThreadPoolExecutor(max_workers=len(jobs)) as ex:
            list(ex.map(_run, jobs))
    if force or jobs or _stale(out, objs):
        _run(["gcc", "-shared", "-o", str(out), *objs,
              "-L", str(LIB), "-lnompk", "-L", str(CUDA_HOME / "lib64"), "-lcudart", "-lnvrtc",
              "-L", pylibdir, f"-l{pyver}", "-ldl", "-lm", "-lpthread",
              "-Wl,-rpath,$ORIGIN", f"-Wl,-rpath,{CUDA_HOME}/lib64"])
    return out


def build_tools(force=False):
    """C drivers under tools/ (launch-overhead sweep of BASELINE.json configs[4])."""
    OBJ.mkdir(exist_ok=True)
    out = OBJ / "launch_overhead"
    src = ROOT / "tools" / "launch_overhead.c"
    if force or _stale(out, [src, LIB / "libnomp.so"]):
        _run(["gcc", "-O2", "-I", str(ROOT / "include"), str(src), "-o", str(out), "-L", str(LIB), "-lnomp",
              f"-Wl,-rpath,{LIB}", "-Wl,-rpath,$ORIGIN/../lib"])
    common = ROOT / "examples" / "sem_common.h"
    for name in ("cg_poisson", "poisson_box"):
        ex = OBJ / name
        ex_src = ROOT / "examples" / f"{name}.c"
        if force or _stale(ex, [ex_src, common, LIB / "libnomp.so"]):
            _run(["gcc", "-O2", "-Wall", "-I", str(ROOT / "include"), str(ex_src), "-o", str(ex), "-L", str(LIB), "-lnomp", "-lm",
                  f"-Wl,-rpath,{LIB}", "-Wl,-rpath,$ORIGIN/../lib"])
    return out


def build_oracle(force=False):
    out = ROOT / "oracle" / "libnomp_oracle.so"
    if force or _stale(out, [ROOT / "oracle" / "nomp_oracle.c"]):
        _run(["make", "-C", str(ROOT / "oracle"), "-B" if force else "-s", "oracle"])
    return out


def build_ref_tests(force=False):
    """Compile the reference's own test programs against our libnomp.so (build container only)."""
    if not (REFERENCE / "tests").is_dir():
        return None
    args = ["make", "-C", str(ROOT / "oracle"), f"REFERENCE={REFERENCE}", "ref-tests"]
    if force:
        args.insert(1, "-B")
    _run(args)
    return ROOT / "oracle" / "_ref" / "tests"


def build_all(force=False, only=None):
    built = {}
    if only in (None, "kernels"):
        built["kernels"] = build_kernels(force)
    if only in (None, "libnomp"):
        built["libnomp"] = build_libnomp(force)
    if only in (None, "tools"):
        built["tools"] = build_tools(force)
    if only in (None, "oracle"):
        built["oracle"] = build_oracle(force)
    if only in (None, "ref-tests"):
        built["ref-tests"] = build_ref_tests(force)
    return built


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--only", choices=["kernels", "libnomp", "tools", "oracle", "ref-tests"])
    a = ap.parse_args()
    for k, v in build_all(a.force, a.only).items():
        print(f"{k}: {v}")


if __name__ == "__main__":
    main()
