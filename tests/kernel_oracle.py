"""Test helper: reference semantics of ANY nomp kernel string = the C function itself, compiled by gcc and run
serially on the host (every nomp kernel is a complete C function, e.g. reference
tests/nomp-api-200-impl.h:36-40).  Test infrastructure, like everything under oracle/."""
from __future__ import annotations

import ctypes as C
import hashlib
import re
import subprocess
import tempfile
from pathlib import Path

_CACHE = {}
_DIR = Path(tempfile.mkdtemp(prefix="nomp-kernel-oracle-"))


def compile_kernel(src: str, extra_defs: str = "") -> C.CFUNCTYPE:
    """gcc -O2 -ffp-contract=off the kernel string; returns the ctypes function (restype None)."""
    key = hashlib.sha256((src + extra_defs).encode()).hexdigest()[:16]
    if key in _CACHE:
        return _CACHE[key]
    m = re.search(r"void\s+([A-Za-z_]\w*)\s*\(", src)
    assert m, "kernel string has no function"
    name = m.group(1)
    c_file = _DIR / f"k{key}.c"
    so_file = _DIR / f"k{key}.so"
    c_file.write_text(extra_defs + "\n" + src + "\n")
    subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-w", "-o", str(so_file), str(c_file)], check=True)
    lib = C.CDLL(str(so_file))
    fn = getattr(lib, name)
    fn.restype = None
    _CACHE[key] = fn
    return fn


def run_kernel(src: str, *args, defs: str = ""):
    """Call the compiled kernel; numpy arrays are passed as pointers, Python ints as C int, floats as C double."""
    import numpy as np
    fn = compile_kernel(src, defs)
    cargs = []
    for a in args:
        if isinstance(a, np.ndarray):
            cargs.append(C.c_void_p(a.ctypes.data))
        elif isinstance(a, (int, np.integer)):
            cargs.append(C.c_int(int(a)))
        elif isinstance(a, float):
            cargs.append(C.c_double(a))
        else:
            cargs.append(a)
    fn(*cargs)
