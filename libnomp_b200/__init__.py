"""libnomp_b200: a B200-native execution path behind libnomp's public C API.

The product is two shared libraries built in-tree by ``libnomp_b200.build``:

* ``lib/libnomp.so``  -- libnomp's public C API (``include/nomp.h``: nomp_init / nomp_update / nomp_jit /
  nomp_run / nomp_sync / nomp_finalize) on a CUDA-only backend;
* ``lib/libnompk.so`` -- the hand-written sm_100a kernels (``include/nompk.h``) the backend dispatches to.

This Python package only holds the build script, ctypes bindings used by the tests and ``bench.py``
(``libnomp_b200.capi``) and the embedded-Python side of the jit bridge (``libnomp_b200/python``), which
libnomp.so imports through its own interpreter exactly like the reference imports ``loopy_api``.
There is no CPU fallback: the bindings raise if the native libraries are missing.
"""
from pathlib import Path

PACKAGE_DIR = Path(__file__).resolve().parent
REPO_ROOT = PACKAGE_DIR.parent
LIB_DIR = PACKAGE_DIR / "lib"
#: value for --nomp-install-dir / NOMP_INSTALL_DIR: libnomp.so appends "<install>/python" to sys.path
INSTALL_DIR = PACKAGE_DIR

__all__ = ["PACKAGE_DIR", "REPO_ROOT", "LIB_DIR", "INSTALL_DIR"]
