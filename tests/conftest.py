import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


# Tests exercise the transform bridge and NVRTC for real: the on-disk JIT cache is off unless a test turns it on
# with its own directory (tests/test_jit_cache_gpu.py).
os.environ.setdefault("NOMP_JIT_CACHE", "0")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_sessionstart(session):
    """A fresh clone has no built libraries (they are git-ignored): build them once, in-tree, before the first test
    (what `python -m libnomp_b200.build` / __graft_entry__.build() do; a couple of minutes, nvcc cross-compiles without a
    GPU).  Nothing is built when they are already there; a failed build leaves the tests to fail loudly."""
    lib = ROOT / "libnomp_b200" / "lib"
    if (lib / "libnompk.so").exists() and (lib / "libnomp.so").exists():
        return
    try:
        from libnomp_b200 import build
        print("\n[tests] building libnompk.so / libnomp.so (first run in this tree) ...", flush=True)
        build.build_all()
    except Exception as exc:  # pragma: no cover
        print(f"[tests] build failed: {exc}", flush=True)


def _has_gpu():
    # tests/test_hostdev_cpu.py re-runs API tests in a child pytest whose "device" is the CUDA test double (tests/hostdev)
    if os.environ.get("NOMP_HOSTDEV_ACTIVE") == "1":
        return True
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def repo_root():
    return ROOT
