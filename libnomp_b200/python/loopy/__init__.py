"""`loopy`-compatible surface for nomp transform / annotation scripts.

User scripts written for the reference do `import loopy as lp` and call lp.split_iname / lp.tag_inames on the
kernel they are handed (reference tests/nomp_api_100.py:6, nomp_api_225.py:4, nomp_api_300.py:5,
nomp_api_400.py:3, tests/sem.py:5, docs/examples.rst:48-60).  The real loopy (and pymbolic, islpy, pytools)
is not a dependency of this implementation; this package maps those calls onto nomp_bridge's own loop-nest IR.
Only the schedule-level API is provided -- anything else raises AttributeError, which libnomp reports as
NOMP_PY_CALL_FAILURE just as it would any other failing user script.
"""
from nomp_bridge.ir import Kernel, KernelError, fix_parameters, split_iname, tag_inames  # noqa: F401

from . import translation_unit  # noqa: F401

TranslationUnit = Kernel
LoopKernel = Kernel
LoopyError = KernelError
VERSION = (2024, 1)
__version__ = "nomp-b200-shim"


class AddressSpace:
    PRIVATE, LOCAL, GLOBAL = 0, 1, 2


class CudaTarget:
    pass


def set_options(knl, *args, **kwargs):
    return knl


def add_inames_for_unused_hw_axes(knl, *args, **kwargs):
    return knl


def prioritize_loops(knl, *args, **kwargs):
    return knl
