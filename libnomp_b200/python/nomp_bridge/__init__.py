"""Embedded-Python side of libnomp's jit bridge (imported by libnomp.so through its own interpreter, the way the
reference imports python/loopy_api.py and python/reduction.py -- reference src/loopy.c:130-182, :247-284).

Pure Python, standard library only (see cparse.py for why)."""
from .api import (annotate_passthrough, c_to_loopy, fix_parameters, get_grid_size, get_knl_name, get_knl_src,  # noqa: F401
                  realize_reduction)
from .ir import Kernel, KernelError  # noqa: F401
