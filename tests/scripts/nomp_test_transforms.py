"""Transform scripts used by tests/test_nomp_api_gpu.py (same style as a nompcc user's transforms.py: the functions
receive the kernel and the context dict and use the loopy-compatible API)."""
import math

import loopy as lp


def tile(knl, context):
    (iname,) = knl.default_entrypoint.all_inames()
    block = min(512, context["device::max_threads_per_block"])
    knl = lp.split_iname(knl, iname, block, inner_iname=f"{iname}_inner", outer_iname=f"{iname}_outer")
    return lp.tag_inames(knl, {f"{iname}_outer": "g.0", f"{iname}_inner": "l.0"})


def tile_small(knl, context):
    (iname,) = knl.default_entrypoint.all_inames()
    knl = lp.split_iname(knl, iname, 32)
    return lp.tag_inames(knl, {f"{iname}_outer": "g.0", f"{iname}_inner": "l.0"})


def tile_outer(knl, context):
    (i, j) = sorted(knl.default_entrypoint.all_inames())
    block = min(512, context["device::max_threads_per_block"])
    knl = lp.split_iname(knl, i, block, inner_iname=f"{i}_inner", outer_iname=f"{i}_outer")
    return lp.tag_inames(knl, {f"{i}_outer": "g.0", f"{i}_inner": "l.0", j: "for"})


def tile_2d(knl, context):
    block = int(math.sqrt(min(1024, context["device::max_threads_per_block"])))
    knl = lp.split_iname(knl, "i", block)
    knl = lp.split_iname(knl, "j", block)
    tags = {"i_outer": "g.0", "i_inner": "l.0", "j_outer": "g.1", "j_inner": "l.1"}
    if "k" in knl.default_entrypoint.all_inames():
        tags["k"] = "for"
    return lp.tag_inames(knl, tags)


def element_dof(knl, context):
    return lp.tag_inames(knl, {"i": "g.0", "j*": "l.0"})


def raises(knl, context):
    return undefined_name  # noqa: F821  (deliberate NameError)


def returns_garbage(knl, context):
    return 42


def checks_context(knl, context):
    assert context["backend::name"] == "cuda" and context["device::vendor"] == "NVIDIA"
    assert isinstance(context["device::max_threads_per_block"], int) and context["device::arch"].startswith("sm_")
    assert isinstance(context["device::driver"], int) and context["device::name"]
    return tile(knl, context)
