"""Annotation script used by tests/test_nomp_api_gpu.py (role of the reference's tests/sem.py)."""
import loopy as lp


def annotate(knl, annotations, context):
    inames = knl.default_entrypoint.all_inames()
    block = min(256, context["device::max_threads_per_block"])
    for key, loop in annotations.items():
        if loop not in inames:
            continue
        if key == "grid_loop":
            knl = lp.split_iname(knl, loop, block)
            knl = lp.tag_inames(knl, [(f"{loop}_outer", "g.0"), (f"{loop}_inner", "l.0")])
        elif key == "element_loop":
            knl = lp.tag_inames(knl, [(loop, "g.0")])
    return knl
