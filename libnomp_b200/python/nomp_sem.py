"""Annotation script for spectral-element kernels (`--nomp-annotations-script nomp_sem`).

Same keys as the reference's tests/sem.py:10-36 -- `grid_loop`, `element_loop`, `dof_loop` -- with the difference that
`dof_loop` takes effect (the reference computes the split and drops the result, tests/sem.py:20-24):

    {"annotate", "element_loop", "e"}    one thread block per element                          (g.0)
    {"annotate", "dof_loop", "i"}        the loops over the points of an element become the    (l.0, then l.1, l.2
    {"annotate", "dof_loop", "j"}        thread axes of the block, in the order of the          for the next ones)
    {"annotate", "dof_loop", "k"}        clauses; every loop of that name in the kernel
    {"annotate", "grid_loop", "i"}       a flat loop: split by the block size, g.0 / l.0

Per-element temporaries declared inside the element loop and written inside the dof loops (`double ur[n][n][n];`)
then live in shared memory, with a barrier after each dof loop nest (nomp_bridge.emit_cuda): the usual one-block-per-
element schedule of SEM operators, for kernels the hand-written families do not cover.  Extents must be known at jit
time (pass n as NOMP_INT | NOMP_JIT) and their product must fit a thread block.
"""
import loopy as lp


def annotate(knl, annotations, context):
    inames = knl.default_entrypoint.all_inames()
    block = min(512, int(context.get("device::max_threads_per_block", 1024)))
    for key, loop in annotations.items():
        if loop not in inames:
            continue
        if key == "element_loop":
            knl = lp.tag_inames(knl, [(loop, "g.0")])
        elif key == "dof_loop":
            used = {t for t in knl.tags().values() if t and t.startswith("l.")}
            axis = len(used)
            if axis > 2:
                raise ValueError(f"dof_loop '{loop}': a thread block has three axes and all are taken ({sorted(used)})")
            knl = lp.tag_inames(knl, [(loop, f"l.{axis}")])
        elif key == "grid_loop":
            knl = lp.split_iname(knl, loop, block)
            knl = lp.tag_inames(knl, [(f"{loop}_outer", "g.0"), (f"{loop}_inner", "l.0")])
    return knl
