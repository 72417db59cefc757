"""World-size-2 checks of the multi-GPU host logic on CPU (gloo): the partitioning the ranks use, the combine step of
a partitioned reduction (what the NCCL allreduce of the per-rank scalar does), and the file rendezvous through which
rank 0 hands the ncclUniqueId to the other ranks."""
import os
import socket
import sys
import tempfile
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def partition(total, world, rank):
    """Contiguous block partition used by bench.py and DESIGN.md (e): rank r owns [r*total/world, (r+1)*total/world)."""
    return total * rank // world, total * (rank + 1) // world


def _worker(rank, world, port, tmp, results):
    import ctypes as C
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from libnomp_b200 import capi
    from oracle import ffi

    # 1. rendezvous: rank 0 publishes 128 bytes, everyone ends up with the same bytes
    lib = capi.nomp()
    blob = (C.c_ubyte * 128)()
    if rank == 0:
        rng = np.random.default_rng(7)
        for i, v in enumerate(rng.integers(0, 256, 128)):
            blob[i] = int(v)
    path = os.path.join(tmp, "id-file").encode()
    assert lib.nomp_b200_exchange_blob(path, rank, blob, 128) == 0
    mine = torch.tensor(list(blob), dtype=torch.int64)
    gathered = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(gathered, mine)
    assert all(torch.equal(gathered[0], g) for g in gathered)

    # 2. partitioned dot product: each rank reduces its slice with the oracle, the scalars are all-reduced
    n = 1 << 18
    x = ffi.fill_int_f64(n, 11, 0, 7)
    y = ffi.fill_int_f64(n, 12, 0, 7)
    lo, hi = partition(n, world, rank)
    part = torch.tensor([ffi.reduce_(0, ffi.F64, x[lo:hi].copy(), y[lo:hi].copy())], dtype=torch.float64)
    dist.all_reduce(part, op=dist.ReduceOp.SUM)
    assert part.item() == ffi.reduce_(0, ffi.F64, x, y)            # exact data: bitwise equal to the 1-rank result
    xi = ffi.fill_i64(n, 5)
    parti = torch.tensor([ffi.reduce_(0, ffi.I64, xi[lo:hi].copy())], dtype=torch.int64)
    dist.all_reduce(parti, op=dist.ReduceOp.SUM)
    assert parti.item() == ffi.reduce_(0, ffi.I64, xi)             # wrap-around adds are associative
    pm = torch.tensor([ffi.reduce_(3, ffi.F64, x[lo:hi].copy())], dtype=torch.float64)
    dist.all_reduce(pm, op=dist.ReduceOp.MAX)
    assert pm.item() == x.max()

    # 3. element-partitioned Ax: the slices stitched together are the global result (no halo: local operator)
    nn, E = 8, 12
    u = ffi.fill_int_f64(E * nn ** 3, 2, -4, 4)
    g = ffi.fill_int_f64(E * 6 * nn ** 3, 3, 0, 3)
    D = ffi.fill_int_f64(nn * nn, 4, -2, 2)
    elo, ehi = partition(E, world, rank)
    n3 = nn ** 3
    w_local = torch.from_numpy(ffi.ax(nn, u[elo * n3:ehi * n3].copy(), g[elo * 6 * n3:ehi * 6 * n3].copy(), D))
    pieces = [torch.zeros((ehi - elo) * n3, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(pieces, w_local)
    assert np.array_equal(torch.cat(pieces).numpy(), ffi.ax(nn, u, g, D))

    # 4. slab-partitioned gather-scatter, the algorithm of gs.cu on the host: fold the local copies of every id, exchange
    #    the partials of the ids both ranks hold (the interface plane, in ascending id order -- the list S(me, peer)),
    #    fold them in rank order, write back.  Must equal the oracle's rank-segmented fold of the whole mesh bit for bit.
    import bench
    n_pts, ex, ey, ez = 4, 3, 2, 2 * world
    nz = ez // world
    ids_full = ffi.box_ids(n_pts, ex, ey, ez)
    ids = bench.box_slab_ids(n_pts, ex, ey, ez, rank * nz, nz)
    per = ids_full.size // world
    assert np.array_equal(ids, ids_full[rank * per:(rank + 1) * per])
    v_full = ffi.fill_uniform_f64(ids_full.size, 31, 0.5, 1.5)
    local = ffi.gs(0, ffi.F64, ids, v_full[rank * per:(rank + 1) * per].copy())
    uniq = torch.from_numpy(np.unique(ids))
    sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([uniq.numel()]))
    padded = torch.zeros(int(max(t.item() for t in sizes)), dtype=torch.int64)
    padded[:uniq.numel()] = uniq
    everyone = [torch.zeros_like(padded) for _ in range(world)]
    dist.all_gather(everyone, padded)
    peer = 1 - rank
    shared = np.intersect1d(uniq.numpy(), everyone[peer][:int(sizes[peer].item())].numpy())   # ascending: S(me, peer)
    assert shared.size == (3 * ex + 1) * (3 * ey + 1)                                        # one plane of points
    first = {int(i): k for k, i in reversed(list(enumerate(ids)))}
    mine = torch.tensor([local[first[int(i)]] for i in shared], dtype=torch.float64)
    partials = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(partials, mine)
    total = partials[0] + partials[1]                                                        # rank order
    out = local.copy()
    where = {int(i): k for k, i in enumerate(shared)}
    for k, i in enumerate(ids):
        if int(i) in where:
            out[k] = total[where[int(i)]].item()
    want = ffi.gs(0, ffi.F64, ids_full, v_full.copy(), [r * per for r in range(world + 1)])
    assert np.array_equal(out, want[rank * per:(rank + 1) * per])
    results[rank] = 1
    dist.destroy_process_group()


def test_two_rank_host_logic():
    world = 2
    for total in (262144, 10, 7):
        spans = [partition(total, 3, r) for r in range(3)]
        assert spans[0][0] == 0 and spans[-1][1] == total and all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    assert partition(262144, 8, 3) == (3 * 32768, 4 * 32768)
    with tempfile.TemporaryDirectory() as tmp:
        ctx = mp.get_context("spawn")
        results = ctx.Manager().dict()
        procs = [ctx.Process(target=_worker, args=(r, world, _free_port() if r < 0 else PORT, tmp, results)) for r in range(world)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(300)
        assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
        assert dict(results) == {0: 1, 1: 1}


PORT = _free_port()
