"""Gather-scatter (direct stiffness summation) on a B200: libnompk's nompk_gs_* and the
nomp_b200_gs_* extension API against the CPU oracle (oracle/nomp_oracle.c: oracle_gs), bit for bit -- the association
order is part of the contract (copies in ascending local index, ranks in ascending rank order)."""
import ctypes as C
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

ROOT = Path(__file__).resolve().parent.parent

from libnomp_b200 import capi  # noqa: E402
from oracle import ffi  # noqa: E402

NP = {capi.F64: np.float64, capi.F32: np.float32, capi.I32: np.int32, capi.I64: np.int64, capi.U32: np.uint32,
      capi.U64: np.uint64}
TORCH = {capi.F64: torch.float64, capi.F32: torch.float32, capi.I32: torch.int32, capi.I64: torch.int64}


class Gs:
    """nompk_gs_* on one GPU (world = 1)."""

    def __init__(self, ids, kernel=None):
        """kernel: None = what the library picks (one copy per lane wherever no group has more than 32 copies),
        "group" = the one-group-per-thread kernel (NOMPK_GS_KERNEL=group at setup time)."""
        self.lib = capi.nompk()
        self.st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        ids_dev = torch.from_numpy(np.ascontiguousarray(ids, dtype=np.int64)).cuda()
        self.h = C.c_void_p()
        capi.nompk_check(self.lib.nompk_gs_create(ids_dev.data_ptr(), ids_dev.numel(), C.byref(self.h), self.st))
        xb = C.c_size_t(99)
        old = os.environ.pop("NOMPK_GS_KERNEL", None)
        if kernel:
            os.environ["NOMPK_GS_KERNEL"] = kernel
        try:
            capi.nompk_check(self.lib.nompk_gs_finalize_setup(self.h, 0, 1, C.byref(xb), self.st))
        finally:
            os.environ.pop("NOMPK_GS_KERNEL", None)
            if old is not None:
                os.environ["NOMPK_GS_KERNEL"] = old
        assert xb.value == 0

    def stats(self):
        out = (C.c_size_t * 8)()
        capi.nompk_check(self.lib.nompk_gs_stats(self.h, C.byref(out)))
        return [int(v) for v in out]

    def apply(self, op, dtype, v):
        capi.nompk_check(self.lib.nompk_gs_apply(self.h, op, dtype, v.data_ptr(), None, self.st))

    def close(self):
        self.lib.nompk_gs_destroy(self.h)


def values(dtype, n, seed):
    rng = np.random.default_rng(seed)
    if dtype in (capi.F64, capi.F32):
        return rng.uniform(0.5, 1.5, n).astype(NP[dtype])
    return rng.integers(-1000, 1000, n).astype(NP[dtype])


@pytest.mark.parametrize("kernel", [None, "group"])
@pytest.mark.parametrize("shape", [(4, 3, 2, 2), (8, 4, 3, 2), (2, 5, 5, 5), (6, 1, 1, 7), (8, 9, 7, 5)])
def test_box_mesh_multiplicities_and_sum(shape, kernel):
    """Summing ones gives the multiplicity of each point of a box of elements: 1 inside, 2 on faces, 4 on edges, 8 at
    corners shared by eight elements; random data agree with the oracle bit for bit."""
    n, ex, ey, ez = shape
    ids = ffi.box_ids(n, ex, ey, ez)
    gs = Gs(ids, kernel)
    st = gs.stats()
    N = n - 1
    points = (N * ex + 1) * (N * ey + 1) * (N * ez + 1)
    assert st[0] == ids.size and st[1] == points and st[4] == 0
    ones = torch.ones(ids.size, dtype=torch.float64, device="cuda")
    gs.apply(capi.RED_SUM, capi.F64, ones)
    want = ffi.gs(capi.RED_SUM, capi.F64, ids, np.ones(ids.size))
    assert np.array_equal(ones.cpu().numpy(), want)
    # every global point is counted once when weighted by 1 / multiplicity
    assert abs(float((1.0 / ones).sum()) - points) < 1e-9 * points
    v = values(capi.F64, ids.size, 3)
    d = torch.from_numpy(v).cuda()
    gs.apply(capi.RED_SUM, capi.F64, d)
    assert np.array_equal(d.cpu().numpy(), ffi.gs(capi.RED_SUM, capi.F64, ids, v.copy()))
    gs.close()


@pytest.mark.parametrize("big_group", [False, True])
@pytest.mark.parametrize("dtype", [capi.F64, capi.F32, capi.I32, capi.I64])
@pytest.mark.parametrize("op", [capi.RED_SUM, capi.RED_PROD, capi.RED_MIN, capi.RED_MAX])
def test_all_types_and_operators_on_random_numbering(dtype, op, big_group):
    """Ids drawn at random (groups of 1 .. ~20 copies, some ids <= 0 that must be left alone): folded by the warp kernel
    in serial order; with one group of 50 copies the handle falls back to the one-group-per-thread kernel."""
    rng = np.random.default_rng(17)
    n = 200003
    ids = rng.integers(-3, n // 8, n).astype(np.int64)
    if big_group:
        ids[rng.choice(n, 50, replace=False)] = 5
    v = values(dtype, n, 5)
    if op == capi.RED_PROD:
        v = (np.sign(v) * (1 + (np.abs(v) % 3) / 4)).astype(NP[dtype]) if dtype in (capi.F64, capi.F32) else (v % 3 + 1).astype(NP[dtype])
    gs = Gs(ids)
    d = torch.from_numpy(v).cuda()
    gs.apply(op, dtype, d)
    got = d.cpu().numpy()
    want = ffi.gs(op, dtype, ids, v.copy())
    assert np.array_equal(got, want)
    assert np.array_equal(got[ids <= 0], v[ids <= 0])
    if op in (capi.RED_MIN, capi.RED_MAX):  # idempotent
        gs.apply(op, dtype, d)
        assert np.array_equal(d.cpu().numpy(), want)
    gs.close()


def test_full_size_properties():
    """BASELINE size (64^3 elements, N = 7: 1.34e8 points, ids built on the device): size-independent properties in
    exact arithmetic -- multiplicities are 1/2/4/8 and sum(1/m) counts every distinct point once; the weighted sum of
    integer data is conserved; min is idempotent and never increases a value."""
    n, e = 8, 64
    N = n - 1
    px = N * e + 1
    el = torch.arange(e ** 3, device="cuda")
    pt = torch.arange(n ** 3, device="cuda")
    gx = (el % e)[:, None] * N + (pt % n)[None]
    gy = ((el // e) % e)[:, None] * N + ((pt // n) % n)[None]
    gz = (el // (e * e))[:, None] * N + (pt // (n * n))[None]
    ids = (1 + gx + px * (gy + px * gz)).reshape(-1).contiguous()
    del gx, gy, gz, el, pt
    lib = capi.nompk()
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    h = C.c_void_p()
    capi.nompk_check(lib.nompk_gs_create(ids.data_ptr(), ids.numel(), C.byref(h), st))
    del ids
    xb = C.c_size_t()
    capi.nompk_check(lib.nompk_gs_finalize_setup(h, 0, 1, C.byref(xb), st))
    stats = (C.c_size_t * 8)()
    lib.nompk_gs_stats(h, C.byref(stats))
    assert stats[0] == e ** 3 * n ** 3 and stats[1] == px ** 3
    m = torch.ones(stats[0], dtype=torch.float64, device="cuda")
    capi.nompk_check(lib.nompk_gs_apply(h, capi.RED_SUM, capi.F64, m.data_ptr(), None, st))
    counts = torch.bincount(m.to(torch.int64), minlength=9)
    assert int(counts[[0, 3, 5, 6, 7]].sum()) == 0
    assert int(counts[8]) == 8 * (e - 1) ** 3                                   # interior vertices of the box
    assert float((1.0 / m).sum()) == float(px ** 3)                            # dyadic terms: exact in any order
    v = torch.randint(-8, 9, (stats[0],), device="cuda").to(torch.float64)
    before = float((v).sum())
    capi.nompk_check(lib.nompk_gs_apply(h, capi.RED_SUM, capi.F64, v.data_ptr(), None, st))
    assert float((v / m).sum()) == before                                      # multiples of 1/8 below 2^40: exact
    w = torch.randint(-1000, 1000, (stats[0],), device="cuda").to(torch.float64)
    w0 = w.clone()
    capi.nompk_check(lib.nompk_gs_apply(h, capi.RED_MIN, capi.F64, w.data_ptr(), None, st))
    w1 = w.clone()
    capi.nompk_check(lib.nompk_gs_apply(h, capi.RED_MIN, capi.F64, w.data_ptr(), None, st))
    assert torch.equal(w, w1) and bool((w1 <= w0).all()) and bool((w1[m == 1] == w0[m == 1]).all())
    lib.nompk_gs_destroy(h)


def test_golden_fixture():
    """tests/golden/gs_cases.json (independent numpy restatement, exact-integer data): the GPU agrees with it too."""
    import json
    ops = {"+": capi.RED_SUM, "*": capi.RED_PROD, "min": capi.RED_MIN, "max": capi.RED_MAX}
    dts = {"float64": capi.F64, "float32": capi.F32, "int64": capi.I64, "int32": capi.I32}
    cases = json.loads((ROOT / "tests" / "golden" / "gs_cases.json").read_text())["cases"]
    handles = {}
    for c in cases:
        if c["name"] not in handles:
            handles[c["name"]] = Gs(np.array(c["ids"], dtype=np.int64))
        d = torch.from_numpy(np.array(c["v"], dtype=c["dtype"])).cuda()
        handles[c["name"]].apply(ops[c["op"]], dts[c["dtype"]], d)
        assert np.array_equal(d.cpu().numpy(), np.array(c["want"], dtype=c["dtype"])), (c["name"], c["op"], c["dtype"])
    for h in handles.values():
        h.close()


def test_degenerate_numberings():
    # no shared id at all: nothing to do, nothing launched
    ids = np.arange(1, 1001, dtype=np.int64)
    gs = Gs(ids)
    assert gs.stats()[2] == 0
    v = torch.arange(1000, dtype=torch.float64, device="cuda")
    before = capi.nompk().nompk_launch_count()
    gs.apply(capi.RED_SUM, capi.F64, v)
    assert capi.nompk().nompk_launch_count() == before
    assert np.array_equal(v.cpu().numpy(), np.arange(1000.0))
    gs.close()
    # one id everywhere: a single group of n copies
    n = 5000
    gs = Gs(np.full(n, 7, dtype=np.int64))
    assert gs.stats()[2:4] == [1, n]
    v = torch.from_numpy(ffi.fill_int_f64(n, 3, 0, 7)).cuda()
    total = float(v.sum())
    gs.apply(capi.RED_SUM, capi.F64, v)
    assert np.all(v.cpu().numpy() == total)
    gs.close()
    # empty
    gs = Gs(np.zeros(0, dtype=np.int64))
    gs.apply(capi.RED_SUM, capi.F64, torch.zeros(1, dtype=torch.float64, device="cuda"))
    gs.close()


def test_public_api_on_mapped_vectors():
    capi.check(capi.init(backend="cuda", device=0, verbose=0))
    lib = capi.nomp()
    try:
        n, ex, ey, ez = 8, 6, 5, 4
        ids = ffi.box_ids(n, ex, ey, ez)
        h = capi.gs_setup(ids)
        info = capi.gs_info(h)
        assert info["n"] == ids.size and info["distinct"] == (7 * ex + 1) * (7 * ey + 1) * (7 * ez + 1)
        assert info["shared_groups"] == 0 and info["neighbours"] == 0
        v = ffi.fill_uniform_f64(ids.size, 9, 0.5, 1.5)
        w = ffi.fill_i64(ids.size, 4)
        want_v = ffi.gs(capi.RED_SUM, capi.F64, ids, v.copy())
        want_w = ffi.gs(capi.RED_MAX, capi.I64, ids, w.copy())
        for a in (v, w):
            capi.check(capi.update(a.ctypes.data, 0, a.size, 8, capi.NOMP_TO))
        capi.gs(h, v, "+")
        capi.gs(h, w, "max")
        for a in (v, w):
            capi.check(capi.update(a.ctypes.data, 0, a.size, 8, capi.NOMP_FROM))
        assert np.array_equal(v, want_v) and np.array_equal(w, want_w)

        # errors: unmapped vector, short mapping, unknown operator, stale handle
        other = np.zeros(ids.size)
        err = lib.nomp_b200_gs(h, other.ctypes.data, 8, capi.NOMP_FLOAT, b"+")
        assert capi.err_info(err)[0] == capi.NOMP_USER_MAP_PTR_IS_INVALID
        short = np.zeros(10)
        capi.check(capi.update(short.ctypes.data, 0, 10, 8, capi.NOMP_TO))
        err = lib.nomp_b200_gs(h, short.ctypes.data, 8, capi.NOMP_FLOAT, b"+")
        assert capi.err_info(err)[0] == capi.NOMP_USER_INPUT_IS_INVALID
        err = lib.nomp_b200_gs(h, v.ctypes.data, 8, capi.NOMP_FLOAT, b"-")
        assert capi.err_info(err)[0] == capi.NOMP_USER_INPUT_IS_INVALID
        capi.check(lib.nomp_b200_gs_free(h))
        err = lib.nomp_b200_gs(h, v.ctypes.data, 8, capi.NOMP_FLOAT, b"+")
        assert capi.err_info(err)[0] == capi.NOMP_USER_INPUT_IS_INVALID
        h2 = capi.gs_setup(ids[: 512 * 4])  # left for nomp_finalize to release
        assert h2 == h
    finally:
        assert lib.nomp_finalize_excluding_interpreter() == 0


GS_WORKER = r"""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, {root!r})
rank, world = int(sys.argv[1]), int(sys.argv[2])
os.environ.update(NOMP_COMM_SIZE=str(world), NOMP_COMM_RANK=str(rank), NOMP_COMM_ID_FILE=sys.argv[3])
from libnomp_b200 import capi
from oracle import ffi
capi.check(capi.init(backend="cuda", device=rank, verbose=1))
lib = capi.nomp()
n, ex, ey, ezr = 8, 6, 5, 3                      # every rank owns a slab of ezr element layers
ids_all = ffi.box_ids(n, ex, ey, ezr * world)
per = ids_all.size // world
seg = [r * per for r in range(world + 1)]
ids = ids_all[seg[rank]:seg[rank + 1]].copy()
h = capi.gs_setup(ids)
info = capi.gs_info(h)
plane = (7 * ex + 1) * (7 * ey + 1)
assert info["neighbours"] == (1 if rank in (0, world - 1) else 2), info
assert info["shared_ids"] == plane * info["neighbours"], info
for rep, (op, name, dtype, fill) in enumerate([(capi.RED_SUM, "+", capi.F64, lambda m: ffi.fill_uniform_f64(m, 21, 0.5, 1.5)),
                                               (capi.RED_MAX, "max", capi.I64, lambda m: ffi.fill_i64(m, 22)),
                                               (capi.RED_SUM, "+", capi.F64, lambda m: ffi.fill_int_f64(m, 23, -4, 4)),
                                               (capi.RED_MIN, "min", capi.F64, lambda m: ffi.fill_uniform_f64(m, 24, 0.5, 1.5))]):
    full = fill(ids_all.size)
    want = ffi.gs(op, dtype, ids_all, full.copy(), seg)[seg[rank]:seg[rank + 1]]
    mine = full[seg[rank]:seg[rank + 1]].copy()
    capi.check(capi.update(mine.ctypes.data, 0, mine.size, 8, capi.NOMP_TO))
    for _ in range(3 if rep == 3 else 1):        # min is idempotent: repeated calls alternate the exchange slots
        capi.gs(h, mine, name)
    capi.check(capi.update(mine.ctypes.data, 0, mine.size, 8, capi.NOMP_FROM))
    capi.check(capi.update(mine.ctypes.data, 0, mine.size, 8, capi.NOMP_FREE))
    assert np.array_equal(mine, want), (rank, rep, int((mine != want).sum()))
# a numbering in which some ids live on every rank (corners of a partition in more than one direction)
rng = np.random.default_rng(5)
ids2_all = rng.integers(1, 4000, 30000 * world).astype(np.int64)
seg2 = [r * 30000 for r in range(world + 1)]
h2 = capi.gs_setup(ids2_all[seg2[rank]:seg2[rank + 1]].copy())
full = ffi.fill_uniform_f64(ids2_all.size, 31, 0.5, 1.5)
want = ffi.gs(capi.RED_SUM, capi.F64, ids2_all, full.copy(), seg2)[seg2[rank]:seg2[rank + 1]]
mine = full[seg2[rank]:seg2[rank + 1]].copy()
capi.check(capi.update(mine.ctypes.data, 0, mine.size, 8, capi.NOMP_TO))
capi.gs(h2, mine, "+")
capi.check(capi.update(mine.ctypes.data, 0, mine.size, 8, capi.NOMP_FROM))
assert np.array_equal(mine, want), (rank, "all-to-all numbering", int((mine != want).sum()))
capi.check(lib.nomp_b200_gs_free(h2))
assert lib.nomp_finalize_excluding_interpreter() == 0
print("rank", rank, "ok")
"""


def test_gather_scatter_across_gpus(tmp_path):
    """Slab-partitioned box mesh on 2-4 GPUs: the interface planes are summed over NVLink peer memory and every rank
    gets the bits of the oracle's two-level (rank-ordered) fold."""
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs at least two GPUs")
    script = tmp_path / "worker.py"
    script.write_text(GS_WORKER.format(root=str(ROOT)))
    idfile = f"/dev/shm/nomp-test-gs-{os.getpid()}"
    procs = [subprocess.Popen([sys.executable, str(script), str(r), str(world), idfile], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(world)]
    outs = [p.communicate(timeout=600)[0] for p in procs]
    for f in [idfile] + [f"{idfile}.ipc.{r}" for r in range(world)]:
        try:
            os.unlink(f)
        except OSError:
            pass
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and f"rank {r} ok" in o, o[-3000:]


class EmulatedRanks:
    """The multi-rank protocol of nompk_gs_* with every "rank" on this one GPU: one handle, one exchange buffer and
    one stream per rank.  The kernels of different ranks run concurrently and talk through the same flags and slots
    they use over NVLink, so layouts and offsets for any number of ranks are testable without that many GPUs."""

    def __init__(self, id_parts):
        self.lib, self.world = capi.nompk(), len(id_parts)
        lib, world = self.lib, self.world
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        self.streams = [torch.cuda.Stream() for _ in range(world)]
        self.handles, uniq = [], []
        for ids in id_parts:
            d = torch.from_numpy(np.ascontiguousarray(ids, dtype=np.int64)).cuda()
            h = C.c_void_p()
            capi.nompk_check(lib.nompk_gs_create(d.data_ptr(), d.numel(), C.byref(h), st))
            ptr, cnt = C.c_void_p(), C.c_size_t()
            capi.nompk_check(lib.nompk_gs_unique(h, C.byref(ptr), C.byref(cnt)))
            self.handles.append(h)
            uniq.append((ptr, cnt.value))
        self.shared = [[0] * world for _ in range(world)]
        for r, h in enumerate(self.handles):
            for p in range(world):
                if p != r:
                    n = C.c_size_t()
                    capi.nompk_check(lib.nompk_gs_match_peer(h, p, world, uniq[p][0], uniq[p][1], C.byref(n), st))
                    self.shared[r][p] = n.value
        self.buffers, offsets, totals = [], [], []
        for r, h in enumerate(self.handles):
            xb = C.c_size_t()
            capi.nompk_check(lib.nompk_gs_finalize_setup(h, r, world, C.byref(xb), st))
            self.buffers.append(torch.zeros(max(int(xb.value), 8), dtype=torch.uint8, device="cuda"))
            off, cnt = (C.c_size_t * world)(), (C.c_size_t * world)()
            capi.nompk_check(lib.nompk_gs_recv_offsets(h, off, cnt))
            assert list(cnt) == self.shared[r]
            offsets.append(list(off))
            totals.append(sum(cnt))
        for r, h in enumerate(self.handles):
            peers = (C.c_void_p * world)(*[b.data_ptr() for b in self.buffers])
            send = (C.c_size_t * world)(*[offsets[p][r] for p in range(world)])
            tot = (C.c_size_t * world)(*totals)
            capi.nompk_check(lib.nompk_gs_connect(h, peers, send, tot, st))
        torch.cuda.synchronize()

    def apply(self, op, dtype, parts):
        dev = [torch.from_numpy(p).cuda() for p in parts]
        torch.cuda.synchronize()
        for r, h in enumerate(self.handles):
            capi.nompk_check(self.lib.nompk_gs_apply(h, op, dtype, dev[r].data_ptr(), None,
                                                     C.c_void_p(self.streams[r].cuda_stream)))
        torch.cuda.synchronize()
        return [d.cpu().numpy() for d in dev]

    def close(self):
        for h in self.handles:
            self.lib.nompk_gs_destroy(h)


@pytest.mark.parametrize("world", [2, 3, 4, 5, 8])
def test_multi_rank_protocol_on_one_gpu(world):
    """Slabs of a box mesh (inner ranks have two neighbours with segments of their own in the exchange buffer) and a
    random numbering whose ids live on any subset of the ranks; repeated calls alternate the two slots."""
    n, ex, ey, ezr = 6, 4, 3, 2
    ids_all = ffi.box_ids(n, ex, ey, ezr * world)
    per = ids_all.size // world
    seg = [r * per for r in range(world + 1)]
    ranks = EmulatedRanks([ids_all[seg[r]:seg[r + 1]] for r in range(world)])
    plane = (5 * ex + 1) * (5 * ey + 1)
    for r in range(world):
        assert sum(ranks.shared[r]) == plane * ((r > 0) + (r < world - 1))
    for rep, (op, dtype) in enumerate([(capi.RED_SUM, capi.F64), (capi.RED_MAX, capi.I64), (capi.RED_SUM, capi.F64),
                                       (capi.RED_MIN, capi.F32), (capi.RED_SUM, capi.F64)]):
        full = values(dtype, ids_all.size, 60 + rep)
        want = ffi.gs(op, dtype, ids_all, full.copy(), seg)
        got = ranks.apply(op, dtype, [full[seg[r]:seg[r + 1]].copy() for r in range(world)])
        for r in range(world):
            assert np.array_equal(got[r], want[seg[r]:seg[r + 1]]), (world, rep, r)
    ranks.close()

    rng = np.random.default_rng(world)
    m = 20000
    ids2 = rng.integers(1, 3000, m * world).astype(np.int64)
    ids2[rng.integers(0, ids2.size, 500)] = 0
    seg2 = [r * m for r in range(world + 1)]
    ranks = EmulatedRanks([ids2[seg2[r]:seg2[r + 1]] for r in range(world)])
    for rep in range(3):
        full = values(capi.F64, ids2.size, 80 + rep)
        want = ffi.gs(capi.RED_SUM, capi.F64, ids2, full.copy(), seg2)
        got = ranks.apply(capi.RED_SUM, capi.F64, [full[seg2[r]:seg2[r + 1]].copy() for r in range(world)])
        for r in range(world):
            assert np.array_equal(got[r], want[seg2[r]:seg2[r + 1]]), (world, rep, r)
    ranks.close()
