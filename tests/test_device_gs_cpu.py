"""The device kernels of the gather-scatter (libnomp_b200/csrc/kernels/gs.cu: gs_local_kernel -- one group per thread --,
gs_local_warp_kernel -- one copy per lane, groups folded with shuffles -- and gs_remote_kernel)
compiled for the HOST -- their text is cut out of gs.cu unchanged -- and executed with the cooperative emulator, one
thread per rank and host memory as "peer memory".  The index structures the CUDA setup builds with CUB are rebuilt here
with numpy from their documented meaning, so this checks the protocol between the ranks (segments and slot strides in
the peers' buffers, flags, the ticket among the CTAs that hold shared groups, the rank-ordered fold) against the oracle
without a GPU; the setup itself is covered by the GPU tests."""
import ctypes as C
import re
import threading
from pathlib import Path

import numpy as np
import pytest

from oracle import ffi
from tests import cuda_emulation as emu

ROOT = Path(__file__).resolve().parent.parent
GS_CU = ROOT / "libnomp_b200" / "csrc" / "kernels" / "gs.cu"

PRELUDE = r"""
enum { NOMPK_RED_SUM = 0, NOMPK_RED_PROD = 1, NOMPK_RED_MIN = 2, NOMPK_RED_MAX = 3 };
template <typename T> static inline T op_add(T a, T b) { return a + b; }
template <typename T> static inline T op_mul(T a, T b) { return a * b; }
constexpr int kGsThreads = 256;
constexpr int kGsWarps = 256 / 32;
constexpr unsigned long long kGsTimeoutNs = 20ull * 1000 * 1000 * 1000;
static inline unsigned long long timer_ns() { return nomp_emu_now_ns(); }
"""
WRAPPERS = r"""
static void local_sum_f64(double *v, GsView s) { gs_local_kernel<NOMPK_RED_SUM, double>(v, s); }
static void warp_sum_f64(double *v, GsView s) { gs_local_warp_kernel<NOMPK_RED_SUM, double, 4>(v, s); }
static void warp_max_i64(long long *v, GsView s) { gs_local_warp_kernel<NOMPK_RED_MAX, long long, 4>(v, s); }
static void remote_sum_f64(double *v, GsView s) { gs_remote_kernel<NOMPK_RED_SUM, double>(v, s); }
static void local_max_i64(long long *v, GsView s) { gs_local_kernel<NOMPK_RED_MAX, long long>(v, s); }
static void remote_max_i64(long long *v, GsView s) { gs_remote_kernel<NOMPK_RED_MAX, long long>(v, s); }
"""


def device_source():
    text = GS_CU.read_text()
    combine = re.search(r"template <int OP, typename T> __device__ __forceinline__ T combine\(T a, T b\) \{.*?\n\}\n", text, re.S)
    a = text.index("struct GsView {")
    b = text.index("template <int OP, typename T> int launch_gs(")
    assert combine and a < b
    return PRELUDE + combine.group(0) + text[a:b] + WRAPPERS


class GsView(C.Structure):
    _fields_ = [("offsets", C.c_void_p), ("indices", C.c_void_p), ("pidx", C.c_void_p), ("heads", C.c_void_p), ("wgroup", C.c_void_p),
                ("rowinfo", C.c_void_p), ("nwarps", C.c_size_t), ("remote_slot", C.c_void_p), ("rgroup", C.c_void_p),
                ("roffsets", C.c_void_p), ("rpos", C.c_void_p), ("rpeer", C.c_void_p), ("partial", C.c_void_p),
                ("ticket", C.c_void_p), ("recv_off", C.c_void_p), ("send_off", C.c_void_p), ("peer_xchg", C.c_void_p),
                ("neighbours", C.c_void_p), ("G", C.c_size_t), ("Q", C.c_size_t), ("values_base", C.c_size_t),
                ("remote_ctas", C.c_uint), ("flags_bytes", C.c_size_t), ("n_neighbours", C.c_int), ("rank", C.c_int),
                ("world", C.c_int), ("slot", C.c_int), ("seq", C.c_ulonglong), ("error_host", C.c_void_p)]


def flags_bytes(world):
    return (2 * world * 8 + 255) // 256 * 256


class Rank:
    """What nompk_gs_create / match_peer / finalize_setup / connect produce for one rank, built with numpy."""

    def __init__(self, rank, id_parts):
        world = len(id_parts)
        ids = id_parts[rank]
        self.rank, self.world = rank, world
        uniq, counts = np.unique(ids[ids > 0], return_counts=True)
        order = np.argsort(ids, kind="stable")            # copies of an id in ascending local index
        sorted_ids = ids[order]
        starts = np.searchsorted(sorted_ids, uniq)
        self.shared = []                                  # per peer: ascending ids both ranks hold = S(me, peer)
        for r in range(world):
            other = np.unique(id_parts[r][id_parts[r] > 0])
            self.shared.append(np.intersect1d(uniq, other) if r != rank else np.zeros(0, dtype=np.int64))
        npeers = np.zeros(uniq.size, dtype=np.int64)
        for r in range(world):
            npeers += np.isin(uniq, self.shared[r])
        active = (counts >= 2) | (npeers > 0)
        act = np.flatnonzero(active)
        first_idx = order[starts[act]]
        # groups shared with a peer first (their partial results leave early), then by their first local copy
        act = act[np.lexsort((first_idx, npeers[act] == 0))]
        offsets, indices, remote_slot, rgroup, roffsets, rpeer, rpos = [0], [], [], [], [0], [], []
        for g, u in enumerate(act):
            idx = order[starts[u]:starts[u] + counts[u]]
            indices += idx.tolist()
            offsets.append(len(indices))
            if npeers[u] == 0:
                remote_slot.append(-1)
                continue
            remote_slot.append(len(rgroup))
            rgroup.append(g)
            for r in range(world):
                if r != rank:
                    pos = np.searchsorted(self.shared[r], uniq[u])
                    if pos < self.shared[r].size and self.shared[r][pos] == uniq[u]:
                        rpeer.append(r)
                        rpos.append(int(pos))
            roffsets.append(len(rpeer))
        u32, i32 = (lambda a: np.array(a, dtype=np.uint32)), (lambda a: np.array(a, dtype=np.int32))
        self.offsets, self.indices, self.remote_slot = u32(offsets), u32(indices + [0]), i32(remote_slot + [0])
        self.rgroup, self.roffsets, self.rpeer, self.rpos = u32(rgroup + [0]), u32(roffsets), i32(rpeer + [0]), u32(rpos + [0])
        self.G, self.Q = len(remote_slot), len(rgroup)
        self.counts = [int(s.size) for s in self.shared]
        self.recv_off = np.concatenate([[0], np.cumsum(self.counts)[:-1]]).astype(np.uint64)
        self.total = int(sum(self.counts))
        self.neighbours = i32([r for r in range(world) if self.counts[r] > 0] + [0])
        self.n_neighbours = sum(1 for c in self.counts if c > 0)
        blocks = {g // 256 for g, q in enumerate(remote_slot) if q >= 0}
        self.remote_ctas = len(blocks)
        # warp layout: tiles of 256 groups start at multiples of 32 slots; inside a tile a group that would straddle a
        # multiple of 32 moves to the next one; heads = first copy of each group; wgroup = group of a warp's first slot
        pidx, heads, wgroup, rowinfo = [], [], [], []
        sizes = np.diff(self.offsets)
        self.warp_layout = sizes.size > 0 and int(sizes.max()) <= 32
        if self.warp_layout:
            for t0 in range(0, self.G, 256):
                pos = len(pidx)
                assert pos % 32 == 0
                for g in range(t0, min(self.G, t0 + 256)):
                    m = int(sizes[g])
                    if len(pidx) % 32 + m > 32:
                        pidx += [0xffffffff] * (-len(pidx) % 32)
                    at = len(pidx)
                    while len(heads) <= at // 32:
                        heads.append(0)
                        wgroup.append(0xffffffff)
                        rowinfo.append(0)
                    rowinfo[at // 32] = max(rowinfo[at // 32] & 0xff, m) | ((at % 32 + m) << 8)   # longest group | occupied slots << 8
                    if at % 32 == 0:
                        wgroup[at // 32] = g
                    heads[at // 32] |= 1 << (at % 32)
                    pidx += indices[offsets[g]:offsets[g + 1]]
                pidx += [0xffffffff] * (-len(pidx) % 32)
        self.pidx, self.heads, self.wgroup = u32(pidx + [0]), u32(heads + [0]), u32(wgroup + [0])
        self.rowinfo = np.array(rowinfo + [0], dtype=np.uint16)
        self.nwarps = len(heads)
        self.remote_ctas_warp = len({w // 32 for w in range(self.nwarps) if wgroup[w] < len(rgroup)})
        self.partial = np.zeros(max(self.Q, 1), dtype=np.uint64)
        self.ticket = np.zeros(2, dtype=np.uint32)
        self.xchg = np.zeros((flags_bytes(world) + 2 * self.total * 8) // 8 + 1, dtype=np.uint64)
        self.error = np.zeros(1, dtype=np.uint64)
        self.seq = 0

    def connect(self, ranks):
        world = self.world
        self.table = np.array([r.xchg.ctypes.data for r in ranks], dtype=np.uint64)
        send = np.zeros(2 * world, dtype=np.uint64)
        for r in range(world):
            send[r] = ranks[r].recv_off[self.rank]
            send[world + r] = ranks[r].total + int(ranks[r].recv_off[self.rank])
        self.send_off = send

    def view(self, warp=False):
        self.seq += 1
        slot = self.seq & 1
        p = lambda a: a.ctypes.data  # noqa: E731
        return GsView(p(self.offsets), p(self.indices), p(self.pidx), p(self.heads), p(self.wgroup), p(self.rowinfo), self.nwarps, p(self.remote_slot),
                      p(self.rgroup), p(self.roffsets), p(self.rpos),
                      p(self.rpeer), p(self.partial), p(self.ticket), p(self.recv_off), p(self.send_off), p(self.table),
                      p(self.neighbours), self.G, self.Q, flags_bytes(self.world) + slot * self.total * 8,
                      self.remote_ctas_warp if warp else self.remote_ctas,
                      flags_bytes(self.world), self.n_neighbours, self.rank, self.world, slot, self.seq, p(self.error))


def apply(ranks, parts, kind, warp=None):
    """`warp`: per rank, use gs_local_warp_kernel (default: wherever the layout exists, as the library does; a list mixes
    the two local kernels between the ranks -- they speak one protocol)."""
    T = {"sum_f64": "double", "max_i64": "long long"}[kind]
    src = device_source()
    errors = []

    def main(r):
        try:
            rk = ranks[r]
            if rk.G == 0:
                return
            use_warp = rk.warp_layout if warp is None else (warp[r] and rk.warp_layout)
            view = rk.view(use_warp)
            v = C.c_void_p(parts[r].ctypes.data)
            if use_warp:
                emu.emulate_cooperative(src, f"warp_{kind}", ((rk.nwarps + 31) // 32, 1, 1), (256, 1, 1), [f"{T} *", "GsView"], [v, view],
                                        instance=20 + r)
            else:
                emu.emulate_cooperative(src, f"local_{kind}", ((rk.G + 255) // 256, 1, 1), (256, 1, 1), [f"{T} *", "GsView"], [v, view],
                                        instance=20 + r)
            if rk.Q:
                emu.emulate_cooperative(src, f"remote_{kind}", ((rk.Q + 255) // 256, 1, 1), (256, 1, 1), [f"{T} *", "GsView"],
                                        [v, view], instance=20 + r)
        except BaseException as exc:   # pragma: no cover
            errors.append(exc)

    threads = [threading.Thread(target=main, args=(r,)) for r in range(len(ranks))]
    for t in threads:
        t.start()
    for t in threads:
        t.join(120)
    assert not errors and not any(t.is_alive() for t in threads), errors
    assert all(int(r.error[0]) == 0 and int(r.ticket[0]) == 0 for r in ranks)


@pytest.mark.parametrize("world", [1, 2, 3, 4])
def test_gs_kernels_between_host_ranks(world):
    n, ex, ey, ezr = 4, 3, 2, 2
    ids_all = ffi.box_ids(n, ex, ey, ezr * world)
    per = ids_all.size // world
    seg = [r * per for r in range(world + 1)]
    id_parts = [ids_all[seg[r]:seg[r + 1]].copy() for r in range(world)]
    ranks = [Rank(r, id_parts) for r in range(world)]
    for r in ranks:
        r.connect(ranks)
    plane = (3 * ex + 1) * (3 * ey + 1)
    assert [rk.total for rk in ranks] == [plane * ((r > 0) + (r < world - 1)) for r in range(world)]
    rng = np.random.default_rng(world)
    for call in range(3):                      # consecutive calls alternate the two slots of the exchange buffers
        full = rng.uniform(0.5, 1.5, ids_all.size)
        want = ffi.gs(0, ffi.F64, ids_all, full.copy(), seg)
        parts = [full[seg[r]:seg[r + 1]].copy() for r in range(world)]
        # call 0: every rank with the warp kernel, call 1: every rank with the group kernel, call 2: mixed
        apply(ranks, parts, "sum_f64", warp=[None, [False] * world, [r % 2 == 0 for r in range(world)]][call])
        for r in range(world):
            assert np.array_equal(parts[r], want[seg[r]:seg[r + 1]]), (world, call, r)
    assert all(rk.warp_layout for rk in ranks)
    full = rng.integers(-1000, 1000, ids_all.size).astype(np.int64)
    want = ffi.gs(3, ffi.I64, ids_all, full.copy(), seg)
    parts = [full[seg[r]:seg[r + 1]].copy() for r in range(world)]
    apply(ranks, parts, "max_i64")
    for r in range(world):
        assert np.array_equal(parts[r], want[seg[r]:seg[r + 1]])


def test_gs_kernels_with_ids_on_every_rank():
    """A numbering in which an id may live on any subset of three ranks, with ids <= 0 that do not take part: segments
    of different sizes per peer, groups shared with two peers at once."""
    world, m = 3, 4000
    rng = np.random.default_rng(9)
    ids = rng.integers(-1, 900, m * world).astype(np.int64)
    seg = [r * m for r in range(world + 1)]
    id_parts = [ids[seg[r]:seg[r + 1]].copy() for r in range(world)]
    ranks = [Rank(r, id_parts) for r in range(world)]
    for r in ranks:
        r.connect(ranks)
    for call in range(2):
        full = rng.uniform(0.5, 1.5, ids.size)
        want = ffi.gs(0, ffi.F64, ids, full.copy(), seg)
        parts = [full[seg[r]:seg[r + 1]].copy() for r in range(world)]
        apply(ranks, parts, "sum_f64")
        for r in range(world):
            assert np.array_equal(parts[r], want[seg[r]:seg[r + 1]]), (call, r)


# ---- the setup kernels: CUB's sort / run-length / scan / select are played by numpy, the library's own kernels run ----------

SETUP_KERNELS = {
    "iota_kernel": ["unsigned *", "size_t"],
    "match_kernel": ["const long long *", "size_t", "const long long *", "size_t", "unsigned *"],
    "position_kernel": ["const unsigned *", "unsigned *", "size_t"],
    "classify_kernel": ["const long long *", "const unsigned *", "unsigned **", "int", "size_t", "unsigned char *", "unsigned *"],
    "first_index_kernel": ["const unsigned *", "const unsigned *", "const unsigned *", "const unsigned *", "unsigned long long *", "size_t"],
    "group_sizes_kernel": ["const unsigned *", "const unsigned *", "const unsigned *", "unsigned *", "unsigned *", "unsigned *", "size_t"],
    "fill_kernel": ["const unsigned *", "const unsigned *", "const unsigned *", "const unsigned *", "const unsigned *", "const unsigned *",
                    "const unsigned *", "const unsigned *", "unsigned **", "int", "unsigned *", "int *", "unsigned *", "unsigned *", "int *",
                    "unsigned *", "size_t"],
    "mark_remote_ctas_kernel": ["const int *", "size_t", "int", "unsigned *"],
}
LAYOUT_KERNELS = {
    "layout_size_kernel": ["const unsigned *", "size_t", "unsigned *"],
    "layout_fill_kernel": ["const unsigned *", "const unsigned *", "size_t", "const unsigned *", "unsigned *", "unsigned *", "unsigned *",
                           "unsigned short *"],
    "mark_remote_warp_ctas_kernel": ["const unsigned *", "size_t", "size_t", "int", "unsigned *"],
}
SETUP_KERNELS.update(LAYOUT_KERNELS)


def setup_source():
    text = GS_CU.read_text()
    a, b = text.index("// ---- setup kernels"), text.index("// ---- apply")      # includes the warp-layout kernels
    return "#include <cstddef>\n" + text[a:b]


def launch(name, n_threads, *args):
    cargs = []
    for t, a in zip(SETUP_KERNELS[name], args):
        if isinstance(a, np.ndarray):
            cargs.append(C.c_void_p(a.ctypes.data))
        elif t == "int":
            cargs.append(C.c_int(a))
        else:
            cargs.append(C.c_size_t(a))
    emu.emulate_cooperative(setup_source(), name, (max(1, (n_threads + 255) // 256), 1, 1), (256, 1, 1), SETUP_KERNELS[name], cargs,
                            instance=60)


def excl(a):
    return np.concatenate([[0], np.cumsum(a)]).astype(np.uint32)


def device_setup(rank, id_parts):
    """nompk_gs_create / match_peer / finalize_setup step by step, as gs.cu orders them."""
    world, ids = len(id_parts), id_parts[rank]
    n = ids.size
    vals = np.zeros(n, dtype=np.uint32)
    launch("iota_kernel", n, vals, n)
    assert np.array_equal(vals, np.arange(n))
    order = np.argsort(ids, kind="stable")                               # cub::DeviceRadixSort::SortPairs (stable)
    keys, sorted_idx = ids[order], vals[order]
    uniq, counts = np.unique(keys, return_counts=True)                    # cub::DeviceRunLengthEncode::Encode
    uniq, counts = uniq.astype(np.int64), counts.astype(np.uint32)
    run_start = excl(counts)                                             # cub::DeviceScan::ExclusiveSum
    U = uniq.size
    peer_pos, shared = [None] * world, [0] * world
    for r in range(world):
        if r == rank:
            continue
        other = np.unique(id_parts[r]).astype(np.int64)
        found, pos = np.zeros(U, dtype=np.uint32), None
        launch("match_kernel", U, uniq, U, other, other.size, found)
        pos = excl(found)
        shared[r] = int(pos[U])
        launch("position_kernel", U, found, pos, U)
        peer_pos[r] = pos if shared[r] else None
    table = np.array([p.ctypes.data if p is not None else 0 for p in peer_pos], dtype=np.uint64)
    active, rcount = np.zeros(U, dtype=np.uint8), np.zeros(U, dtype=np.uint32)
    launch("classify_kernel", U, uniq, counts, table, world, U, active, rcount)
    sel = np.flatnonzero(active).astype(np.uint32)                        # cub::DeviceSelect::Flagged
    G = sel.size
    first = np.zeros(max(G, 1), dtype=np.uint64)
    launch("first_index_kernel", G, sel, run_start, sorted_idx, rcount, first, G)
    order_g = sel[np.argsort(first[:G], kind="stable")]                  # SortPairs(first, sel)
    cnt, rcnt, rflag = (np.zeros(max(G, 1), dtype=np.uint32) for _ in range(3))
    launch("group_sizes_kernel", G, order_g, counts, rcount, cnt, rcnt, rflag, G)
    offsets, rstart, rslot = excl(cnt[:G]), excl(rcnt[:G]), excl(rflag[:G])
    nnz, R, Q = int(offsets[G]), int(rstart[G]), int(rslot[G])
    indices, remote_slot = np.zeros(nnz + 1, dtype=np.uint32), np.zeros(G + 1, dtype=np.int32)
    rgroup, roffsets = np.zeros(Q + 1, dtype=np.uint32), np.zeros(Q + 1, dtype=np.uint32)
    rpeer, rpos = np.zeros(R + 1, dtype=np.int32), np.zeros(R + 1, dtype=np.uint32)
    launch("fill_kernel", G, order_g, run_start, counts, sorted_idx, offsets, rcnt, rstart, rslot, table, world, indices, remote_slot,
           rgroup, roffsets, rpeer, rpos, G)
    roffsets[Q] = R
    blk = np.zeros((G + 255) // 256 + 1, dtype=np.uint32)
    launch("mark_remote_ctas_kernel", G, remote_slot, G, 256, blk)
    out = dict(G=G, Q=Q, offsets=offsets, indices=indices[:nnz], remote_slot=remote_slot[:G], rgroup=rgroup[:Q], roffsets=roffsets,
               rpeer=rpeer[:R], rpos=rpos[:R], shared=shared, remote_ctas=int(blk.sum()), nwarps=0)
    if G and int(cnt[:G].max()) <= 32:                                    # cub::DeviceReduce::Max
        ntiles = (G + 255) // 256
        tile_warps = np.zeros(ntiles, dtype=np.uint32)
        launch("layout_size_kernel", ntiles, offsets, G, tile_warps)
        tile_start = excl(tile_warps)
        nwarps = int(tile_start[ntiles])
        pidx = np.full(nwarps * 32 + 1, 0xffffffff, dtype=np.uint32)
        heads, wgroup = np.zeros(nwarps + 1, dtype=np.uint32), np.full(nwarps + 1, 0xffffffff, dtype=np.uint32)
        rowinfo = np.zeros(nwarps + 1, dtype=np.uint16)
        launch("layout_fill_kernel", ntiles, offsets, indices, G, tile_start, pidx, heads, wgroup, rowinfo)
        wblk = np.zeros((nwarps + 31) // 32 + 1, dtype=np.uint32)
        launch("mark_remote_warp_ctas_kernel", nwarps, wgroup, nwarps, Q, 32, wblk)
        out.update(nwarps=nwarps, pidx=pidx[:nwarps * 32], heads=heads[:nwarps], wgroup=wgroup[:nwarps], rowinfo=rowinfo[:nwarps],
                   remote_ctas_warp=int(wblk.sum()))
    return out


@pytest.mark.parametrize("kind", ["slabs", "random"])
def test_setup_kernels_build_the_documented_structures(kind):
    """The library's setup kernels (with numpy standing in for CUB) produce, for every rank, exactly the groups, CSR
    arrays, peer positions and remote-CTA count that the `Rank` class above derives from their documented meaning --
    so the protocol tests above and the real setup talk about the same structures."""
    world = 3
    if kind == "slabs":
        ids_all = ffi.box_ids(4, 3, 2, 2 * world)
        per = ids_all.size // world
        id_parts = [ids_all[r * per:(r + 1) * per].copy() for r in range(world)]
    else:
        rng = np.random.default_rng(4)
        id_parts = [rng.integers(-1, 700, 3000).astype(np.int64) for _ in range(world)]
        id_parts[1][100:140] = 77            # a group of more than 32 copies: rank 1 keeps the one-group-per-thread kernel
    for rank in range(world):
        got, want = device_setup(rank, id_parts), Rank(rank, id_parts)
        assert got["G"] == want.G and got["Q"] == want.Q and got["shared"] == want.counts
        assert got["remote_ctas"] == want.remote_ctas
        for name in ("offsets", "roffsets"):
            assert np.array_equal(got[name], getattr(want, name)), (rank, name)
        for name in ("indices", "remote_slot", "rgroup", "rpeer", "rpos"):
            assert np.array_equal(got[name], getattr(want, name)[:got[name].size]), (rank, name)
        assert (got["nwarps"] > 0) == want.warp_layout
        if want.warp_layout:
            assert got["nwarps"] == want.nwarps and got["remote_ctas_warp"] == want.remote_ctas_warp
            for name in ("pidx", "heads", "wgroup", "rowinfo"):
                assert np.array_equal(got[name], getattr(want, name)[:got[name].size]), (rank, name)
            assert np.array_equal(got["rgroup"], np.arange(got["Q"])), "the groups shared with a peer must be the first groups"
    if kind == "random":
        assert [Rank(r, id_parts).warp_layout for r in range(world)] == [True, False, True]
