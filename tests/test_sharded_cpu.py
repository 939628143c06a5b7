"""Host-side logic of the multi-GPU path on CPU: key splitters, and the byte routing of the key-range all-to-all run
over gloo with world_size 2 against a stand-in for the engine (rows with the real wire layout, no arithmetic)."""
import ctypes as C
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import helpers  # noqa: F401
from intronprospector_b200 import sharded

JROW = np.dtype([("first", "<u8", 3), ("cnt", "<u8", 3), ("tid", "<i4"), ("i0", "<u4"), ("i1", "<u4"), ("a0", "<u4"),
                 ("a1", "<u4"), ("pad", "<u4")])
HROW = np.dtype([("jrow", "<u4"), ("offset", "<u2"), ("cnt", "<u2")])


def test_wire_layout_matches_the_library(built):
    from intronprospector_b200 import _native
    lib = _native.load_library()
    jb, hb = C.c_uint32(), C.c_uint32()
    assert lib.ipx_partial_row_bytes(C.byref(jb), C.byref(hb)) == 0
    assert (jb.value, hb.value) == (JROW.itemsize, HROW.itemsize)


def test_key_splitters_follow_name_order():
    names = ["chr1", "chr2", "chr10", "chrX", "chrM", "chr3"]
    order = sorted(range(len(names)), key=lambda i: names[i].encode())
    for parts in (1, 2, 3, 4, 8):
        tid, start = sharded.key_splitters(names, parts)
        assert len(tid) == len(start) == parts - 1
        ranks = [order.index(t) for t in tid]
        assert ranks == sorted(ranks)
        assert all(s == 0 for s in start)
    assert sharded.key_splitters(names, 2)[0] == [order[3]]


def test_global_numbering():
    a = np.array([50, 10, 70], np.uint64)
    b = np.array([20, 60], np.uint64)
    na, nb = sharded.global_numbering([a, b])
    assert list(na) == [2, 0, 4] and list(nb) == [1, 3]


class StandInEngine:
    """Speaks the ipx_partial_* protocol with the real row layouts; its 'table' is a list of junction rows in key order
    and offset rows that name their junction by position."""

    def __init__(self, names, rows_j, rows_h):
        self.names = names
        self.rank_of = {t: r for r, t in enumerate(sorted(range(len(names)), key=lambda i: names[i].encode()))}
        order = sorted(range(len(rows_j)), key=lambda i: (self.rank_of[int(rows_j[i]["tid"])], int(rows_j[i]["i0"]), int(rows_j[i]["i1"])))
        new_index = {old: new for new, old in enumerate(order)}
        self.j = np.array([rows_j[i] for i in order], dtype=JROW)
        h = np.array(rows_h, dtype=HROW)
        h["jrow"] = [new_index[int(x)] for x in h["jrow"]]
        self.h = h[np.argsort(h["jrow"], kind="stable")]
        self.got = []

    def partial_row_bytes(self):
        return JROW.itemsize, HROW.itemsize

    def _part_of_junctions(self, n_parts, split_tid):
        bounds = [self.rank_of[t] for t in split_tid] + [len(self.names)]
        return np.array([next(p for p in range(n_parts) if self.rank_of[int(r["tid"])] < bounds[p]) for r in self.j], np.int64)

    def partial_sizes(self, n_parts, split_tid, split_start):
        self.pj = self._part_of_junctions(n_parts, split_tid)
        ph = self.pj[self.h["jrow"]]
        return [int((self.pj == p).sum()) for p in range(n_parts)], [int((ph == p).sum()) for p in range(n_parts)]

    def partial_export(self, ptr_j, ptr_h):
        # junction rows are in key order, so each part is a contiguous run; offset rows follow their junctions
        first = {int(p): int(np.flatnonzero(self.pj == p)[0]) for p in np.unique(self.pj)}
        h = self.h.copy()
        h["jrow"] = [int(x) - first[int(self.pj[int(x)])] for x in self.h["jrow"]]
        C.memmove(ptr_j, self.j.ctypes.data, self.j.nbytes)
        C.memmove(ptr_h, h.ctypes.data, h.nbytes)

    def partial_import(self, ptr_j, nj, ptr_h, nh):
        j = np.frombuffer(C.string_at(ptr_j, nj * JROW.itemsize), dtype=JROW).copy()
        h = np.frombuffer(C.string_at(ptr_h, nh * HROW.itemsize), dtype=HROW).copy()
        assert (h["jrow"] < max(1, nj)).all()
        self.got.append((j, h))


def make_rows(rank, n_targets, seed):
    rng = np.random.default_rng(seed + rank)
    nj = 400 + 37 * rank
    rows_j = np.zeros(nj, JROW)
    rows_j["tid"] = rng.integers(0, n_targets, nj)
    rows_j["i0"] = rng.integers(0, 1 << 28, nj)
    rows_j["i1"] = rows_j["i0"] + rng.integers(70, 100000, nj)
    rows_j["cnt"][:, 0] = rank * 1000 + np.arange(nj)
    nh = 900 + 11 * rank
    rows_h = np.zeros(nh, HROW)
    rows_h["jrow"] = rng.integers(0, nj, nh)
    rows_h["cnt"] = rank * 7 + 1
    rows_h["offset"] = rng.integers(0, 65536, nh)
    return list(rows_j), list(rows_h)


def resolved(rows_j, rows_h):
    """offset rows with their junction's key instead of its index: what must survive the exchange"""
    return sorted((int(rows_j[int(h["jrow"])]["tid"]), int(rows_j[int(h["jrow"])]["i0"]), int(rows_j[int(h["jrow"])]["i1"]),
                   int(h["offset"]), int(h["cnt"])) for h in rows_h)


def _worker(rank, world, port, names, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        eng = StandInEngine(names, *make_rows(rank, len(names), 99))
        moved = sharded.exchange_dist(eng, names, torch.device("cpu"))
        for src, (j, h) in enumerate(eng.got):
            np.save(os.path.join(out_dir, "j%d_%d.npy" % (rank, src)), j)
            np.save(os.path.join(out_dir, "h%d_%d.npy" % (rank, src)), h)
        assert moved > 0 and len(eng.got) == world
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_key_range_exchange_over_gloo(tmp_path):
    world = 2
    names = ["chr%d" % i for i in range(1, 12)] + ["chrX"]
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(world, port, names, str(tmp_path)), nprocs=world, join=True)
    rank_of = {t: r for r, t in enumerate(sorted(range(len(names)), key=lambda i: names[i].encode()))}
    split_tid, _ = sharded.key_splitters(names, world)
    bounds = [rank_of[t] for t in split_tid] + [len(names)]
    sent_j, sent_h = [], []
    for r in range(world):
        rj, rh = make_rows(r, len(names), 99)
        sent_j += [bytes(np.array(x, dtype=JROW).tobytes()) for x in rj]
        sent_h += resolved(rj, rh)
    got_j, got_h = [], []
    for dst in range(world):
        lo = 0 if dst == 0 else bounds[dst - 1]
        for src in range(world):
            j = np.load(str(tmp_path / ("j%d_%d.npy" % (dst, src))))
            h = np.load(str(tmp_path / ("h%d_%d.npy" % (dst, src))))
            # every junction row reached the owner of its key range, from the right source
            ranks = np.array([rank_of[int(t)] for t in j["tid"]])
            assert ((ranks >= lo) & (ranks < bounds[dst])).all()
            assert ((j["cnt"][:, 0] // 1000) == src).all()
            got_j += [bytes(x.tobytes()) for x in j]
            got_h += resolved(list(j), list(h))
    # nothing was lost, duplicated or attached to another junction
    assert sorted(got_j) == sorted(sent_j)
    assert sorted(got_h) == sorted(sent_h)
