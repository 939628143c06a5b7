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
HROW = np.dtype([("tid", "<i4"), ("i0", "<u4"), ("i1", "<u4"), ("cnt", "<u4"), ("offset", "<u4"), ("pad", "<u4")])


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
    """Speaks the ipx_partial_* protocol with real row layouts; its 'table' is a list of rows in key order."""

    def __init__(self, names, rows_j, rows_h):
        self.names = names
        self.rank_of = {t: r for r, t in enumerate(sorted(range(len(names)), key=lambda i: names[i].encode()))}
        key = lambda r: (self.rank_of[int(r["tid"])], int(r["i0"]), int(r["i1"]))
        self.j = np.array(sorted(rows_j, key=key), dtype=JROW)
        self.h = np.array(sorted(rows_h, key=key), dtype=HROW)
        self.got_j = self.got_h = None

    def partial_row_bytes(self):
        return JROW.itemsize, HROW.itemsize

    def _part(self, rows, n_parts, split_tid):
        bounds = [self.rank_of[t] for t in split_tid] + [len(self.names)]
        part = np.zeros(len(rows), np.int64)
        for i, r in enumerate(rows):
            part[i] = next(p for p in range(n_parts) if self.rank_of[int(r["tid"])] < bounds[p])
        return part

    def partial_sizes(self, n_parts, split_tid, split_start):
        pj, ph = self._part(self.j, n_parts, split_tid), self._part(self.h, n_parts, split_tid)
        return [int((pj == p).sum()) for p in range(n_parts)], [int((ph == p).sum()) for p in range(n_parts)]

    def partial_export(self, ptr_j, ptr_h):
        C.memmove(ptr_j, self.j.ctypes.data, self.j.nbytes)
        C.memmove(ptr_h, self.h.ctypes.data, self.h.nbytes)

    def partial_import(self, ptr_j, nj, ptr_h, nh):
        self.got_j = np.frombuffer(C.string_at(ptr_j, nj * JROW.itemsize), dtype=JROW).copy()
        self.got_h = np.frombuffer(C.string_at(ptr_h, nh * HROW.itemsize), dtype=HROW).copy()


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
    pick = rng.integers(0, nj, nh)
    for k in ("tid", "i0", "i1"):
        rows_h[k] = rows_j[k][pick]
    rows_h["cnt"] = rank * 7 + 1
    rows_h["offset"] = rng.integers(0, 65536, nh)
    return list(rows_j), list(rows_h)


def _worker(rank, world, port, names, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        eng = StandInEngine(names, *make_rows(rank, len(names), 99))
        moved = sharded.exchange_dist(eng, names, torch.device("cpu"))
        np.save(os.path.join(out_dir, "j%d.npy" % rank), eng.got_j)
        np.save(os.path.join(out_dir, "h%d.npy" % rank), eng.got_h)
        assert moved > 0
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
    sent_j = np.concatenate([np.array(make_rows(r, len(names), 99)[0], dtype=JROW) for r in range(world)])
    sent_h = np.concatenate([np.array(make_rows(r, len(names), 99)[1], dtype=HROW) for r in range(world)])
    got_j = [np.load(str(tmp_path / ("j%d.npy" % r))) for r in range(world)]
    got_h = [np.load(str(tmp_path / ("h%d.npy" % r))) for r in range(world)]
    # every row reached exactly the owner of its key range, nothing was lost or duplicated
    for r in range(world):
        lo = 0 if r == 0 else bounds[r - 1]
        for rows in (got_j[r], got_h[r]):
            ranks = np.array([rank_of[int(t)] for t in rows["tid"]])
            assert ((ranks >= lo) & (ranks < bounds[r])).all()
    assert sorted(np.concatenate(got_j).tobytes()[i:i + JROW.itemsize] for i in range(0, sent_j.nbytes, JROW.itemsize)) == \
        sorted(sent_j.tobytes()[i:i + JROW.itemsize] for i in range(0, sent_j.nbytes, JROW.itemsize))
    assert sorted(np.concatenate(got_h).tobytes()[i:i + HROW.itemsize] for i in range(0, sent_h.nbytes, HROW.itemsize)) == \
        sorted(sent_h.tobytes()[i:i + HROW.itemsize] for i in range(0, sent_h.nbytes, HROW.itemsize))
    # rows of one source arrive contiguously, sources in rank order (the owner folds partials in stream order)
    for r in range(world):
        src = got_j[r]["cnt"][:, 0] // 1000
        assert (np.diff(src.astype(np.int64)) >= 0).all()
