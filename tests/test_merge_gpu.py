"""BASELINE config 5 on the GPU: intronProspectorMerge over many synthetic per-sample tables, byte-identical to the
unmodified reference binary; and the key-range sharded fold (one handle per chromosome range) equal to the whole."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import helpers
import merge_synth
from intronprospector_b200 import _native, sharded

pytestmark = pytest.mark.gpu


def run_merge(binary, workdir, tag, inputs, extra=()):
    out = {s: os.path.join(workdir, "%s.%s" % (tag, s)) for s in ("introns.tsv", "juncs.bed", "introns.bed")}
    cmd = [binary, "--intron-calls=" + out["introns.tsv"], "--junction-bed=" + out["juncs.bed"],
           "--intron-bed=" + out["introns.bed"]] + list(extra) + list(inputs)
    proc = subprocess.run(cmd, cwd=workdir, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    return proc.returncode, proc.stderr, {s: open(p, "rb").read() if os.path.exists(p) else None for s, p in out.items()}


@pytest.mark.parametrize("extra", [[], ["--sj-filter=canon"]], ids=["all", "canon"])
def test_merge_cli_matches_reference(built, tmp_path, extra):
    paths, _ = merge_synth.write_tables(str(tmp_path), n_samples=24, n_universe=30_000)
    rc_r, err_r, ref = run_merge(helpers.REF_MERGE, str(tmp_path), "ref", paths, extra)
    rc_g, err_g, gpu = run_merge(helpers.GPU_MERGE, str(tmp_path), "gpu", paths, extra)
    assert rc_r == 0 and rc_g == 0, (err_r, err_g)
    assert err_r == err_g  # strand-conflict lines, in file/row order
    assert b"different strand" in err_r
    for k in ref:
        assert ref[k] is not None and ref[k] == gpu[k], "%s differs" % k


@pytest.mark.parametrize("n_parts", [2, 5])
def test_merge_cli_sharded_matches_reference(built, tmp_path, n_parts, monkeypatch):
    """BASELINE config 5 through the tool: the fold sharded over key ranges (one handle and one thread per range; on a
    multi-GPU host one device per range) against the reference binary -- files, and the strand-conflict lines on
    stderr in input order although they are raised on different ranges."""
    paths, _ = merge_synth.write_tables(str(tmp_path), n_samples=12, n_universe=30_000, conflicts=40)
    rc_r, err_r, ref = run_merge(helpers.REF_MERGE, str(tmp_path), "ref", paths)
    monkeypatch.setenv("IPX_MERGE_PARTS", str(n_parts))
    rc_g, err_g, gpu = run_merge(helpers.GPU_MERGE, str(tmp_path), "gpu", paths)
    assert rc_r == 0 and rc_g == 0, (err_r, err_g)
    assert err_r == err_g and err_r.count(b"different strand") > 40
    for k in ref:
        assert ref[k] is not None and ref[k] == gpu[k], "%s differs" % k


def test_merge_sharded_by_key_range(built, tmp_path):
    """Rows partitioned by chromosome range (no exchange needed: a key lives on one rank), one fold per rank; sjN is
    the index in anchor order, so a rank's numbers are its local ones plus the junction count of the ranks before."""
    paths, _ = merge_synth.write_tables(str(tmp_path), n_samples=8, n_universe=20_000, seed=9)
    io = _native.load_io_library()
    arr = (C.c_char_p * len(paths))(*[p.encode() for p in paths])
    calls = C.c_void_p()
    assert io.ipxio_calls_load(arr, len(paths), _native.FILTER_NULL, C.byref(calls)) == 0
    rows = io.ipxio_calls_rows(calls)
    names = [io.ipxio_calls_chroms(calls)[i].decode() for i in range(io.ipxio_calls_n_chroms(calls))]

    def fold(rows_ptr):
        eng = _native.Engine(device=0)
        eng.set_targets(names)
        table, log = eng.merge_rows(rows_ptr)
        eng.close()
        return table, log

    whole, whole_log = fold(rows)
    world = 3
    parts = [sharded.merge_rows_for_part(rows.contents, names, world, p) for p in range(world)]
    tables, logs, base = [], [], 0
    for part in parts:
        t, log = fold(C.pointer(part.struct))
        t.set_ijunc(t.ijunc + np.uint32(base))
        base += t.n
        tables.append(t)
        logs.append(log)
    assert sum(t.n for t in tables) == whole.n > 10_000
    for fmt in (_native.FMT_INTRON_TSV, _native.FMT_INTRON_BED9, _native.FMT_JUNCTION_BED):
        assert b"".join(t.text(fmt, 0.0) for t in tables) == whole.text(fmt, 0.0)
    assert sorted(b"".join(logs).split(b"\n")) == sorted(whole_log.split(b"\n"))
    io.ipxio_calls_free(calls)
