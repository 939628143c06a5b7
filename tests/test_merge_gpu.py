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


def _write(path, text, gz=False):
    import gzip
    data = text.encode() if isinstance(text, str) else text
    with (gzip.open(path, "wb") if gz else open(path, "wb")) as f:
        f.write(data)


TSV_HEADER = merge_synth.HEADER
ROWS_PLAIN = ("chr1\t100\t900\t+\t5\t1\t0\t30\t40\t2.5\tGT/AG\tsj0\n"
              "chr1\t100\t900\t+\t7\t0\t2\t50\t20\t0.918296\tGT/AG\tsj1\n"
              "chr2\t5000\t9000\t-\t1\t0\t0\t8\t8\t0\tCT/AC\tsj2\n")
# rows the device leaves to the host: exponents, many fractional digits, signs, blanks, integers beyond 32 bits, long strings,
# an empty strand, a confidence strtof rounds in the last bit
ROWS_ODD = ("chr1\t100\t900\t+\t5\t1\t0\t30\t40\t1e-05\tGT/AG\tsj0\n"
            "chr1\t100\t900\t+\t5\t1\t0\t30\t40\t0.000123456789\tGT/AG\tsj0\n"
            "chr1\t100\t900\t+\t+5\t1\t0\t30\t40\t1.5\tGT/AG\tsj0\n"
            "chr1\t 100\t900\t+\t5\t1\t0\t30\t40\t1.5\tGT/AG\tsj0\n"
            "chr1\t100\t900\t+\t4294967297\t1\t0\t30\t40\t1.5\tGT/AG\tsj0\n"
            "chr3\t100\t900\t\t5\t1\t0\t30\t40\t16777217\tGTAGGTAGGT/AGAG\tsj0\n"
            "chr3\t100\t950\t.\t5\t1\t0\t200\t40\t.5\t\tsj0\n"
            "chr3\t100\t960\t+\t5\t1\t0\t2\t40\t5.\tgt/ag\tsj0\n"
            "chr3\t100\t970\t+\t0\t0\t9\t2\t40\t0.30000001192092896\tGT/AG\tsj0\n"
            "chr3\t100\t980\t+\t0\t0\t9\t2\t40\tinf\tGT/AG\tsj0\n")


@pytest.mark.parametrize("extra", [[], ["--sj-filter=canon"]], ids=["all", "canon"])
def test_tokenizer_rows_the_device_leaves_to_the_host(built, tmp_path, extra, monkeypatch):
    """the merge tool's rows are tokenised on the device; what is not in the plain form is parsed by
    the host with the reference's rules.  Plain and odd rows mixed, columns in another order with one more column, gzip,
    a last line without its newline: files and stderr of the reference binary."""
    _write(tmp_path / "a.tsv", TSV_HEADER + ROWS_PLAIN + ROWS_ODD + ROWS_PLAIN)
    _write(tmp_path / "b.tsv.gz", TSV_HEADER + ROWS_ODD + ROWS_PLAIN.rstrip("\n"), gz=True)
    cols = TSV_HEADER.rstrip("\n").split("\t")
    order = [11, 3, 0, 2, 1, 10, 9, 8, 7, 6, 5, 4]
    swapped = "\t".join(cols[k] for k in order) + "\textra\n"
    for ln in (ROWS_PLAIN + ROWS_ODD).splitlines():
        f = ln.split("\t")
        swapped += "\t".join(f[k] for k in order) + "\tx\n"
    _write(tmp_path / "c.tsv", swapped)
    paths = [str(tmp_path / n) for n in ("a.tsv", "b.tsv.gz", "c.tsv")]
    monkeypatch.setenv("IPX_TSV_DEVICE", "1")  # every file through the device (the tool starts on the host's parser while the device comes up)
    rc_r, err_r, ref = run_merge(helpers.REF_MERGE, str(tmp_path), "ref", paths, extra)
    rc_g, err_g, gpu = run_merge(helpers.GPU_MERGE, str(tmp_path), "gpu", paths, extra)
    monkeypatch.setenv("IPX_TSV_HOST", "1")
    rc_h, err_h, host = run_merge(helpers.GPU_MERGE, str(tmp_path), "host", paths, extra)
    assert rc_r == 0, err_r
    assert (rc_g, err_g) == (rc_r, err_r) and (rc_h, err_h) == (rc_r, err_r)
    for k in ref:
        assert ref[k] is not None and ref[k] == gpu[k] == host[k], "%s differs" % k


BAD = {
    "integer": "chr1\t100\t9x0\t+\t5\t1\t0\t30\t40\t2.5\tGT/AG\tsj0\n",
    "float": "chr1\t100\t900\t+\t5\t1\t0\t30\t40\t2.5.1\tGT/AG\tsj0\n",
    "columns": "chr1\t100\t900\t+\t5\t1\t0\t30\t40\t2.5\tGT/AG\n",
    "empty_line": "\n",
    "carriage_return": "chr1\t100\t900\t+\t5\t1\t0\t30\t40\t2.5\tGT/AG\tsj0\r\n",
}


@pytest.mark.parametrize("kind", sorted(BAD))
def test_tokenizer_bad_rows_fail_like_the_reference(built, tmp_path, kind, monkeypatch):
    """A malformed row in the second of three files: the same exit code and the same stderr (the conflict lines of the
    rows before it, then the reference's error text), and no output beyond what the reference leaves."""
    good = TSV_HEADER + ROWS_PLAIN * 50
    conflict = ROWS_PLAIN.replace("\t+\t", "\t-\t", 1)
    _write(tmp_path / "a.tsv", good + conflict)
    _write(tmp_path / "b.tsv", good + BAD[kind] + ROWS_PLAIN)
    _write(tmp_path / "c.tsv", good)
    paths = [str(tmp_path / n) for n in ("a.tsv", "b.tsv", "c.tsv")]
    monkeypatch.setenv("IPX_TSV_DEVICE", "1")
    rc_r, err_r, ref = run_merge(helpers.REF_MERGE, str(tmp_path), "ref", paths)
    rc_g, err_g, gpu = run_merge(helpers.GPU_MERGE, str(tmp_path), "gpu", paths)
    assert (rc_g, err_g) == (rc_r, err_r)
    if kind != "carriage_return":
        assert rc_r != 0
    assert ref == gpu


def test_tokenizer_large_tables_match_the_host_parser(built, tmp_path, monkeypatch):
    """Many rows per file (several tiles of the newline scan, every thread block shape) against the host parser."""
    paths, _ = merge_synth.write_tables(str(tmp_path), n_samples=6, n_universe=120_000, conflicts=20)
    monkeypatch.setenv("IPX_TSV_DEVICE", "1")
    rc_g, err_g, gpu = run_merge(helpers.GPU_MERGE, str(tmp_path), "gpu", paths)
    monkeypatch.setenv("IPX_TSV_HOST", "1")
    rc_h, err_h, host = run_merge(helpers.GPU_MERGE, str(tmp_path), "host", paths)
    assert rc_g == 0 and (rc_g, err_g) == (rc_h, err_h)
    assert gpu == host and len(gpu["introns.tsv"]) > 4_000_000


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
