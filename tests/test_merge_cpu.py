"""Merge path on CPU: the product's intronProspectorMerge host code (TSV parsing, CLI, formatters) linked against the
CPU oracle reproduces the unmodified reference binary on synthetic per-sample tables."""
import os
import subprocess

import pytest

import helpers
import merge_synth


@pytest.mark.parametrize("extra", [[], ["--sj-filter=canon"]], ids=["all", "canon"])
def test_oracle_merge_cli_matches_reference(built, tmp_path, extra):
    paths, _ = merge_synth.write_tables(str(tmp_path), n_samples=10, n_universe=8_000)
    outs = {}
    for tag, binary in (("ref", helpers.REF_MERGE), ("ora", helpers.ORACLE_MERGE)):
        o = {s: str(tmp_path / ("%s.%s" % (tag, s))) for s in ("introns.tsv", "juncs.bed", "introns.bed")}
        proc = subprocess.run([binary, "--intron-calls=" + o["introns.tsv"], "--junction-bed=" + o["juncs.bed"],
                               "--intron-bed=" + o["introns.bed"]] + extra + paths, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
        assert proc.returncode == 0, proc.stderr
        outs[tag] = (proc.stderr, {s: open(p, "rb").read() for s, p in o.items()})
    assert outs["ref"][0] == outs["ora"][0] and b"different strand" in outs["ref"][0]
    for s in outs["ref"][1]:
        assert outs["ref"][1][s] == outs["ora"][1][s], s


def _load_rows(paths, threads, monkeypatch):
    """the host loader of the merge tool through its C ABI (ipx_io.h), with the given number of parsing threads"""
    import ctypes as C
    import hashlib
    import numpy as np
    from intronprospector_b200 import _native
    monkeypatch.setenv("IPX_PARSE_THREADS", str(threads))
    io = _native.load_io_library()
    arr = (C.c_char_p * len(paths))(*[p.encode() for p in paths])
    calls = C.c_void_p()
    rc = io.ipxio_calls_load(arr, len(paths), _native.FILTER_NULL, C.byref(calls))
    err = io.ipxio_last_error().decode() if rc else ""
    rows = io.ipxio_calls_rows(calls).contents
    n = int(rows.n_rows)

    def column(ptr, ctype, count):
        return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ctype)), shape=(count,)).tobytes() if count else b""
    digest = hashlib.sha1(b"".join([
        column(rows.chrom, C.c_int32, n), column(rows.intron_start, C.c_uint32, n), column(rows.intron_end, C.c_uint32, n),
        column(rows.anchor_start, C.c_uint32, n), column(rows.anchor_end, C.c_uint32, n), column(rows.strand, C.c_char, n),
        column(rows.sites, C.c_uint32, n), column(rows.counts, C.c_uint32, 3 * n), column(rows.confidence, C.c_float, n)])).hexdigest()
    chroms = [io.ipxio_calls_chroms(calls)[i] for i in range(io.ipxio_calls_n_chroms(calls))]
    sites = [rows.site_strings[i] for i in range(int(rows.n_site_strings))]
    io.ipxio_calls_free(calls)
    return rc, err, n, digest, chroms, sites


def test_parallel_tsv_loading_equals_file_by_file(built, tmp_path, monkeypatch):
    """Files parsed by a pool of threads: same rows, same dictionary numbering (first use, in file order)."""
    paths, _ = merge_synth.write_tables(str(tmp_path), n_samples=9, n_universe=6_000, n_contigs=7)
    # a file that introduces chromosome names and motifs the others use later, out of the usual order
    with open(paths[4], "a") as f:
        f.write("zz_late\t10\t500\t+\t1\t0\t0\t20\t30\t0\tAT/AC\tsj99999999\n")
        f.write("aa_early\t10\t500\t-\t1\t0\t0\t20\t30\t0\tGG/TT\tsj99999998\n")
    one = _load_rows(paths, 1, monkeypatch)
    assert one[0] == 0 and one[2] > 20_000
    for threads in (2, 5, 16):
        assert _load_rows(paths, threads, monkeypatch) == one


def test_parallel_tsv_loading_fails_like_file_by_file(built, tmp_path, monkeypatch):
    """A bad line: the rows before it (earlier files, and the bad file's own earlier rows) are kept, the first bad
    file's error is reported, later files -- broken or not -- do not matter."""
    paths, _ = merge_synth.write_tables(str(tmp_path), n_samples=8, n_universe=3_000, n_contigs=5)
    lines = open(paths[3]).read().split("\n")
    lines[len(lines) // 2] = lines[len(lines) // 2].replace("\t", "\tnot-a-number\t", 1).rsplit("\t", 1)[0]
    open(paths[3], "w").write("\n".join(lines))
    open(paths[6], "w").write("chrom\tintron_start\n1\t2\t3\n")
    one = _load_rows(paths, 1, monkeypatch)
    assert one[0] != 0 and "invalid integer" in one[1] and one[2] > 3 * 1000
    for threads in (3, 16):
        assert _load_rows(paths, threads, monkeypatch) == one
