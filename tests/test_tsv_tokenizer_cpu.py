"""The host side of the device tokeniser (csrc/host/calls_tsv.cc: load through ipx_tsv_tokenize's contract) on CPU: a
stand-in tokeniser written here in Python follows the contract of include/ipx.h -- plain rows come back as columns, every
other row flagged -- and the loader must produce the rows, dictionaries and errors of the host parser.  (The device's
own tokeniser is held to the reference binary in tests/test_merge_gpu.py.)"""
import ctypes as C
import re

import numpy as np
import pytest

import helpers  # noqa: F401
import test_merge_gpu as inputs
from intronprospector_b200 import _native


class Columns(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("n_columns", "chrom", "intron_start", "intron_end", "strand", "uniq_mapped_count",
                                         "multi_mapped_count", "unsure_mapped_count", "max_left_overhang", "max_right_overhang",
                                         "confidence", "splice_sites")]


class Rows(C.Structure):
    _fields_ = [("n_rows", C.c_uint64), ("intron_start", C.c_void_p), ("intron_end", C.c_void_p), ("anchor_start", C.c_void_p),
                ("anchor_end", C.c_void_p), ("counts", C.c_void_p), ("confidence", C.c_void_p), ("strand", C.c_void_p),
                ("sites8", C.c_void_p), ("sites_len", C.c_void_p), ("chrom_off", C.c_void_p), ("chrom_len", C.c_void_p),
                ("line_off", C.c_void_p), ("line_len", C.c_void_p), ("flags", C.c_void_p)]


TOKENIZE = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_uint64, C.POINTER(Columns), C.POINTER(C.POINTER(Rows)))
ROWS_FREE = C.CFUNCTYPE(None, C.POINTER(Rows))
UINT = re.compile(rb"^[0-9]{1,10}$")
CONF = re.compile(rb"^(?:[0-9]+\.?[0-9]*|\.[0-9]+)$")


class StandIn:
    """ipx_tsv_tokenize's contract in numpy: plain unsigned integers that fit 32 bits, decimal fractions with at most 15
    significant and 8 fractional digits, strings without control characters; anything else flags the row."""

    def __init__(self):
        self.alive = {}
        self.calls = 0
        self.flagged = 0
        self.tokenize = TOKENIZE(self._tokenize)
        self.rows_free = ROWS_FREE(self._free)

    def _tokenize(self, bytes_ptr, n_bytes, cols_ptr, out):
        self.calls += 1
        cols = cols_ptr.contents
        data = C.string_at(bytes_ptr, n_bytes) if n_bytes else b""
        lines, off = [], 0
        while off < len(data):
            nl = data.find(b"\n", off)
            end = len(data) if nl < 0 else nl
            lines.append((off, end - off))
            off = end + 1
        n = len(lines)
        a = {"intron_start": np.zeros(n, np.uint32), "intron_end": np.zeros(n, np.uint32), "anchor_start": np.zeros(n, np.uint32),
             "anchor_end": np.zeros(n, np.uint32), "counts": np.zeros((n, 3), np.uint32), "confidence": np.zeros(n, np.float32),
             "strand": np.zeros(n, np.uint8), "sites8": np.zeros(n, np.uint64), "sites_len": np.zeros(n, np.uint8),
             "chrom_off": np.zeros(n, np.uint32), "chrom_len": np.zeros(n, np.uint16), "line_off": np.zeros(n, np.uint32),
             "line_len": np.zeros(n, np.uint32), "flags": np.ones(n, np.uint8)}
        for i, (lo, ln) in enumerate(lines):
            a["line_off"][i], a["line_len"][i] = lo, ln
            fields, pos = [], lo
            for f in data[lo:lo + ln].split(b"\t"):
                fields.append((pos, f))
                pos += len(f) + 1
            if len(fields) != cols.n_columns:
                continue
            get = lambda c: fields[c][1]  # noqa: E731
            ints = [get(c) for c in (cols.intron_start, cols.intron_end, cols.max_left_overhang, cols.max_right_overhang,
                                     cols.uniq_mapped_count, cols.multi_mapped_count, cols.unsure_mapped_count)]
            conf, sites, chrom = get(cols.confidence), get(cols.splice_sites), get(cols.chrom)
            if not all(UINT.match(x) and int(x) <= 0xFFFFFFFF for x in ints) or not CONF.match(conf):
                continue
            digits = conf.replace(b".", b"")
            frac = len(conf) - conf.index(b".") - 1 if b"." in conf else 0
            if len(digits.lstrip(b"0")) > 15 or frac > 8 or len(sites) > 8 or len(chrom) > 0xFFFF:
                continue
            if any(ch < 0x20 for ch in sites + chrom):
                continue
            v = [int(x) for x in ints]
            a["intron_start"][i], a["intron_end"][i] = v[0], v[1]
            a["anchor_start"][i], a["anchor_end"][i] = (v[0] - v[2]) & 0xFFFFFFFF, (v[1] + v[3]) & 0xFFFFFFFF
            a["counts"][i] = v[4:7]
            a["confidence"][i] = np.float32(int(digits) / 10.0 ** frac)
            strand = get(cols.strand)
            a["strand"][i] = strand[0] if strand else 0
            a["sites8"][i] = int.from_bytes(sites, "little")
            a["sites_len"][i] = len(sites)
            a["chrom_off"][i], a["chrom_len"][i] = fields[cols.chrom][0], len(chrom)
            a["flags"][i] = 0
        self.flagged += int(a["flags"].sum())
        rows = Rows()
        rows.n_rows = n
        for k, v in a.items():
            setattr(rows, k, v.ctypes.data)
        self.alive[C.addressof(rows)] = (rows, a)
        out[0] = C.pointer(rows)
        return 0

    def _free(self, rows_ptr):
        self.alive.pop(C.addressof(rows_ptr.contents), None)


def load(paths, filt, tok=None):
    io = _native.load_io_library()
    io.ipxio_calls_load_tokenized.argtypes = [C.POINTER(C.c_char_p), C.c_int32, C.c_int32, TOKENIZE, ROWS_FREE, C.POINTER(C.c_void_p)]
    arr = (C.c_char_p * len(paths))(*[p.encode() for p in paths])
    calls = C.c_void_p()
    if tok is None:
        rc = io.ipxio_calls_load(arr, len(paths), filt, C.byref(calls))
    else:
        rc = io.ipxio_calls_load_tokenized(arr, len(paths), filt, tok.tokenize, tok.rows_free, C.byref(calls))
    err = io.ipxio_last_error().decode() if rc else ""
    r = io.ipxio_calls_rows(calls).contents
    n = int(r.n_rows)

    def col(ptr, ctype, width=1):
        return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ctype)), shape=(n * width,)).copy() if n else np.zeros(0)

    got = {"chrom": col(r.chrom, C.c_int32), "i0": col(r.intron_start, C.c_uint32), "i1": col(r.intron_end, C.c_uint32),
           "a0": col(r.anchor_start, C.c_uint32), "a1": col(r.anchor_end, C.c_uint32), "strand": col(r.strand, C.c_uint8),
           "sites": col(r.sites, C.c_uint32), "counts": col(r.counts, C.c_uint32, 3), "conf": col(r.confidence, C.c_float).view(np.uint32)}
    got["site_strings"] = [r.site_strings[i] for i in range(r.n_site_strings)]
    got["chroms"] = [io.ipxio_calls_chroms(calls)[i] for i in range(io.ipxio_calls_n_chroms(calls))]
    io.ipxio_calls_free(calls)
    return rc, err, got


def same(a, b):
    assert a[0] == b[0] and a[1] == b[1]
    for k in a[2]:
        x, y = a[2][k], b[2][k]
        assert (x == y) if isinstance(x, list) else np.array_equal(x, y), k


@pytest.mark.parametrize("filt", [_native.FILTER_NULL, _native.FILTER_CANON], ids=["all", "canon"])
def test_plain_and_odd_rows_through_a_tokeniser(built, tmp_path, filt):
    inputs._write(tmp_path / "a.tsv", inputs.TSV_HEADER + inputs.ROWS_PLAIN + inputs.ROWS_ODD + inputs.ROWS_PLAIN)
    inputs._write(tmp_path / "b.tsv.gz", inputs.TSV_HEADER + inputs.ROWS_ODD + inputs.ROWS_PLAIN.rstrip("\n"), gz=True)
    cols = inputs.TSV_HEADER.rstrip("\n").split("\t")
    order = [11, 3, 0, 2, 1, 10, 9, 8, 7, 6, 5, 4]
    swapped = "\t".join(cols[k] for k in order) + "\textra\n"
    for ln in (inputs.ROWS_PLAIN + inputs.ROWS_ODD).splitlines():
        f = ln.split("\t")
        swapped += "\t".join(f[k] for k in order) + "\tx\n"
    inputs._write(tmp_path / "c.tsv", swapped)
    paths = [str(tmp_path / n) for n in ("a.tsv", "b.tsv.gz", "c.tsv")]
    tok = StandIn()
    want, got = load(paths, filt), load(paths, filt, tok)
    same(got, want)
    assert want[0] == 0 and tok.calls == 3 and tok.flagged >= 3 * 8 and not tok.alive
    assert len(want[2]["chrom"]) > 20


@pytest.mark.parametrize("kind", sorted(inputs.BAD))
def test_bad_rows_fail_alike(built, tmp_path, kind):
    good = inputs.TSV_HEADER + inputs.ROWS_PLAIN * 50
    inputs._write(tmp_path / "a.tsv", good)
    inputs._write(tmp_path / "b.tsv", good + inputs.BAD[kind] + inputs.ROWS_PLAIN)
    inputs._write(tmp_path / "c.tsv", good)
    paths = [str(tmp_path / n) for n in ("a.tsv", "b.tsv", "c.tsv")]
    tok = StandIn()
    want, got = load(paths, _native.FILTER_NULL), load(paths, _native.FILTER_NULL, tok)
    same(got, want)
    assert (want[0] != 0) == (kind != "carriage_return") and not tok.alive


def test_header_without_a_column_goes_to_the_host_parser(built, tmp_path):
    header = inputs.TSV_HEADER.replace("\tconfidence", "\tconf")
    inputs._write(tmp_path / "a.tsv", header + inputs.ROWS_PLAIN)
    tok = StandIn()
    want, got = load([str(tmp_path / "a.tsv")], _native.FILTER_NULL), load([str(tmp_path / "a.tsv")], _native.FILTER_NULL, tok)
    same(got, want)
    assert want[0] != 0 and "confidence" in want[1] and tok.calls == 0
