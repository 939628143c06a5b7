"""ctypes binding of include/ipx.h and include/ipx_io.h.  No computation happens here."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
BUILD = os.path.join(HERE, "csrc", "build")
LIBIPX = os.path.join(BUILD, "libipx.so")
LIBIPX_IO = os.path.join(BUILD, "libipx_io.so")
CLI_EXTRACT = os.path.join(BUILD, "intronProspector")
CLI_MERGE = os.path.join(BUILD, "intronProspectorMerge")

IPX_OK = 0
FILTER_NULL, FILTER_ALL, FILTER_CANON = 0, 1, 2
STRAND_UN, STRAND_RF, STRAND_FR = 0, 1, 2
FMT_JUNCTION_BED, FMT_INTRON_BED9, FMT_INTRON_BED6, FMT_INTRON_TSV = 0, 1, 2, 3


class IpxError(RuntimeError):
    pass


class Params(C.Structure):
    _fields_ = [
        ("min_anchor_length", C.c_uint32),
        ("max_anchor_indel_size", C.c_uint32),
        ("min_intron_length", C.c_uint32),
        ("max_intron_length", C.c_uint32),
        ("junction_filter", C.c_int32),
        ("strandness", C.c_int32),
        ("excludes", C.c_uint32),
        ("skip_missing_targets", C.c_int32),
        ("device", C.c_int32),
        ("profile", C.c_int32),
        ("stream", C.c_void_p),
        ("table_capacity", C.c_uint64),
        ("hist_capacity", C.c_uint64),
        ("deferred_counting", C.c_int32),
    ]


class Batch(C.Structure):
    _fields_ = [
        ("n_records", C.c_uint64),
        ("n_cigar", C.c_uint64),
        ("seq_base", C.c_uint64),
        ("tid", C.c_void_p),
        ("pos", C.c_void_p),
        ("flag", C.c_void_p),
        ("nh_class", C.c_void_p),
        ("xs", C.c_void_p),
        ("cigar_off", C.c_void_p),
        ("cigar", C.c_void_p),
    ]


class BatchCompact(C.Structure):
    _fields_ = [
        ("n_records", C.c_uint64),
        ("n_cigar", C.c_uint64),
        ("seq_base", C.c_uint64),
        ("pos", C.c_void_p),
        ("flag", C.c_void_p),
        ("nh_class", C.c_void_p),
        ("xs", C.c_void_p),
        ("n_cigar_ops", C.c_void_p),
        ("cigar", C.c_void_p),
        ("n_tid_runs", C.c_uint64),
        ("run_tid", C.c_void_p),
        ("run_start", C.c_void_p),
    ]


class TableStruct(C.Structure):
    _fields_ = [
        ("n_junctions", C.c_uint64),
        ("chrom", C.POINTER(C.c_int32)),
        ("intron_start", C.POINTER(C.c_uint32)),
        ("intron_end", C.POINTER(C.c_uint32)),
        ("anchor_start", C.POINTER(C.c_uint32)),
        ("anchor_end", C.POINTER(C.c_uint32)),
        ("strand", C.POINTER(C.c_char)),
        ("splice_sites", C.POINTER(C.c_char)),
        ("read_counts", C.POINTER(C.c_uint64)),
        ("ijunc", C.POINTER(C.c_uint32)),
        ("confidence", C.POINTER(C.c_float)),
        ("hist_off", C.POINTER(C.c_uint64)),
        ("hist_cnt", C.POINTER(C.c_uint32)),
        ("n_hist", C.c_uint64),
        ("first_error_seq", C.c_uint64),
        ("status", C.c_int32),
        ("error_tid", C.c_int32),
        ("error_text", C.c_char * 256),
        ("missing_targets", C.POINTER(C.c_int32)),
        ("n_missing_targets", C.c_uint64),
        ("creation_seq", C.POINTER(C.c_uint64)),
        ("site_strings", C.POINTER(C.c_char_p)),
        ("site_ids", C.POINTER(C.c_uint32)),
    ]


class Timings(C.Structure):
    _fields_ = [
        ("extract_ms", C.c_double),
        ("aggregate_ms", C.c_double),
        ("motif_ms", C.c_double),
        ("finalize_ms", C.c_double),
        ("merge_ms", C.c_double),
        ("extract_launches", C.c_uint64),
        ("aggregate_launches", C.c_uint64),
        ("motif_launches", C.c_uint64),
        ("finalize_launches", C.c_uint64),
        ("merge_launches", C.c_uint64),
        ("records", C.c_uint64),
        ("cigar_live", C.c_uint64),
        ("events", C.c_uint64),
        ("junctions", C.c_uint64),
        ("hist_runs", C.c_uint64),
        ("motif_keys", C.c_uint64),
        ("merge_rows", C.c_uint64),
        ("export_ms", C.c_double),
        ("import_ms", C.c_double),
        ("exchange_launches", C.c_uint64),
    ]


class Text(C.Structure):
    _fields_ = [("data", C.POINTER(C.c_char)), ("len", C.c_size_t)]


class MergeRows(C.Structure):
    _fields_ = [
        ("n_rows", C.c_uint64),
        ("chrom", C.c_void_p),
        ("intron_start", C.c_void_p),
        ("intron_end", C.c_void_p),
        ("anchor_start", C.c_void_p),
        ("anchor_end", C.c_void_p),
        ("strand", C.c_void_p),
        ("sites", C.c_void_p),
        ("counts", C.c_void_p),
        ("confidence", C.c_void_p),
        ("site_strings", C.POINTER(C.c_char_p)),
        ("n_site_strings", C.c_uint32),
    ]


ABI_SYMBOLS = [
    "ipx_version", "ipx_abi_version", "ipx_device_count", "ipx_create", "ipx_destroy", "ipx_last_error",
    "ipx_set_targets", "ipx_set_genome", "ipx_push_batch", "ipx_push_batch_device", "ipx_finalize_device",
    "ipx_flush", "ipx_table_free", "ipx_clear", "ipx_merge_rows_fold", "ipx_partial_row_bytes", "ipx_partial_sizes",
    "ipx_partial_export", "ipx_partial_import", "ipx_format", "ipx_format_tsv_header", "ipx_text_free",
    "ipx_get_timings", "ipx_reset_timings", "ipx_stream", "ipx_thresholds_rows", "ipx_thresholds_export",
    "ipx_thresholds_import", "ipx_count_deferred", "ipx_table_set_ijunc", "ipx_push_batch_compact",
    "ipx_mailbox_create", "ipx_mailbox_ipc_handle", "ipx_mailbox_ptr", "ipx_mailbox_connect", "ipx_partial_export_mail",
    "ipx_partial_import_mail", "ipx_exchange_settle", "ipx_sample_keys", "ipx_host_alloc", "ipx_host_free", "ipx_sync_uploads",
    "ipx_merge_rows_fold_part", "ipx_free", "ipx_tsv_tokenize", "ipx_tsv_rows_free",
]
IO_SYMBOLS = [
    "ipxio_last_error", "ipxio_open", "ipxio_close", "ipxio_n_targets", "ipxio_target_name", "ipxio_read_batch",
    "ipxio_genome_load", "ipxio_genome_free", "ipxio_genome_n_seqs", "ipxio_genome_names", "ipxio_genome_lengths",
    "ipxio_genome_seqs", "ipxio_faidx_build", "ipxio_calls_load", "ipxio_calls_load_tokenized", "ipxio_calls_free", "ipxio_calls_rows",
    "ipxio_calls_n_chroms", "ipxio_calls_chroms", "ipxio_pipeline_read",
]

_lib = None
_io = None


def _bind(lib):
    lib.ipx_version.restype = C.c_char_p
    lib.ipx_last_error.restype = C.c_char_p
    lib.ipx_last_error.argtypes = [C.c_void_p]
    lib.ipx_create.argtypes = [C.POINTER(Params), C.POINTER(C.c_void_p)]
    lib.ipx_destroy.argtypes = [C.c_void_p]
    lib.ipx_destroy.restype = None
    lib.ipx_set_targets.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_char_p)]
    lib.ipx_set_genome.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_char_p), C.POINTER(C.c_uint64), C.POINTER(C.c_char_p)]
    lib.ipx_push_batch.argtypes = [C.c_void_p, C.POINTER(Batch)]
    lib.ipx_push_batch_device.argtypes = [C.c_void_p, C.POINTER(Batch)]
    lib.ipx_finalize_device.argtypes = [C.c_void_p]
    lib.ipx_flush.argtypes = [C.c_void_p, C.POINTER(C.POINTER(TableStruct))]
    lib.ipx_table_free.argtypes = [C.POINTER(TableStruct)]
    lib.ipx_table_free.restype = None
    lib.ipx_clear.argtypes = [C.c_void_p]
    lib.ipx_merge_rows_fold.argtypes = [C.c_void_p, C.POINTER(MergeRows), C.POINTER(C.POINTER(TableStruct)), C.POINTER(Text)]
    lib.ipx_format.argtypes = [C.POINTER(TableStruct), C.POINTER(C.c_char_p), C.c_int, C.c_float, C.c_uint64, C.c_uint64, C.POINTER(Text)]
    lib.ipx_format_tsv_header.argtypes = [C.POINTER(Text)]
    lib.ipx_text_free.argtypes = [C.POINTER(Text)]
    lib.ipx_text_free.restype = None
    lib.ipx_get_timings.argtypes = [C.c_void_p, C.POINTER(Timings)]
    lib.ipx_reset_timings.argtypes = [C.c_void_p]
    lib.ipx_stream.argtypes = [C.c_void_p]
    lib.ipx_stream.restype = C.c_void_p
    if not hasattr(lib, "ipx_partial_row_bytes"):
        return  # the CPU oracle's facade (tests only) has no multi-GPU exchange
    lib.ipx_partial_row_bytes.argtypes = [C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    lib.ipx_partial_sizes.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    lib.ipx_partial_export.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.ipx_partial_import.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64]
    lib.ipx_thresholds_rows.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint32)]
    lib.ipx_thresholds_export.argtypes = [C.c_void_p, C.c_void_p]
    lib.ipx_thresholds_import.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
    lib.ipx_count_deferred.argtypes = [C.c_void_p]
    lib.ipx_table_set_ijunc.argtypes = [C.POINTER(TableStruct), C.c_void_p]
    lib.ipx_push_batch_compact.argtypes = [C.c_void_p, C.POINTER(BatchCompact)]
    lib.ipx_mailbox_create.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_uint64, C.c_uint64]
    lib.ipx_mailbox_ipc_handle.argtypes = [C.c_void_p, C.c_void_p]
    lib.ipx_mailbox_ptr.argtypes = [C.c_void_p]
    lib.ipx_mailbox_ptr.restype = C.c_void_p
    lib.ipx_mailbox_connect.argtypes = [C.c_void_p, C.c_int32, C.c_void_p]
    lib.ipx_partial_export_mail.argtypes = [C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_uint32)]
    lib.ipx_partial_import_mail.argtypes = [C.c_void_p]
    lib.ipx_exchange_settle.argtypes = [C.c_void_p, C.POINTER(C.c_int32)]
    lib.ipx_sample_keys.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64)]


def load_library(path=None):
    """Load the CUDA library.  Raises when it has not been built: there is no fallback implementation."""
    global _lib
    if _lib is None or path is not None:
        p = path or LIBIPX
        if not os.path.exists(p):
            raise IpxError("%s not found: build it with `make -C intronprospector_b200/csrc` (needs nvcc, sm_100a)" % p)
        lib = C.CDLL(p)
        _bind(lib)
        if path is not None:
            return lib
        _lib = lib
    return _lib


def load_io_library():
    global _io
    if _io is None:
        if not os.path.exists(LIBIPX_IO):
            raise IpxError("%s not found: build it with `make -C intronprospector_b200/csrc`" % LIBIPX_IO)
        io = C.CDLL(LIBIPX_IO)
        io.ipxio_last_error.restype = C.c_char_p
        io.ipxio_open.argtypes = [C.c_char_p, C.POINTER(C.c_void_p)]
        io.ipxio_close.argtypes = [C.c_void_p]
        io.ipxio_close.restype = None
        io.ipxio_n_targets.argtypes = [C.c_void_p]
        io.ipxio_target_name.argtypes = [C.c_void_p, C.c_int32]
        io.ipxio_target_name.restype = C.c_char_p
        io.ipxio_read_batch.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64] + [C.c_void_p] * 7 + [C.POINTER(C.c_uint64)]
        io.ipxio_read_batch.restype = C.c_int64
        io.ipxio_pipeline_read.argtypes = [C.c_char_p, C.c_int32, C.c_uint64, C.c_uint64] + [C.c_void_p] * 7 + [C.POINTER(C.c_uint64)]
        io.ipxio_pipeline_read.restype = C.c_int64
        io.ipxio_genome_load.argtypes = [C.c_char_p, C.POINTER(C.c_void_p)]
        io.ipxio_genome_free.argtypes = [C.c_void_p]
        io.ipxio_genome_free.restype = None
        io.ipxio_genome_n_seqs.argtypes = [C.c_void_p]
        io.ipxio_genome_names.argtypes = [C.c_void_p]
        io.ipxio_genome_names.restype = C.POINTER(C.c_char_p)
        io.ipxio_genome_lengths.argtypes = [C.c_void_p]
        io.ipxio_genome_lengths.restype = C.POINTER(C.c_uint64)
        io.ipxio_genome_seqs.argtypes = [C.c_void_p]
        io.ipxio_genome_seqs.restype = C.POINTER(C.c_char_p)
        io.ipxio_faidx_build.argtypes = [C.c_char_p]
        io.ipxio_calls_load.argtypes = [C.POINTER(C.c_char_p), C.c_int32, C.c_int32, C.POINTER(C.c_void_p)]
        io.ipxio_calls_free.argtypes = [C.c_void_p]
        io.ipxio_calls_free.restype = None
        io.ipxio_calls_rows.argtypes = [C.c_void_p]
        io.ipxio_calls_rows.restype = C.POINTER(MergeRows)
        io.ipxio_calls_n_chroms.argtypes = [C.c_void_p]
        io.ipxio_calls_chroms.argtypes = [C.c_void_p]
        io.ipxio_calls_chroms.restype = C.POINTER(C.c_char_p)
        _io = io
    return _io


def _arr(ptr, n, dtype):
    if n == 0:
        return np.zeros(0, dtype=dtype)
    return np.ctypeslib.as_array(ptr, shape=(n,)).view(dtype).copy()


def _view(ptr, n, dtype):
    """numpy view of a library-owned column (no copy): valid while the native table is alive"""
    if n == 0:
        return np.zeros(0, dtype=dtype)
    a = np.ctypeslib.as_array(ptr, shape=(n,)).view(dtype)
    a.flags.writeable = False
    return a


class Table:
    """An ipx_table.  The columns are numpy views of the library's own buffers, made on first use (so ipx_flush costs
    what the C call costs); they stay valid until close(), which frees the native table.  Copy what must outlive it."""

    _COLUMNS = {
        "chrom": ("chrom", np.int32, 1), "intron_start": ("intron_start", np.uint32, 1), "intron_end": ("intron_end", np.uint32, 1),
        "anchor_start": ("anchor_start", np.uint32, 1), "anchor_end": ("anchor_end", np.uint32, 1),
        "strand": ("strand", np.uint8, 1), "splice_sites": ("splice_sites", np.uint8, 8), "read_counts": ("read_counts", np.uint64, 3),
        "ijunc": ("ijunc", np.uint32, 1), "confidence": ("confidence", np.float32, 1),
    }

    def __init__(self, lib, tp, names, free):
        t = tp.contents
        self.n = int(t.n_junctions)
        self.n_hist = int(t.n_hist)
        self.names = list(names)
        self.status = int(t.status)
        self.first_error_seq = int(t.first_error_seq)
        self.error_tid = int(t.error_tid)
        self.error_text = t.error_text.decode(errors="replace")
        self.missing_targets = _arr(t.missing_targets, int(t.n_missing_targets), np.int32)
        self._lib, self._tp, self._free = lib, tp, free
        self._cnames = (C.c_char_p * max(1, len(names)))(*[s.encode() for s in names])
        self._text = {}
        self._cols = {}

    def __getattr__(self, name):
        # columns on demand; only called for attributes not set in __init__
        if name.startswith("_") or self.__dict__.get("_tp") is None:
            raise AttributeError(name)
        cols = self.__dict__["_cols"]
        if name in cols:
            return cols[name]
        t = self._tp.contents
        n = self.n
        if name in Table._COLUMNS:
            field, dtype, width = Table._COLUMNS[name]
            ptr = getattr(t, field)
            if dtype == np.uint8:
                ptr = C.cast(ptr, C.POINTER(C.c_uint8))
            a = _view(ptr, n * width, dtype)
            a = a.reshape(n, width) if width > 1 else a
        elif name == "hist_off":
            a = _view(t.hist_off, n + 1, np.uint64)
        elif name == "hist_cnt":
            a = _view(t.hist_cnt, self.n_hist, np.uint32)
        elif name == "creation_seq":
            a = _view(t.creation_seq, n, np.uint64) if t.creation_seq else np.zeros(0, np.uint64)
        else:
            raise AttributeError(name)
        cols[name] = a
        return a

    def text(self, fmt, min_confidence=0.0, begin=0, end=None):
        """Formatted rows [begin, end) of the table (all of it by default), sorted as the reference sorts them."""
        end = self.n if end is None else end
        key = (fmt, float(min_confidence), begin, end)
        if key not in self._text:
            if self._tp is None:
                raise IpxError("table already closed")
            txt = Text()
            rc = self._lib.ipx_format(self._tp, self._cnames, fmt, key[1], begin, end, C.byref(txt))
            if rc != IPX_OK:
                raise IpxError("ipx_format failed: %d" % rc)
            self._text[key] = C.string_at(txt.data, txt.len) if txt.len else b""
            self._lib.ipx_text_free(C.byref(txt))
        return self._text[key]

    def text_by_target(self, fmt, min_confidence=0.0):
        """The text of a coordinate-sorted single-file run: the rows of each target in header order, each target's rows
        sorted on their own (extract_junctions_by_target, intronProspector.cc:356-371)."""
        if self.n == 0:
            return b""
        chrom = np.asarray(self.chrom)
        cuts = np.flatnonzero(np.concatenate(([True], chrom[1:] != chrom[:-1], [True])))
        ranges = sorted((int(chrom[b]), int(b), int(e)) for b, e in zip(cuts[:-1], cuts[1:]))
        return b"".join(self.text(fmt, min_confidence, b, e) for _, b, e in ranges)

    def tsv_header(self):
        txt = Text()
        if self._lib.ipx_format_tsv_header(C.byref(txt)) != IPX_OK:
            raise IpxError("ipx_format_tsv_header failed")
        out = C.string_at(txt.data, txt.len)
        self._lib.ipx_text_free(C.byref(txt))
        return out

    def set_ijunc(self, ijunc):
        """Renumber the junctions (multi-GPU: global ranks of creation_seq)."""
        a = np.ascontiguousarray(ijunc, dtype=np.uint32)
        if a.shape[0] != self.n or self._tp is None:
            raise IpxError("set_ijunc: wrong length or closed table")
        if self._lib.ipx_table_set_ijunc(self._tp, a.ctypes.data) != IPX_OK:
            raise IpxError("ipx_table_set_ijunc failed")
        self._cols.pop("ijunc", None)  # the library may have moved the column
        self._text = {}

    def close(self):
        if self.__dict__.get("_tp") is not None:
            self._cols.clear()
            self._free(self._tp)
            self._tp = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def columns(self):
        """Detached copies of every column (safe to keep after close())."""
        return {k: np.array(getattr(self, k)) for k in ("chrom", "intron_start", "intron_end", "anchor_start", "anchor_end", "strand",
                                                        "splice_sites", "read_counts", "ijunc", "confidence", "hist_off", "hist_cnt")}


class Engine:
    """One ipx_handle.  `lib` lets tests drive the CPU oracle's ABI facade through the same code."""

    def __init__(self, lib=None, device=0, stream=None, profile=False, **kw):
        self.lib = lib or load_library()
        p = Params()
        p.min_anchor_length = kw.pop("min_anchor_length", 8)
        p.max_anchor_indel_size = kw.pop("max_anchor_indel_size", 0)
        p.min_intron_length = kw.pop("min_intron_length", 70)
        p.max_intron_length = kw.pop("max_intron_length", 500000)
        p.junction_filter = kw.pop("junction_filter", FILTER_NULL)
        p.strandness = kw.pop("strandness", STRAND_UN)
        p.excludes = kw.pop("excludes", 0)
        p.skip_missing_targets = int(kw.pop("skip_missing_targets", False))
        p.device = device
        p.profile = int(profile)
        p.stream = stream
        p.table_capacity = kw.pop("table_capacity", 0)
        p.hist_capacity = kw.pop("hist_capacity", 0)
        p.deferred_counting = int(kw.pop("deferred_counting", False))
        if kw:
            raise TypeError("unknown parameters: %s" % sorted(kw))
        self.params = p
        self.h = C.c_void_p()
        rc = self.lib.ipx_create(C.byref(p), C.byref(self.h))
        if rc != IPX_OK:
            raise IpxError("ipx_create failed (%d): no usable CUDA device; there is no CPU fallback" % rc)
        self.names = []
        self._keep = []

    def close(self):
        if self.h:
            self.lib.ipx_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != IPX_OK:
            raise IpxError("%s (code %d)" % (self.lib.ipx_last_error(self.h).decode(errors="replace"), rc))

    def set_targets(self, names):
        self.names = list(names)
        arr = (C.c_char_p * max(1, len(names)))(*[s.encode() for s in names])
        self._check(self.lib.ipx_set_targets(self.h, len(names), arr))

    def set_genome(self, names, seqs):
        n = len(names)
        cn = (C.c_char_p * max(1, n))(*[s.encode() for s in names])
        bufs = [s if isinstance(s, bytes) else s.encode() for s in seqs]
        cs = (C.c_char_p * max(1, n))(*bufs)
        lens = (C.c_uint64 * max(1, n))(*[len(b) for b in bufs])
        self._check(self.lib.ipx_set_genome(self.h, n, cn, lens, cs))

    @staticmethod
    def _ptr(a):
        if hasattr(a, "data_ptr"):
            return a.data_ptr()
        return a.ctypes.data

    def make_batch(self, cols, seq_base=0):
        b = Batch()
        b.n_records = int(cols["tid"].shape[0])
        b.n_cigar = int(cols["cigar"].shape[0])
        b.seq_base = seq_base
        for k in ("tid", "pos", "flag", "nh_class", "xs", "cigar_off", "cigar"):
            setattr(b, k, self._ptr(cols[k]))
        return b

    def push(self, cols, seq_base=0):
        """cols: dict of numpy arrays (host) with the ipx_batch columns."""
        b = self.make_batch(cols, seq_base)
        self._keep = [cols]
        self._check(self.lib.ipx_push_batch(self.h, C.byref(b)))

    @staticmethod
    def compact_columns(cols, alloc=None):
        """ipx_batch columns -> ipx_batch_compact columns (host side; what a record packer would produce directly):
        operation counts instead of offsets, target runs instead of a target column."""
        off = np.asarray(cols["cigar_off"]).astype(np.int64)
        n_ops = np.diff(off)
        if n_ops.size and n_ops.max() > 0xFFFF:
            raise IpxError("a record has more than 65535 CIGAR operations: use the plain batch")
        tid = np.asarray(cols["tid"])
        starts = np.flatnonzero(np.concatenate(([True], tid[1:] != tid[:-1]))) if tid.size else np.zeros(0, np.int64)
        out = {k: cols[k] for k in ("pos", "flag", "nh_class", "xs", "cigar")}
        ops = n_ops.astype(np.uint16)
        if alloc is not None:
            buf = alloc(ops.shape, np.uint16)
            (buf.numpy() if hasattr(buf, "numpy") else buf)[...] = ops
            ops = buf
        out["n_cigar_ops"] = ops
        out["run_start"] = np.ascontiguousarray(starts.astype(np.uint64))
        out["run_tid"] = np.ascontiguousarray(tid[starts].astype(np.int32)) if tid.size else np.zeros(0, np.int32)
        return out

    def push_compact(self, ccols, seq_base=0):
        """ccols: what compact_columns returns (host arrays / pinned tensors)."""
        b = BatchCompact()
        b.n_records = int(ccols["pos"].shape[0])
        b.n_cigar = int(ccols["cigar"].shape[0])
        b.seq_base = seq_base
        for k in ("pos", "flag", "nh_class", "xs", "n_cigar_ops", "cigar", "run_tid", "run_start"):
            setattr(b, k, self._ptr(ccols[k]))
        b.n_tid_runs = int(ccols["run_tid"].shape[0])
        self._keep = [ccols]
        self._check(self.lib.ipx_push_batch_compact(self.h, C.byref(b)))

    def push_device(self, cols, seq_base=0):
        """cols: dict of torch CUDA tensors."""
        b = self.make_batch(cols, seq_base)
        self._check(self.lib.ipx_push_batch_device(self.h, C.byref(b)))

    def finalize_device(self):
        self._check(self.lib.ipx_finalize_device(self.h))

    def flush(self):
        tp = C.POINTER(TableStruct)()
        self._check(self.lib.ipx_flush(self.h, C.byref(tp)))
        return Table(self.lib, tp, self.names, self.lib.ipx_table_free)

    def clear(self):
        self._check(self.lib.ipx_clear(self.h))

    def merge_rows(self, rows_ptr):
        tp = C.POINTER(TableStruct)()
        log = Text()
        rc = self.lib.ipx_merge_rows_fold(self.h, rows_ptr, C.byref(tp), C.byref(log))
        text = C.string_at(log.data, log.len) if log.len else b""
        self.lib.ipx_text_free(C.byref(log))
        self._check(rc)
        return Table(self.lib, tp, self.names, self.lib.ipx_table_free), text

    # ---- multi-GPU exchange (see sharded.py) ----
    def partial_row_bytes(self):
        jb, hb = C.c_uint32(), C.c_uint32()
        self._check(self.lib.ipx_partial_row_bytes(C.byref(jb), C.byref(hb)))
        return jb.value, hb.value

    def partial_sizes(self, n_parts, split_tid, split_start):
        st = (C.c_int32 * max(1, n_parts - 1))(*split_tid)
        ss = (C.c_uint32 * max(1, n_parts - 1))(*split_start)
        jr, hr = (C.c_uint64 * n_parts)(), (C.c_uint64 * n_parts)()
        self._check(self.lib.ipx_partial_sizes(self.h, n_parts, st, ss, jr, hr))
        return [int(x) for x in jr], [int(x) for x in hr]

    def partial_export(self, dev_junction_rows, dev_hist_rows):
        self._check(self.lib.ipx_partial_export(self.h, dev_junction_rows, dev_hist_rows))

    def partial_import(self, dev_junction_rows, n_junction_rows, dev_hist_rows, n_hist_rows):
        self._check(self.lib.ipx_partial_import(self.h, dev_junction_rows, n_junction_rows, dev_hist_rows, n_hist_rows))

    # exchange over peer memory (ipx.h: ipx_mailbox_*, ipx_partial_*_mail)
    def mailbox_create(self, world, rank, junction_rows_per_source, hist_rows_per_source):
        self._check(self.lib.ipx_mailbox_create(self.h, world, rank, junction_rows_per_source, hist_rows_per_source))
        self.mailbox_world = world

    def mailbox_ipc_handle(self):
        buf = C.create_string_buffer(64)
        self._check(self.lib.ipx_mailbox_ipc_handle(self.h, buf))
        return buf.raw

    def mailbox_ptr(self):
        return int(self.lib.ipx_mailbox_ptr(self.h) or 0)

    def mailbox_connect_ipc(self, handles):
        """handles: the 64-byte IPC handles of all ranks in rank order (this rank's own entry is ignored)"""
        blob = b"".join(handles)
        assert len(blob) == 64 * self.mailbox_world
        self._check(self.lib.ipx_mailbox_connect(self.h, 1, C.c_char_p(blob)))
        self.mailbox_connected = True

    def mailbox_connect_ptrs(self, ptrs):
        """ptrs: the mailbox device pointers of all ranks in rank order (handles of the same process)"""
        arr = (C.c_void_p * self.mailbox_world)(*ptrs)
        self._check(self.lib.ipx_mailbox_connect(self.h, 0, arr))
        self.mailbox_connected = True

    def partial_export_mail(self, split_tid, split_start):
        n = max(1, len(split_tid))
        st = (C.c_int32 * n)(*split_tid)
        ss = (C.c_uint32 * n)(*split_start)
        self._check(self.lib.ipx_partial_export_mail(self.h, st, ss))

    def partial_import_mail(self):
        self._check(self.lib.ipx_partial_import_mail(self.h))

    def exchange_settle(self):
        """Wait for this rank's share of the exchange; True when events were folded after its export (a table had been
        full): every rank then repeats the exchange before flushing."""
        r = C.c_int32()
        self._check(self.lib.ipx_exchange_settle(self.h, C.byref(r)))
        return bool(r.value)

    def sample_keys(self, stride=1, max_keys=1 << 16):
        """Every stride-th junction key of the local table as (target index, intron start) arrays."""
        tid = np.empty(max_keys, np.int32)
        i0 = np.empty(max_keys, np.uint32)
        n = C.c_uint64()
        self._check(self.lib.ipx_sample_keys(self.h, stride, max_keys, tid.ctypes.data, i0.ctypes.data, C.byref(n)))
        return tid[:n.value], i0[:n.value]

    def thresholds_rows(self):
        n, b = C.c_uint64(), C.c_uint32()
        self._check(self.lib.ipx_thresholds_rows(self.h, C.byref(n), C.byref(b)))
        return int(n.value), int(b.value)

    def thresholds_export(self, dev_rows):
        self._check(self.lib.ipx_thresholds_export(self.h, dev_rows))

    def thresholds_import(self, dev_rows, n_rows):
        self._check(self.lib.ipx_thresholds_import(self.h, dev_rows, n_rows))

    def count_deferred(self):
        self._check(self.lib.ipx_count_deferred(self.h))

    def timings(self):
        t = Timings()
        self._check(self.lib.ipx_get_timings(self.h, C.byref(t)))
        return {k: getattr(t, k) for k, _ in Timings._fields_}

    def reset_timings(self):
        self._check(self.lib.ipx_reset_timings(self.h))

    def stream(self):
        return self.lib.ipx_stream(self.h)
