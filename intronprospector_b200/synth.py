"""ctypes binding of csrc/synth/synth.h: seeded synthetic alignment streams for tests and bench.py."""
import ctypes as C
import os

import numpy as np

from ._native import BUILD

SHORT_UNSTRANDED, SHORT_STRANDED, LONG = 0, 1, 2


class Spec(C.Structure):
    _fields_ = [
        ("n_records", C.c_uint64),
        ("seed", C.c_uint64),
        ("model_seed", C.c_uint64),
        ("kind", C.c_int32),
        ("n_contigs", C.c_int32),
        ("genome_len", C.c_uint64),
        ("n_genes", C.c_int32),
        ("threads", C.c_int32),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(os.path.join(BUILD, "libipx_synth.so"))
        L.ipx_synth_generate.argtypes = [C.POINTER(Spec)]
        L.ipx_synth_generate.restype = C.c_void_p
        L.ipx_synth_free.argtypes = [C.c_void_p]
        L.ipx_synth_free.restype = None
        L.ipx_synth_n_records.argtypes = [C.c_void_p]
        L.ipx_synth_n_records.restype = C.c_uint64
        L.ipx_synth_n_cigar.argtypes = [C.c_void_p]
        L.ipx_synth_n_cigar.restype = C.c_uint64
        L.ipx_synth_copy.argtypes = [C.c_void_p] * 8
        L.ipx_synth_copy.restype = None
        L.ipx_synth_n_contigs.argtypes = [C.c_void_p]
        L.ipx_synth_contig_name.argtypes = [C.c_void_p, C.c_int32]
        L.ipx_synth_contig_name.restype = C.c_char_p
        L.ipx_synth_contig_len.argtypes = [C.c_void_p, C.c_int32]
        L.ipx_synth_contig_len.restype = C.c_uint64
        L.ipx_synth_write_bam.argtypes = [C.c_void_p, C.c_char_p, C.c_uint64, C.c_uint64, C.c_int]
        L.ipx_synth_write_bam_full.argtypes = [C.c_void_p, C.c_char_p, C.c_uint64, C.c_uint64, C.c_int, C.c_int]
        L.ipx_synth_write_genome.argtypes = [C.c_void_p, C.c_char_p]
        _lib = L
    return _lib


class Synth:
    """A generated stream.  columns(alloc) copies it into arrays made by `alloc(shape, dtype)` (numpy by default)."""

    def __init__(self, n_records, kind=SHORT_UNSTRANDED, seed=20261018, model_seed=20261018, n_contigs=25,
                 genome_len=3_000_000_000, n_genes=20000, threads=0):
        self.spec = Spec(n_records, seed, model_seed, kind, n_contigs, genome_len, n_genes, threads)
        self.p = lib().ipx_synth_generate(C.byref(self.spec))
        if not self.p:
            raise RuntimeError("synthetic generation failed")
        self.n = int(lib().ipx_synth_n_records(self.p))
        self.n_cigar = int(lib().ipx_synth_n_cigar(self.p))
        self.names = [lib().ipx_synth_contig_name(self.p, i).decode() for i in range(lib().ipx_synth_n_contigs(self.p))]
        self.lengths = [int(lib().ipx_synth_contig_len(self.p, i)) for i in range(len(self.names))]

    def columns(self, alloc=None):
        alloc = alloc or (lambda shape, dtype: np.empty(shape, dtype=dtype))
        cols = {
            "tid": alloc(self.n, np.int32), "pos": alloc(self.n, np.int32), "flag": alloc(self.n, np.uint16),
            "nh_class": alloc(self.n, np.uint8), "xs": alloc(self.n, np.uint8),
            "cigar_off": alloc(self.n + 1, np.uint32), "cigar": alloc(max(1, self.n_cigar), np.uint32),
        }

        def ptr(a):
            return a.data_ptr() if hasattr(a, "data_ptr") else a.ctypes.data

        lib().ipx_synth_copy(self.p, *[ptr(cols[k]) for k in ("tid", "pos", "flag", "nh_class", "xs", "cigar_off", "cigar")])
        if self.n_cigar == 0:
            cols["cigar"] = cols["cigar"][:0]
        return cols

    def write_raw_bam(self, path, begin=0, end=None):
        rc = lib().ipx_synth_write_bam(self.p, path.encode(), begin, self.n if end is None else end, 1)
        if rc != 0:
            raise RuntimeError("writing %s failed" % path)

    def write_full_bam(self, path, begin=0, end=None, threads=0, level=1):
        """The records as an aligner writes them (names, bases, qualities, a STAR-style tag set), BGZF-compressed."""
        rc = lib().ipx_synth_write_bam_full(self.p, path.encode(), begin, self.n if end is None else end, threads, level)
        if rc != 0:
            raise RuntimeError("writing %s failed" % path)

    def write_genome(self, fasta_path):
        if lib().ipx_synth_write_genome(self.p, fasta_path.encode()) != 0:
            raise RuntimeError("writing %s failed" % fasta_path)

    def close(self):
        if self.p:
            lib().ipx_synth_free(self.p)
            self.p = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
