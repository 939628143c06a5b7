"""smoke(): one small extraction through the C ABI on cuda:0, checked against the CPU oracle (test-only code
under oracle/; loaded here purely as the checker)."""
import os

import numpy as np

from . import _native
from .synth import SHORT_UNSTRANDED, Synth

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_ABI = os.path.join(REPO, "oracle", "_ref", "libipx_oracle_abi.so")

COLUMNS = ("chrom", "intron_start", "intron_end", "anchor_start", "anchor_end", "strand", "splice_sites",
           "read_counts", "ijunc", "confidence", "hist_off", "hist_cnt")


def oracle_engine(**kw):
    """An Engine bound to the CPU oracle's ABI facade (checker only)."""
    lib = _native.load_library(ORACLE_ABI)
    return _native.Engine(lib=lib, **kw)


def compare_tables(a, b):
    """Returns a list of differences between two Table objects (empty = identical)."""
    diffs = []
    if a.n != b.n:
        diffs.append("junction count %d != %d" % (a.n, b.n))
        return diffs
    ca, cb = a.columns(), b.columns()
    for k in COLUMNS:
        if ca[k].shape != cb[k].shape or not np.array_equal(ca[k], cb[k]):
            bad = "shape" if ca[k].shape != cb[k].shape else int(np.argmax(np.any((ca[k] != cb[k]).reshape(len(ca[k]), -1), axis=1)))
            diffs.append("column %s differs (first at %s)" % (k, bad))
    # the stream position of each junction's creating event (what sjN ranks, what multi-GPU runs agree on)
    if len(a.creation_seq) == a.n and len(b.creation_seq) == b.n and not np.array_equal(a.creation_seq, b.creation_seq):
        diffs.append("column creation_seq differs")
    for fmt in (_native.FMT_INTRON_BED9, _native.FMT_INTRON_BED6, _native.FMT_INTRON_TSV):
        for minc in (0.0, 1.0):
            if a.text(fmt, minc) != b.text(fmt, minc):
                diffs.append("text format %d (min confidence %g) differs" % (fmt, minc))
    # junction BED: the reference's std::sort leaves ties between equal-anchor rows unspecified; the oracle keeps key
    # order, so compare as multisets here (byte order is pinned against the reference binary in the CLI tests)
    for minc in (0.0, 1.0):
        if sorted(a.text(_native.FMT_JUNCTION_BED, minc).split(b"\n")) != sorted(b.text(_native.FMT_JUNCTION_BED, minc).split(b"\n")):
            diffs.append("junction BED (min confidence %g) differs" % minc)
    return diffs


def run_engine(eng, names, cols, genome=None, n_batches=1):
    eng.set_targets(names)
    if genome is not None:
        eng.set_genome(genome[0], genome[1])
    n = int(cols["tid"].shape[0])
    bounds = [n * i // n_batches for i in range(n_batches + 1)]
    for i in range(n_batches):
        lo, hi = bounds[i], bounds[i + 1]
        if hi == lo:
            continue
        c0, c1 = int(cols["cigar_off"][lo]), int(cols["cigar_off"][hi])
        part = {k: np.ascontiguousarray(cols[k][lo:hi]) for k in ("tid", "pos", "flag", "nh_class", "xs")}
        part["cigar_off"] = np.ascontiguousarray(cols["cigar_off"][lo:hi + 1] - np.uint32(c0))
        part["cigar"] = np.ascontiguousarray(cols["cigar"][c0:c1])
        eng.push(part, seq_base=lo)
    return eng.flush()


def smoke():
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("smoke() needs a CUDA device")
    syn = Synth(200_000, kind=SHORT_UNSTRANDED, genome_len=200_000_000, n_genes=2000)
    cols = syn.columns()
    gpu = _native.Engine(device=0, profile=True)
    got = run_engine(gpu, syn.names, cols, n_batches=3)
    tm = gpu.timings()
    gpu.close()
    ora = oracle_engine()
    want = run_engine(ora, syn.names, cols)
    ora.close()
    diffs = compare_tables(got, want)
    if diffs:
        raise AssertionError("CUDA path differs from the CPU oracle: " + "; ".join(diffs))
    if got.n == 0 or tm["extract_launches"] == 0:
        raise AssertionError("smoke produced no junctions or launched no kernels")
    print("smoke ok: %d records -> %d junctions, %d events, extract %.3f ms over %d launches" % (
        syn.n, got.n, tm["events"], tm["extract_ms"], tm["extract_launches"]))
