"""Table-level parity through the C ABI: the CUDA engine against the CPU oracle on the same seeded batches,
bit-exact on every column (integer work) and on the formatted text (which carries the float32 confidence)."""
import numpy as np
import pytest

import helpers  # noqa: F401
from intronprospector_b200 import _native, selftest
from intronprospector_b200.synth import LONG, SHORT_STRANDED, SHORT_UNSTRANDED, Synth

pytestmark = pytest.mark.gpu


def both(names, cols, n_batches=1, genome=None, **params):
    gpu = _native.Engine(device=0, **params)
    got = selftest.run_engine(gpu, names, cols, genome=genome, n_batches=n_batches)
    gpu.close()
    ora = selftest.oracle_engine(**{k: v for k, v in params.items() if k not in ("table_capacity", "hist_capacity")})
    want = selftest.run_engine(ora, names, cols, genome=genome)
    ora.close()
    return got, want


def genome_of(syn, tmp_path):
    import ctypes as C
    fa = str(tmp_path / "g.fa")
    syn.write_genome(fa)
    io = _native.load_io_library()
    g = C.c_void_p()
    assert io.ipxio_genome_load(fa.encode(), C.byref(g)) == 0
    n = io.ipxio_genome_n_seqs(g)
    names = [io.ipxio_genome_names(g)[i].decode() for i in range(n)]
    lens = [int(io.ipxio_genome_lengths(g)[i]) for i in range(n)]
    seqs = [C.string_at(io.ipxio_genome_seqs(g)[i], lens[i]) for i in range(n)]
    io.ipxio_genome_free(g)
    return names, seqs


@pytest.mark.parametrize("n_batches", [1, 7])
def test_short_unstranded(built, n_batches):
    syn = Synth(2_000_000, kind=SHORT_UNSTRANDED, genome_len=300_000_000, n_genes=4000)
    got, want = both(syn.names, syn.columns(), n_batches=n_batches)
    assert want.n > 5000
    assert selftest.compare_tables(got, want) == []


def test_small_tables_force_collisions(built):
    syn = Synth(200_000, kind=SHORT_UNSTRANDED, genome_len=100_000_000, n_genes=500)
    got, want = both(syn.names, syn.columns(), n_batches=2, table_capacity=8192, hist_capacity=1 << 18)
    assert selftest.compare_tables(got, want) == []


@pytest.mark.parametrize("n_batches", [1, 5])
def test_tables_grow_from_tiny_capacities(built, n_batches):
    """1024 hash slots / 512 junction rows / a 1024-entry offset log against thousands of junctions: the fold leaves the
    events it has no room for in the event buffer, the host grows the tables (rows doubled, slots rehashed, log sized
    ahead of each batch) and repeats the fold until nothing is left -- same table as the oracle's."""
    syn = Synth(200_000, kind=SHORT_UNSTRANDED, genome_len=100_000_000, n_genes=500)
    got, want = both(syn.names, syn.columns(), n_batches=n_batches, table_capacity=1024, hist_capacity=1024)
    assert want.n > 1500
    assert selftest.compare_tables(got, want) == []


def test_tables_grow_with_the_canon_filter_and_long_reads(built, tmp_path):
    syn = Synth(300_000, kind=SHORT_STRANDED, genome_len=100_000_000, n_genes=1000, n_contigs=5)
    genome = genome_of(syn, tmp_path)
    got, want = both(syn.names, syn.columns(), n_batches=3, genome=genome, strandness=_native.STRAND_RF,
                     junction_filter=_native.FILTER_CANON, table_capacity=1024, hist_capacity=1024)
    assert want.n > 1000
    assert selftest.compare_tables(got, want) == []
    syn = Synth(20_000, kind=LONG, genome_len=300_000_000, n_genes=2000)
    got, want = both(syn.names, syn.columns(), n_batches=2, max_anchor_indel_size=36, table_capacity=2048, hist_capacity=1024)
    assert want.n > 1000
    assert selftest.compare_tables(got, want) == []


def test_more_targets_than_the_key_has_bits_for(built):
    """20 000 contigs at the default --max-intron-length (13 bits of the key are left for the target index), and a few
    hundred with -I 10000000 (8 bits): the tables compare the full target index kept with every junction."""
    rng = np.random.default_rng(3)
    def w(length, op):
        return (length << 4) | "MIDNSHP=XB".index(op)
    for n_targets, max_intron in ((20_000, 500_000), (600, 10_000_000)):
        names = ["scaffold_%d" % i for i in range(n_targets)]
        n = 60_000
        tid = np.sort(rng.integers(0, n_targets, n)).astype(np.int32)
        pos = rng.integers(0, 2000, n).astype(np.int32)
        order = np.lexsort((pos, tid))
        tid, pos = tid[order], pos[order]
        # the same few introns on every target: identical (start, length) under different target indexes
        starts = rng.integers(40, 60, n)
        lens = rng.choice([100, 250, 400_000 if max_intron >= 400_000 else 900], n)
        cigar = np.empty((n, 3), np.uint32)
        cigar[:, 0] = (starts << 4) | 0
        cigar[:, 1] = (lens << 4) | 3
        cigar[:, 2] = (50 << 4) | 0
        cols = {"tid": tid, "pos": pos, "flag": np.zeros(n, np.uint16), "nh_class": np.ones(n, np.uint8),
                "xs": np.full(n, ord("+"), np.uint8), "cigar_off": (np.arange(n + 1) * 3).astype(np.uint32),
                "cigar": cigar.reshape(-1)}
        got, want = both(names, cols, n_batches=2, max_intron_length=max_intron)
        assert want.n > 10_000
        assert selftest.compare_tables(got, want) == []


@pytest.mark.parametrize("flt", [_native.FILTER_CANON, _native.FILTER_ALL])
def test_stranded_with_genome(built, tmp_path, flt):
    syn = Synth(600_000, kind=SHORT_STRANDED, genome_len=100_000_000, n_genes=1000, n_contigs=5)
    genome = genome_of(syn, tmp_path)
    got, want = both(syn.names, syn.columns(), n_batches=4, genome=genome, strandness=_native.STRAND_RF,
                     min_anchor_length=12, junction_filter=flt)
    assert want.n > 1000
    assert selftest.compare_tables(got, want) == []


def test_unstranded_with_genome_mixed_strands(built, tmp_path):
    """XS-derived strands plus '.' reads against the canon filter: exercises strand correction and the rule that
    events before the first filter-passing event of a junction are dropped."""
    syn = Synth(400_000, kind=SHORT_UNSTRANDED, genome_len=100_000_000, n_genes=1000, n_contigs=5)
    cols = syn.columns()
    rng = np.random.default_rng(5)
    xs = cols["xs"]
    spliced = xs != 0
    r = rng.random(xs.shape[0])
    xs[spliced & (r < 0.25)] = 0             # no tag -> '.'
    xs[spliced & (r >= 0.25) & (r < 0.40)] = ord("-")
    xs[spliced & (r >= 0.40) & (r < 0.55)] = ord("+")
    xs[spliced & (r >= 0.55) & (r < 0.58)] = 0xFF   # XS present but not a character value
    xs[spliced & (r >= 0.58) & (r < 0.60)] = ord("?")
    genome = genome_of(syn, tmp_path)
    for flt in (_native.FILTER_CANON, _native.FILTER_ALL):
        got, want = both(syn.names, cols, n_batches=3, genome=genome, junction_filter=flt)
        assert selftest.compare_tables(got, want) == []


def test_missing_genome_targets_are_skipped(built, tmp_path):
    syn = Synth(200_000, kind=SHORT_UNSTRANDED, genome_len=100_000_000, n_genes=1000, n_contigs=5)
    names, seqs = genome_of(syn, tmp_path)
    genome = (names[:2] + names[3:], seqs[:2] + seqs[3:])
    got, want = both(syn.names, syn.columns(), genome=genome, junction_filter=_native.FILTER_CANON, skip_missing_targets=True)
    assert selftest.compare_tables(got, want) == []
    assert list(got.missing_targets) == list(want.missing_targets) == [2]


@pytest.fixture(params=[None, "600"], ids=["whole-records", "segmented"])
def segment_threshold(request, monkeypatch):
    """IPX_K1_SEG_MIN: records with more operations than this are walked segment by segment (default 16384: only
    outliers); 600 sends a good part of every long-read test through the segmented kernels."""
    if request.param is not None:
        monkeypatch.setenv("IPX_K1_SEG_MIN", request.param)
    return request.param


@pytest.mark.parametrize("indel", [0, 36])
def test_long_reads(built, indel, segment_threshold):
    syn = Synth(30_000, kind=LONG, genome_len=300_000_000, n_genes=2000)
    got, want = both(syn.names, syn.columns(), n_batches=3, max_anchor_indel_size=indel)
    assert want.n > 1000
    assert selftest.compare_tables(got, want) == []


def _hand_made_long_batch(extra=()):
    """A batch in the long-read regime (more than 8 operations per record on average) built from random operation
    soup: every state transition, clips in the middle, dead and one-operation records, records of every length
    modulo the kernel's 8-word groups."""
    rng = np.random.default_rng(77)
    ops = np.array([0, 0, 7, 7, 8, 1, 2, 3, 3, 4, 5], np.uint32)  # M = X I D N S H
    recs = []
    pos = 1000
    for i in range(4000):
        n = int(rng.integers(0, 3)) if i % 50 == 0 else int(rng.integers(2, 90))
        if i % 400 == 7:
            n = int(rng.integers(1500, 4000))
        op = ops[rng.integers(0, len(ops), n)]
        ln = np.where(rng.random(n) < 0.3, rng.integers(1, 4, n), rng.integers(1, 120, n)).astype(np.uint32)
        ln = np.where(op == 3, rng.integers(60, 3000, n), ln).astype(np.uint32)
        ln[rng.random(n) < 0.01] = 0
        flag = 0x400 if i % 97 == 0 else (0x100 if i % 89 == 0 else (16 if i % 2 else 0))
        recs.append((0 if i % 211 else -1, pos, flag, ((ln << 4) | op).tolist()))
        pos += int(rng.integers(0, 40))
    recs.extend(extra)
    cols = {"tid": np.array([r[0] for r in recs], np.int32), "pos": np.array([r[1] for r in recs], np.int32),
            "flag": np.array([r[2] for r in recs], np.uint16), "nh_class": (np.arange(len(recs)) % 3).astype(np.uint8),
            "xs": np.where(np.arange(len(recs)) % 4 == 0, ord("+"), np.where(np.arange(len(recs)) % 4 == 1, ord("-"), 0)).astype(np.uint8),
            "cigar_off": np.cumsum([0] + [len(r[3]) for r in recs]).astype(np.uint32),
            "cigar": np.array([x for r in recs for x in r[3]], np.uint32)}
    assert cols["cigar"].shape[0] > 8 * len(recs)
    return cols


@pytest.mark.parametrize("indel", [0, 5, 36])
def test_long_regime_operation_soup(built, indel, segment_threshold):
    cols = _hand_made_long_batch()
    for n_batches in (1, 3):
        got, want = both(["c1"], cols, n_batches=n_batches, max_anchor_indel_size=indel, min_anchor_length=3)
        assert want.n > 500
        assert selftest.compare_tables(got, want) == []


def test_long_regime_positions_beyond_32_bit_offsets(built, segment_threshold):
    """A CIGAR whose reference span outgrows the long-read walker's 32-bit relative positions: the walk is handed to
    the 64-bit automaton mid-record and the junction after the hand-over is still found where the oracle finds it."""
    def w(length, op):
        return (length << 4) | "MIDNSHP=XB".index(op)
    big = (1 << 28) - 1
    giant = [w(30, "M")] + [w(big, "D")] * 7 + [w(5000, "M"), w(40, "M"), w(300, "N"), w(45, "M"), w(2, "I"), w(20, "M"),
                                                  w(900, "N"), w(33, "M")] + [w(3, "M"), w(1, "X")] * 6
    # the same span inside a record long enough to be walked segment by segment (more than 1024 operations)
    filler = [w(20, "M"), w(1, "X"), w(2, "I"), w(30, "="), w(1, "D")] * 300
    very_long = filler[:700] + giant + filler
    cols = _hand_made_long_batch(extra=[(0, 100, 0, giant), (0, 120, 16, giant[:-3]), (0, 90, 0, very_long)])
    got, want = both(["c1"], cols, max_anchor_indel_size=36)
    assert int(np.max(want.intron_start)) > 7 * big
    assert selftest.compare_tables(got, want) == []


def test_long_regime_unknown_cigar_op_is_reported(built, segment_threshold):
    def w(length, op):
        return (length << 4) | op
    bad_late = [w(10, 0)] * 40 + [w(3, 9)] + [w(10, 0)] * 5   # 'B' deep inside a whole group of 8
    bad_early = [w(10, 0), w(4, 6), w(10, 0)] + [w(7, 0)] * 30  # 'P' among the record's first words
    cols = _hand_made_long_batch(extra=[(0, 500, 0, bad_late), (0, 600, 0, bad_early)])
    got, want = both(["c1"], cols)
    assert got.status == want.status == -3
    assert got.error_text == want.error_text == "Unknown cigar operation 'B'"
    assert got.first_error_seq == want.first_error_seq == 4000
    # ... and inside a record that is walked segment by segment
    segmented = [w(10, 0), w(1, 8)] * 900 + [w(3, 6)] + [w(10, 0)] * 700
    cols = _hand_made_long_batch(extra=[(0, 500, 0, segmented), (0, 600, 0, bad_early)])
    got, want = both(["c1"], cols)
    assert got.status == want.status == -3
    assert got.error_text == want.error_text == "Unknown cigar operation 'P'"
    assert got.first_error_seq == want.first_error_seq == 4000


@pytest.mark.parametrize("bounds", ["64,64,64", "1024,200000,256", "100000,64,100000"])
def test_finalize_grows_from_bounds_that_are_too_small(built, monkeypatch, bounds):
    """The device finalize runs inside buffers sized by grow-only bounds (junctions, offset-log entries, histogram
    runs) and runs again when one was too small.  Long reads put most read offsets into the log, so every combination
    of a bound that is too small with data that needs it comes up here."""
    monkeypatch.setenv("IPX_FIN_BOUNDS", bounds)
    syn = Synth(30_000, kind=LONG, genome_len=300_000_000, n_genes=2000)
    got, want = both(syn.names, syn.columns(), n_batches=2, max_anchor_indel_size=36)
    assert want.n > 1000
    assert selftest.compare_tables(got, want) == []


@pytest.mark.parametrize("bounds", ["64,64,64", "70000,64,4096"])
def test_finalize_grows_from_small_bounds_short_reads(built, tmp_path, monkeypatch, bounds):
    """The same for short reads (offsets in the dense rows, a few in the log), with and without the canon filter, and
    over several flushes of one handle (the bounds only ever grow)."""
    monkeypatch.setenv("IPX_FIN_BOUNDS", bounds)
    syn = Synth(400_000, kind=SHORT_UNSTRANDED, genome_len=100_000_000, n_genes=1000, n_contigs=5)
    cols = syn.columns()
    got, want = both(syn.names, cols, n_batches=3)
    assert want.n > 3000
    assert selftest.compare_tables(got, want) == []
    genome = genome_of(syn, tmp_path)
    got, want = both(syn.names, cols, n_batches=2, genome=genome, junction_filter=_native.FILTER_CANON)
    assert selftest.compare_tables(got, want) == []
    # one handle, three jobs of different sizes: small, large, small again (sjN keeps counting across the flushes of a
    # handle, junctions_extractor.hh:129, so the oracle handle lives as long)
    gpu, ora = _native.Engine(device=0), selftest.oracle_engine()
    for n_rec in (20_000, 400_000, 50_000):
        part = {k: (v[:n_rec] if k not in ("cigar", "cigar_off") else v) for k, v in cols.items()}
        part["cigar_off"] = cols["cigar_off"][:n_rec + 1]
        part["cigar"] = cols["cigar"][:int(cols["cigar_off"][n_rec])]
        got = selftest.run_engine(gpu, syn.names, part)
        want = selftest.run_engine(ora, syn.names, part)
        assert want.n > 200
        assert selftest.compare_tables(got, want) == []
    gpu.close()
    ora.close()


def test_empty_and_ragged_batches(built):
    names = ["c1", "c2"]
    empty = {k: np.zeros(0, dt) for k, dt in (("tid", np.int32), ("pos", np.int32), ("flag", np.uint16), ("nh_class", np.uint8),
                                              ("xs", np.uint8), ("cigar", np.uint32))}
    empty["cigar_off"] = np.zeros(1, np.uint32)
    got, want = both(names, empty)
    assert got.n == want.n == 0
    # records with no CIGAR, unmapped, duplicates, one-op reads and a real junction mixed together
    def w(length, op):
        return (length << 4) | "MIDNSHP=XB".index(op)
    recs = [
        (0, 100, 0, [w(50, "M"), w(200, "N"), w(50, "M")]),
        (-1, 0, 4, []),
        (0, 120, 1024, [w(30, "M"), w(180, "N"), w(70, "M")]),
        (1, 5, 0, [w(100, "M")]),
        (1, 7, 16, [w(5, "S"), w(20, "M"), w(100, "N"), w(10, "M"), w(300, "N"), w(40, "M"), w(3, "H")]),
        (1, 7, 0, [w(20, "M"), w(100, "N"), w(4, "I"), w(300, "N"), w(40, "M")]),
    ]
    cols = {"tid": np.array([r[0] for r in recs], np.int32), "pos": np.array([r[1] for r in recs], np.int32),
            "flag": np.array([r[2] for r in recs], np.uint16), "nh_class": np.array([1, 0, 1, 1, 2, 0], np.uint8),
            "xs": np.array([ord("+"), 0, 0, 0, ord("-"), 0], np.uint8),
            "cigar_off": np.cumsum([0] + [len(r[3]) for r in recs]).astype(np.uint32),
            "cigar": np.array([x for r in recs for x in r[3]], np.uint32)}
    for indel in (0, 36):
        got, want = both(names, cols, max_anchor_indel_size=indel)
        assert want.n >= 2
        assert selftest.compare_tables(got, want) == []


def test_unknown_cigar_op_is_reported(built):
    def w(length, op):
        return (length << 4) | op
    cols = {"tid": np.array([0, 0], np.int32), "pos": np.array([10, 20], np.int32), "flag": np.zeros(2, np.uint16),
            "nh_class": np.ones(2, np.uint8), "xs": np.zeros(2, np.uint8), "cigar_off": np.array([0, 3, 6], np.uint32),
            "cigar": np.array([w(50, 0), w(100, 3), w(50, 0), w(50, 0), w(3, 6), w(50, 0)], np.uint32)}
    got, want = both(["c1"], cols)
    assert got.status == want.status == -3
    assert got.error_text == want.error_text == "Unknown cigar operation 'P'"
    assert got.first_error_seq == want.first_error_seq == 1


def test_compact_batches_equal_plain_batches(built):
    """ipx_push_batch_compact (operation counts + target runs, expanded on the device) against ipx_push_batch."""
    syn = Synth(700_000, kind=SHORT_UNSTRANDED, genome_len=300_000_000, n_genes=3000, n_contigs=9)
    cols = syn.columns()
    cols["tid"][1000:1010] = -1  # a few unmapped records in the middle: more runs, negative targets
    n = int(cols["tid"].shape[0])
    plain = _native.Engine(device=0)
    want = selftest.run_engine(plain, syn.names, cols, n_batches=3)
    plain.close()
    eng = _native.Engine(device=0)
    eng.set_targets(syn.names)
    bounds = [0, n // 3, n // 3 + 1, n]
    for lo, hi in zip(bounds[:-1], bounds[1:]):
        c0, c1 = int(cols["cigar_off"][lo]), int(cols["cigar_off"][hi])
        part = {k: np.ascontiguousarray(cols[k][lo:hi]) for k in ("tid", "pos", "flag", "nh_class", "xs")}
        part["cigar_off"] = np.ascontiguousarray(cols["cigar_off"][lo:hi + 1] - np.uint32(c0))
        part["cigar"] = np.ascontiguousarray(cols["cigar"][c0:c1])
        eng.push_compact(_native.Engine.compact_columns(part), seq_base=lo)
    got = eng.flush()
    eng.close()
    assert want.n > 5000
    assert selftest.compare_tables(got, want) == []
