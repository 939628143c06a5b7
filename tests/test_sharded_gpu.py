"""Multi-GPU path with the ranks emulated on one device: several ipx handles, each fed a contiguous shard of the
stream, exchange partial tables by key range (and, with the canon filter, agree on creation thresholds first).  The
concatenation of the owners' tables must equal the CPU oracle's table of the whole stream, bit for bit."""
import numpy as np
import pytest
import torch

import helpers  # noqa: F401
from intronprospector_b200 import _native, selftest, sharded
from intronprospector_b200.synth import SHORT_STRANDED, SHORT_UNSTRANDED, LONG, Synth
from test_parity_gpu import genome_of

pytestmark = pytest.mark.gpu


def shard_columns(cols, lo, hi):
    c0, c1 = int(cols["cigar_off"][lo]), int(cols["cigar_off"][hi])
    part = {k: np.ascontiguousarray(cols[k][lo:hi]) for k in ("tid", "pos", "flag", "nh_class", "xs")}
    part["cigar_off"] = np.ascontiguousarray(cols["cigar_off"][lo:hi + 1] - np.uint32(c0))
    part["cigar"] = np.ascontiguousarray(cols["cigar"][c0:c1])
    return part


def run_sharded(names, cols, world, genome=None, batches_per_rank=2, mailbox=None, rounds=1, **params):
    """mailbox = (junction rows, offset rows) per source: exchange through peer-memory mailboxes instead of packed
    all-to-all buffers.  rounds > 1 repeats the whole job on the same handles (ipx_clear in between)."""
    dev = torch.device("cuda", 0)
    n = int(cols["tid"].shape[0])
    canon = genome is not None and params.get("junction_filter") == _native.FILTER_CANON
    engines = [_native.Engine(device=0, deferred_counting=canon, **params) for _ in range(world)]
    bounds = [n * r // world for r in range(world + 1)]
    for r, eng in enumerate(engines):
        eng.set_targets(names)
        if genome is not None:
            eng.set_genome(genome[0], genome[1])
    if mailbox:
        sharded.connect_mailboxes_local(engines, mailbox[0], mailbox[1])
    tables = None
    for _ in range(rounds):
        for r, eng in enumerate(engines):
            eng.clear()
            sub = [bounds[r] + (bounds[r + 1] - bounds[r]) * b // batches_per_rank for b in range(batches_per_rank + 1)]
            for b in range(batches_per_rank):
                if sub[b + 1] > sub[b]:
                    eng.push(shard_columns(cols, sub[b], sub[b + 1]), seq_base=sub[b])
        if canon:
            sharded.agree_thresholds_local(engines, dev)
        sharded.exchange_local(engines, names, dev)
        tables = [eng.flush() for eng in engines]
    numbers = sharded.global_numbering([t.creation_seq for t in tables])
    for t, num in zip(tables, numbers):
        t.set_ijunc(num)
    for eng in engines:
        eng.close()
    return tables


def check_against_oracle(tables, names, cols, genome=None, **params):
    ora = selftest.oracle_engine(**params)
    want = selftest.run_engine(ora, names, cols, genome=genome)
    ora.close()
    assert sum(t.n for t in tables) == want.n
    assert want.n > 500
    wc = want.columns()
    for k in ("chrom", "intron_start", "intron_end", "anchor_start", "anchor_end", "strand", "splice_sites", "read_counts",
              "ijunc", "confidence", "hist_cnt"):
        got = np.concatenate([t.columns()[k] for t in tables])
        assert got.shape == wc[k].shape and np.array_equal(got, wc[k]), "column %s differs" % k
    for fmt in (_native.FMT_INTRON_TSV, _native.FMT_INTRON_BED9):
        for minc in (0.0, 1.0):
            assert b"".join(t.text(fmt, minc) for t in tables) == want.text(fmt, minc)
    # every rank owns a contiguous, disjoint range of chromosome names
    order = {t: r for r, t in enumerate(sorted(range(len(names)), key=lambda i: names[i].encode()))}
    last = -1
    for t in tables:
        if t.n:
            ranks = [order[int(c)] for c in t.chrom]
            assert min(ranks) > last
            last = max(ranks)


@pytest.mark.parametrize("world", [2, 3])
def test_unstranded_shards(built, world):
    syn = Synth(600_000, kind=SHORT_UNSTRANDED, genome_len=200_000_000, n_genes=2000, n_contigs=7)
    cols = syn.columns()
    tables = run_sharded(syn.names, cols, world)
    check_against_oracle(tables, syn.names, cols)


def test_long_read_shards(built):
    syn = Synth(20_000, kind=LONG, genome_len=200_000_000, n_genes=1500, n_contigs=5)
    cols = syn.columns()
    tables = run_sharded(syn.names, cols, 2, max_anchor_indel_size=36)
    check_against_oracle(tables, syn.names, cols, max_anchor_indel_size=36)


@pytest.mark.parametrize("world", [2, 4])
def test_canon_filter_shards(built, tmp_path, world):
    """Mixed read strands against the canon filter: the creating event of a junction may sit on another rank, and
    events before it must be dropped everywhere."""
    syn = Synth(400_000, kind=SHORT_UNSTRANDED, genome_len=100_000_000, n_genes=1000, n_contigs=5)
    cols = syn.columns()
    rng = np.random.default_rng(7)
    xs = cols["xs"]
    spliced = xs != 0
    r = rng.random(xs.shape[0])
    xs[spliced & (r < 0.25)] = 0
    xs[spliced & (r >= 0.25) & (r < 0.45)] = ord("-")
    xs[spliced & (r >= 0.45) & (r < 0.65)] = ord("+")
    genome = genome_of(syn, tmp_path)
    p = dict(junction_filter=_native.FILTER_CANON)
    tables = run_sharded(syn.names, cols, world, genome=genome, **p)
    check_against_oracle(tables, syn.names, cols, genome=genome, **p)


def test_stranded_canon_shards(built, tmp_path):
    syn = Synth(400_000, kind=SHORT_STRANDED, genome_len=100_000_000, n_genes=1000, n_contigs=5)
    cols = syn.columns()
    genome = genome_of(syn, tmp_path)
    p = dict(junction_filter=_native.FILTER_CANON, strandness=_native.STRAND_RF, min_anchor_length=12)
    tables = run_sharded(syn.names, cols, 2, genome=genome, **p)
    check_against_oracle(tables, syn.names, cols, genome=genome, **p)


# ---- the same jobs with the exchange through peer-memory mailboxes (ipx_mailbox_*, ipx_partial_*_mail) ----
@pytest.mark.parametrize("world", [2, 3])
def test_unstranded_shards_through_mailboxes(built, world):
    syn = Synth(600_000, kind=SHORT_UNSTRANDED, genome_len=200_000_000, n_genes=2000, n_contigs=7)
    cols = syn.columns()
    tables = run_sharded(syn.names, cols, world, mailbox=(100_000, 2_000_000), rounds=3)  # three exchanges: the slots alternate
    check_against_oracle(tables, syn.names, cols)


def test_long_read_shards_through_mailboxes(built):
    """long reads put most read offsets beyond the dense rows: the offset log is exported too"""
    syn = Synth(20_000, kind=LONG, genome_len=200_000_000, n_genes=1500, n_contigs=5)
    cols = syn.columns()
    tables = run_sharded(syn.names, cols, 2, mailbox=(100_000, 2_000_000), max_anchor_indel_size=36)
    check_against_oracle(tables, syn.names, cols, max_anchor_indel_size=36)


def test_canon_filter_shards_through_mailboxes(built, tmp_path):
    syn = Synth(400_000, kind=SHORT_UNSTRANDED, genome_len=100_000_000, n_genes=1000, n_contigs=5)
    cols = syn.columns()
    rng = np.random.default_rng(7)
    xs = cols["xs"]
    spliced = xs != 0
    r = rng.random(xs.shape[0])
    xs[spliced & (r < 0.25)] = 0
    xs[spliced & (r >= 0.25) & (r < 0.45)] = ord("-")
    xs[spliced & (r >= 0.45) & (r < 0.65)] = ord("+")
    genome = genome_of(syn, tmp_path)
    p = dict(junction_filter=_native.FILTER_CANON)
    tables = run_sharded(syn.names, cols, 4, genome=genome, mailbox=(100_000, 2_000_000), **p)
    check_against_oracle(tables, syn.names, cols, genome=genome, **p)


def test_a_mailbox_slot_that_is_too_small_is_reported(built):
    syn = Synth(300_000, kind=SHORT_UNSTRANDED, genome_len=200_000_000, n_genes=2000, n_contigs=7)
    with pytest.raises(_native.IpxError, match="mailbox"):
        run_sharded(syn.names, syn.columns(), 2, mailbox=(64, 1_000_000))
    with pytest.raises(_native.IpxError, match="mailbox"):
        run_sharded(syn.names, syn.columns(), 2, mailbox=(100_000, 256))
