#!/usr/bin/env python3
"""Device-timed stage breakdown of the other BASELINE.json configs (not the contract benchmark: that is bench.py).

    python profiles/workloads.py [short] [stranded] [long] [merge]

  short     configs[1]  100 M 2x150 unstranded records, no genome (what bench.py measures)
  stranded  configs[2]  stranded records, --strandness=RF --min-anchor-length=12, genome resident in HBM, canon filter
                        (scaled: 50 M records over a 1 Gb genome, so that one run stays within a minute)
  long      configs[3]  long reads (~10 kb, ~800 CIGAR operations), --allow-anchor-indels (scaled: 1 M reads)
  merge     configs[4]  intronProspectorMerge fold of 64 per-sample tables (rows already parsed, host columns)

Prints one JSON line per workload: records (rows) per second over clear + push_device + finalize_device, the stage
times from the library's own CUDA events, and unit counts.  Results of the run kept in profiles/ are quoted in DESIGN.md.
"""
import ctypes as C
import json
import os
import sys
import tempfile
import time

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))

from intronprospector_b200 import _native  # noqa: E402
from intronprospector_b200.synth import LONG, SHORT_STRANDED, SHORT_UNSTRANDED, Synth  # noqa: E402


def load_genome(syn, tmpdir):
    fa = os.path.join(tmpdir, "g.fa")
    syn.write_genome(fa)
    io = _native.load_io_library()
    g = C.c_void_p()
    assert io.ipxio_genome_load(fa.encode(), C.byref(g)) == 0
    n = io.ipxio_genome_n_seqs(g)
    names = [io.ipxio_genome_names(g)[i].decode() for i in range(n)]
    lens = [int(io.ipxio_genome_lengths(g)[i]) for i in range(n)]
    seqs = [C.string_at(io.ipxio_genome_seqs(g)[i], lens[i]) for i in range(n)]
    io.ipxio_genome_free(g)
    os.unlink(fa)
    return names, seqs


def extraction(name, n, kind, genome_len, steps=5, warmup=3, with_genome=False, **params):
    dev = torch.device("cuda", 0)
    syn = Synth(n, kind=kind, genome_len=genome_len, seed=20261018 + kind, model_seed=20261018)
    to_torch = {np.int32: torch.int32, np.uint16: torch.uint16, np.uint8: torch.uint8, np.uint32: torch.uint32}
    host = syn.columns(lambda shape, dtype: torch.empty(shape, dtype=to_torch[dtype], pin_memory=True))
    eng = _native.Engine(device=0, profile=True, **params)
    eng.set_targets(syn.names)
    if with_genome:
        with tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as td:
            g = load_genome(syn, td)
        eng.set_genome(g[0], g[1])
        del g
    n_rec, n_cigar = syn.n, syn.n_cigar
    syn.close()
    devc = {k: v.to(dev) for k, v in host.items()}
    stream = torch.cuda.ExternalStream(eng.stream(), device=dev)

    def step():
        eng.clear()
        eng.push_device(devc)
        eng.finalize_device()

    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    eng.reset_timings()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for a, b in ev:
        a.record(stream)
        step()
        b.record(stream)
    torch.cuda.synchronize()
    ms = sum(a.elapsed_time(b) for a, b in ev) / steps
    tm = eng.timings()
    eng.clear()
    eng.push_device(devc)
    table = eng.flush()
    line = {"workload": name, "records": n_rec, "cigar_words": n_cigar, "ms_per_step": ms, "records_per_s": n_rec / (ms / 1e3),
            "stage_ms": {k: tm[k + "_ms"] / steps for k in ("extract", "aggregate", "motif", "finalize")},
            "live_cigar_words": tm["cigar_live"] / steps, "events": tm["events"] / steps, "junctions": int(table.n),
            "hist_runs": int(table.hist_cnt.shape[0]), "params": {k: int(v) for k, v in params.items()}}
    alg = 16.0 * n_rec + 4.0 * line["live_cigar_words"] + 32.0 * line["events"]
    line["k1_algorithmic_GBps"] = alg / (line["stage_ms"]["extract"] / 1e3) / 1e9 if line["stage_ms"]["extract"] > 0 else None
    eng.close()
    print(json.dumps(line), flush=True)


def merge(n_samples=64, n_universe=400_000, steps=5, warmup=2):
    import merge_synth
    with tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as td:
        t0 = time.perf_counter()
        paths, _ = merge_synth.write_tables(td, n_samples, n_universe, n_contigs=25)
        gen_s = time.perf_counter() - t0
        io = _native.load_io_library()
        arr = (C.c_char_p * len(paths))(*[p.encode() for p in paths])
        calls = C.c_void_p()
        t0 = time.perf_counter()
        assert io.ipxio_calls_load(arr, len(paths), _native.FILTER_NULL, C.byref(calls)) == 0
        parse_s = time.perf_counter() - t0
    rows = io.ipxio_calls_rows(calls)
    names = [io.ipxio_calls_chroms(calls)[i].decode() for i in range(io.ipxio_calls_n_chroms(calls))]
    eng = _native.Engine(device=0, profile=True)
    eng.set_targets(names)
    for _ in range(warmup):
        eng.merge_rows(rows)[0].close()
    eng.reset_timings()
    t0 = time.perf_counter()
    for _ in range(steps):
        table, log = eng.merge_rows(rows)
        nj = table.n
        table.close()
    wall_ms = 1e3 * (time.perf_counter() - t0) / steps
    tm = eng.timings()
    n_rows = int(rows.contents.n_rows)
    print(json.dumps({"workload": "merge", "tables": n_samples, "rows": n_rows, "junctions": int(nj),
                      "device_fold_ms": tm["merge_ms"] / steps, "call_ms_host_rows_in_table_out": wall_ms,
                      "rows_per_s_call": n_rows / (wall_ms / 1e3), "tsv_parse_s": parse_s, "generate_s": gen_s}), flush=True)
    io.ipxio_calls_free(calls)
    eng.close()


if __name__ == "__main__":
    which = sys.argv[1:] or ["short", "stranded", "long", "merge"]
    if "short" in which:
        extraction("short unstranded 100M (configs[1])", 100_000_000, SHORT_UNSTRANDED, 3_000_000_000)
    if "stranded" in which:
        extraction("stranded RF + genome + canon filter, min anchor 12, 50M records / 1 Gb genome (configs[2] scaled)", 50_000_000,
                   SHORT_STRANDED, 1_000_000_000, with_genome=True, strandness=_native.STRAND_RF, min_anchor_length=12,
                   junction_filter=_native.FILTER_CANON)
    if "long" in which:
        extraction("long reads 1M, --allow-anchor-indels (configs[3] scaled)", 1_000_000, LONG, 3_000_000_000,
                   max_anchor_indel_size=36, hist_capacity=1 << 27)
    if "merge" in which:
        merge()
