#!/usr/bin/env python3
"""bench.py -- headline benchmark of the junction-extraction hot path (BASELINE.json: aligned reads/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--records R] [--impl ours|reference] [--quick]

Workload (config.workload): BASELINE.json configs[1], a synthetic coordinate-sorted 2x150 short-read unstranded
RNA-Seq stream of 100 M alignment records over a 25-contig 3 Gb genome (seeded generator, csrc/synth).  One "step" is
one pass of the hot path over that stream: clear the junction table, extract + aggregate every record, finish the
table on the device (key order, sjN ranking, ordered offset histograms).

  value        records/s with the record columns already resident in HBM, device-timed with CUDA events on the
               launching stream (max over ranks), whole job (all ranks' records / slowest rank's time)
  e2e          the same pass through the public C ABI with HOST (pinned) columns: ipx_push_batch_compact (H2D copies
               inside the timed region) + ipx_flush (table D2H + float32 confidences on the host)
  roofline     the extract kernel K1: algorithmic bytes (16 B per record + 4 B per live CIGAR word + 32 B per junction
               event) / its CUDA-event time, against the measured HBM copy bandwidth in MEASURED_PEAKS.json
  cpu_baseline the unmodified reference binary (oracle/_ref/intronProspector, 1 thread) on a bounded prefix of the
               same stream written as an uncompressed BAM, timed on this box's host cores
  parity       (N = 1) the GPU path over the same prefix through ipx_push_batch_compact + ipx_flush + ipx_format against
               the three files that reference run wrote: byte-identical or the bench fails
  other_workloads  (N = 1) BASELINE configs[2..4]: stranded + 3 Gb genome + canon filter, long reads (both option sets),
               intronProspectorMerge of 64 tables -- device-timed with the K1 roofline, and tool-level parity: the GPU
               command-line tool against the reference binary on the same files
  cli_e2e      (N = 1) what a user of the tools waits for: both command-line tools on the same BGZF-compressed BAM of
               aligner-style records (names, bases, qualities, a dozen tags), wall clock, outputs compared

With N > 1 (torchrun, one process per GPU) every rank extracts its own 100 M-record shard (weak scaling; shards are
consecutive in the global stream, like `intronProspector a.bam b.bam ...`), then the junctions a rank does not own are
stored into their owners' memory over NVLink and folded there (no collective); `phases_ms` splits the step.

`--impl reference` times the reference's own CPU implementation instead (all host cores, one process per shard).
"""
import argparse
import ctypes as C
import hashlib
import json
import os
import shutil
import statistics
import subprocess
import sys
import tempfile
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

METRIC = "aligned_reads_per_sec"
UNIT = "records/s"
WORKLOAD = "synthetic 100M-read 2x150 short-read unstranded RNA-Seq, coordinate sorted, 25 contigs / 3 Gb (BASELINE configs[1])"
OPTIONS = "defaults (-C 1.0), -c -j -n"
REF_BIN = os.path.join(REPO, "oracle", "_ref", "intronProspector")
REF_MERGE = os.path.join(REPO, "oracle", "_ref", "intronProspectorMerge")
HTS_TOOL = os.path.join(REPO, "oracle", "_ref", "hts_tool")
GPU_BIN = os.path.join(REPO, "intronprospector_b200", "csrc", "build", "intronProspector")
GPU_MERGE = os.path.join(REPO, "intronprospector_b200", "csrc", "build", "intronProspectorMerge")
SHM = "/dev/shm" if os.path.isdir("/dev/shm") else None


# Libraries (NCCL, torchrun banners) may print to stdout; the contract is ONE JSON line there.  Everything this process
# prints goes to stderr from here on, and emit() alone writes to the real stdout.
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(line):
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


def log(msg):
    sys.stderr.write("[bench %.1fs] %s\n" % (time.perf_counter() - _T0, msg))
    sys.stderr.flush()


_T0 = time.perf_counter()


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


class ClockSampler:
    """nvidia-smi clocks / throttle reasons while the timed region runs (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if f[5 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_hbm_peak():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except (KeyError, ValueError):
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic():
    """DRAM bytes per record of the dominant kernel from the committed `ncu --set full` capture, with the capture's
    code version: stale numbers are labelled, not passed off as this run's."""
    p = os.path.join(REPO, "profiles", "traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p))
        except ValueError:
            return None
    return None


def sha(data):
    return hashlib.sha256(data).hexdigest()[:16]


def outputs_of(prefix):
    return {k: prefix + "." + k for k in ("introns.tsv", "juncs.bed", "introns.bed")}


def tool_command(binary, prefix, inputs, extra=()):
    o = outputs_of(prefix)
    return [binary, "--intron-calls=" + o["introns.tsv"], "--junction-bed=" + o["juncs.bed"], "--intron-bed=" + o["introns.bed"]] + \
        list(extra) + list(inputs)


def run_tool(binary, prefix, inputs, extra=(), env=None):
    """One command-line run; returns (wall seconds, {output: bytes}, stderr)."""
    t0 = time.perf_counter()
    proc = subprocess.run(tool_command(binary, prefix, inputs, extra), stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, env=env)
    dt = time.perf_counter() - t0
    if proc.returncode != 0:
        raise RuntimeError("%s failed: %s" % (os.path.basename(binary), proc.stderr.decode(errors="replace")[-500:]))
    return dt, {k: open(p, "rb").read() for k, p in outputs_of(prefix).items()}, proc.stderr.decode(errors="replace")


class Background:
    """A reference run started early and collected late (the reference is single-threaded; the box has cores to spare).
    A thread waits for the process, so that `seconds` is the run's own wall clock, not the time until it was collected."""

    def __init__(self, binary, prefix, inputs, extra=()):
        self.prefix = prefix
        self.t0 = time.perf_counter()
        self.proc = subprocess.Popen(tool_command(binary, prefix, inputs, extra), stdout=subprocess.DEVNULL, stderr=subprocess.PIPE)
        self._err = b""
        self._end = None
        self._waiter = threading.Thread(target=self._wait, daemon=True)
        self._waiter.start()

    def _wait(self):
        _, self._err = self.proc.communicate()
        self._end = time.perf_counter()

    def collect(self):
        self._waiter.join()
        self.seconds = self._end - self.t0
        if self.proc.returncode != 0:
            raise RuntimeError("reference failed: " + self._err.decode(errors="replace")[-500:])
        self.stderr = self._err.decode(errors="replace")
        return {k: open(p, "rb").read() for k, p in outputs_of(self.prefix).items()}


def compare_outputs(got, want):
    """{'identical': bool, per-file sha of the reference's outputs, first difference}"""
    res = {"identical": True, "sha256": {k: sha(v) for k, v in want.items()}}
    for k in want:
        if got[k] != want[k]:
            res["identical"] = False
            a, b = got[k].split(b"\n"), want[k].split(b"\n")
            first = next((i for i, (x, y) in enumerate(zip(a, b)) if x != y), min(len(a), len(b)))
            res.setdefault("differences", []).append("%s: line %d: %r vs reference %r (%d vs %d lines)" % (
                k, first + 1, a[first][:120] if first < len(a) else None, b[first][:120] if first < len(b) else None, len(a), len(b)))
    return res


def time_reference(bam_paths, tmpdir):
    """Run the unmodified reference over every file concurrently (it is single-threaded: one process per file)."""
    procs = []
    t0 = time.perf_counter()
    for i, bam in enumerate(bam_paths):
        procs.append(subprocess.Popen(tool_command(REF_BIN, os.path.join(tmpdir, "ref%d" % i), [bam]), stdout=subprocess.DEVNULL,
                                      stderr=subprocess.PIPE))
    for p in procs:
        _, err = p.communicate()
        if p.returncode != 0:
            raise RuntimeError("reference failed: " + err.decode(errors="replace")[-500:])
    return time.perf_counter() - t0


def time_decode(bam):
    t0 = time.perf_counter()
    subprocess.run([HTS_TOOL, "count", bam], stdout=subprocess.DEVNULL, check=True)
    return time.perf_counter() - t0


def run_reference_arm(args):
    """--impl reference: the reference's own CPU path on this box's host cores, bounded sample per step."""
    rank = env_int("RANK", 0)
    if rank != 0:
        return 0
    if not os.path.exists(REF_BIN):
        emit({"impl": "reference", "unavailable": "oracle/_ref/intronProspector has not been built"})
        return 0
    from intronprospector_b200.synth import SHORT_UNSTRANDED, Synth
    cores = min(os.cpu_count() or 1, 32)
    per = args.ref_sample
    tmpdir = tempfile.mkdtemp(prefix="ipx_ref_", dir=SHM)
    bams = []
    try:
        for i in range(cores):
            syn = Synth(per, kind=SHORT_UNSTRANDED, seed=20261018 + i, model_seed=20261018)
            path = os.path.join(tmpdir, "shard%d.bam" % i)
            syn.write_raw_bam(path)
            syn.close()
            bams.append(path)
        for _ in range(args.warmup):
            time_reference(bams[:1], tmpdir)
        times = [time_reference(bams, tmpdir) for _ in range(args.steps)]
        decode_s = time_decode(bams[0])
    finally:
        shutil.rmtree(tmpdir, ignore_errors=True)
    total = per * cores
    ms = 1000.0 * sum(times) / len(times)
    value = total / (ms / 1000.0)
    sample = ("%d concurrent single-threaded reference processes, each on the first %d records of a shard of the workload "
              "(uncompressed BAM in tmpfs); the reference's time is linear in records, so records/s carries over to the "
              "full 100 M-record shards" % (cores, per))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int64", "data": "synthetic",
        # the same keys as the GPU arm's config: the workload and options are the same, the step is a bounded sample of it
        "config": {"workload": WORKLOAD, "records_per_gpu": args.records, "options": OPTIONS, "sample": sample,
                   "records_per_step": total, "decode_only_s_per_shard": decode_s},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "reference", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


# ---------------------------------------------------------------------------------------------------------------------
def device_timed(eng, devc, steps, warmup, torch, dev):
    """clear + push_device + finalize_device, CUDA events on the engine's stream; returns (ms per step, timings)"""
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
    return sum(a.elapsed_time(b) for a, b in ev) / steps, eng.timings()


def roofline_of(n_rec, tm, steps, peak):
    k_ms = tm["extract_ms"] / steps
    live, events = tm["cigar_live"] / steps, tm["events"] / steps
    alg = 16.0 * n_rec + 4.0 * live + 32.0 * events
    achieved = alg / (k_ms / 1000.0) / 1e9 if k_ms > 0 else 0.0
    return {"kernel_ms": k_ms, "algorithmic_bytes_per_launch": alg, "achieved": achieved, "frac": achieved / peak if peak else None,
            "live_cigar_words": live, "junction_events": events}


def load_genome(fa):
    from intronprospector_b200 import _native
    io = _native.load_io_library()
    g = C.c_void_p()
    if io.ipxio_genome_load(fa.encode(), C.byref(g)) != 0:
        raise RuntimeError("cannot load " + fa)
    n = io.ipxio_genome_n_seqs(g)
    names = [io.ipxio_genome_names(g)[i].decode() for i in range(n)]
    lens = [int(io.ipxio_genome_lengths(g)[i]) for i in range(n)]
    seqs = [C.string_at(io.ipxio_genome_seqs(g)[i], lens[i]) for i in range(n)]
    io.ipxio_genome_free(g)
    return names, seqs


def extraction_workload(name, n_rec, kind, tmpdir, torch, dev, peak, parity_records, with_genome=False, cli_sets=(), steps=5, warmup=3,
                        **params):
    """One of the other extraction configs: device-timed pass + K1 roofline, and tool-level parity on a prefix.
    cli_sets: [(label, [reference CLI options])]; the first set's options correspond to **params."""
    import numpy as np
    from intronprospector_b200 import _native
    from intronprospector_b200.synth import Synth
    t0 = time.perf_counter()
    syn = Synth(n_rec, kind=kind, seed=20261018 + kind, model_seed=20261018)
    to_torch = {np.int32: torch.int32, np.uint16: torch.uint16, np.uint8: torch.uint8, np.uint32: torch.uint32}
    host = syn.columns(lambda shape, dtype: torch.empty(shape, dtype=to_torch[dtype], pin_memory=True))
    bam = os.path.join(tmpdir, "w%d.bam" % kind)
    syn.write_raw_bam(bam, 0, min(parity_records, syn.n))
    fa = None
    if with_genome:
        fa = os.path.join(tmpdir, "g.fa")
        syn.write_genome(fa)
    gen_s = time.perf_counter() - t0
    # the reference runs start now and are collected after the GPU measurement
    refs = [(label, opts, Background(REF_BIN, os.path.join(tmpdir, "ref_%d_%d" % (kind, i)), [bam], opts))
            for i, (label, opts) in enumerate(cli_sets)]
    eng = _native.Engine(device=dev.index or 0, profile=True, **params)
    eng.set_targets(syn.names)
    if with_genome:
        g = load_genome(fa)
        eng.set_genome(g[0], g[1])
        del g
    n, n_cigar = syn.n, syn.n_cigar
    syn.close()
    devc = {k: v.to(dev) for k, v in host.items()}
    del host
    ms, tm = device_timed(eng, devc, steps, warmup, torch, dev)
    eng.clear()
    eng.push_device(devc)
    table = eng.flush()
    line = {"workload": name, "records": n, "cigar_words": n_cigar, "ms_per_step": ms, "value": n / (ms / 1e3), "unit": UNIT,
            "stage_ms": {k: tm[k + "_ms"] / steps for k in ("extract", "aggregate", "motif", "finalize")},
            "junctions": int(table.n), "hist_runs": int(table.n_hist), "generate_s": gen_s,
            "roofline": dict(roofline_of(n, tm, steps, peak), bound="hbm", peak=peak, unit="GB/s")}
    table.close()
    eng.close()
    del devc
    torch.cuda.empty_cache()
    parity = []
    for i, (label, opts, bg) in enumerate(refs):
        want = bg.collect()
        _, got, _ = run_tool(GPU_BIN, os.path.join(tmpdir, "gpu_%d_%d" % (kind, i)), [bam], opts)
        res = compare_outputs(got, want)
        res.update({"options": " ".join(opts) or "(defaults)", "records": min(parity_records, n), "reference_s": bg.seconds,
                    "how": "GPU command-line tool vs reference binary on the same BAM, three output files"})
        parity.append(res)
    line["parity"] = parity
    for f in os.listdir(tmpdir):
        if f.startswith(("ref_", "gpu_")) or f in ("g.fa", "g.fa.fai", os.path.basename(bam)):
            os.unlink(os.path.join(tmpdir, f))
    return line


def merge_workload(tmpdir, n_samples=64, n_universe=400_000):
    """BASELINE configs[4]: both merge tools on the same 64 tables (tool-level parity incl. stderr), the library call
    and its device fold timed."""
    sys.path.insert(0, os.path.join(REPO, "tests"))
    import merge_synth
    from intronprospector_b200 import _native
    t0 = time.perf_counter()
    paths, _ = merge_synth.write_tables(tmpdir, n_samples, n_universe, n_contigs=25)
    gen_s = time.perf_counter() - t0
    ref = Background(REF_MERGE, os.path.join(tmpdir, "merge_ref"), paths)  # ~30 s on one core: collected at the very end
    io = _native.load_io_library()
    arr = (C.c_char_p * len(paths))(*[p.encode() for p in paths])
    calls = C.c_void_p()
    t0 = time.perf_counter()
    if io.ipxio_calls_load(arr, len(paths), _native.FILTER_NULL, C.byref(calls)) != 0:
        raise RuntimeError("cannot parse the call tables")
    parse_s = time.perf_counter() - t0
    rows = io.ipxio_calls_rows(calls)
    names = [io.ipxio_calls_chroms(calls)[i].decode() for i in range(io.ipxio_calls_n_chroms(calls))]
    eng = _native.Engine(device=0, profile=True)
    eng.set_targets(names)
    for _ in range(2):
        eng.merge_rows(rows)[0].close()
    eng.reset_timings()
    steps = 3
    t0 = time.perf_counter()
    for _ in range(steps):
        table, _ = eng.merge_rows(rows)
        nj = table.n
        table.close()
    call_ms = 1e3 * (time.perf_counter() - t0) / steps
    tm = eng.timings()
    n_rows = int(rows.contents.n_rows)
    io.ipxio_calls_free(calls)
    eng.close()
    fold_ms = tm["merge_ms"] / steps
    alg = 42.0 * n_rows + 48.0 * nj
    t_cli, got, got_err = run_tool(GPU_MERGE, os.path.join(tmpdir, "merge_gpu"), paths)
    # the same with the rows parsed by the host's thread pool instead of the device tokeniser
    t_cli_host, got_h, got_err_h = run_tool(GPU_MERGE, os.path.join(tmpdir, "merge_gpu_h"), paths, env=dict(os.environ, IPX_TSV_HOST="1"))
    if got_h != got or got_err_h != got_err:
        raise RuntimeError("the device tokeniser and the host parser disagree")
    line = {"workload": "intronProspectorMerge of %d per-sample tables (BASELINE configs[4]), one GPU" % n_samples, "tables": n_samples,
            "rows": n_rows, "junctions": int(nj), "generate_s": gen_s, "tsv_parse_s": parse_s, "call_ms_host_rows_in_table_out": call_ms,
            "value": n_rows / (call_ms / 1e3), "unit": "rows/s", "device_fold_ms": fold_ms, "tool_wall_s": t_cli,
            "tool_wall_host_tsv_parser_s": t_cli_host, "devices_visible_to_the_tool": _native.load_library().ipx_device_count(),
            "roofline": {"bound": "hbm", "kernel": "k_merge_fold (K4)", "kernel_ms": fold_ms, "algorithmic_bytes_per_launch": alg,
                         "achieved": alg / (fold_ms / 1e3) / 1e9 if fold_ms > 0 else 0.0, "unit": "GB/s"}}
    return line, ref, got, got_err


def cli_end_to_end(tmpdir, n_rec):
    """Both command-line tools on the same BGZF BAM of aligner-style records: wall clock and outputs."""
    from intronprospector_b200.synth import SHORT_UNSTRANDED, Synth
    t0 = time.perf_counter()
    syn = Synth(n_rec, kind=SHORT_UNSTRANDED, seed=20261021, model_seed=20261018)
    bam = os.path.join(tmpdir, "aligner_style.bam")
    syn.write_full_bam(bam, threads=0, level=1)
    n = syn.n
    syn.close()
    gen_s = time.perf_counter() - t0
    size = os.path.getsize(bam)
    # the reference binary reads through this repo's htslib stand-in, whose BGZF reader would inflate on several threads:
    # one thread, as htslib does unless asked otherwise (the reference never asks)
    one_thread = dict(os.environ, IPX_DECODE_THREADS="1")
    t_ref, want, _ = run_tool(REF_BIN, os.path.join(tmpdir, "ref_cli"), [bam], env=one_thread)
    t_gpu, got, _ = run_tool(GPU_BIN, os.path.join(tmpdir, "gpu_cli"), [bam])  # first run: includes CUDA context + table set-up
    t_gpu2, got2, _ = run_tool(GPU_BIN, os.path.join(tmpdir, "gpu_cli"), [bam])
    res = compare_outputs(got, want)
    res2 = compare_outputs(got2, want)
    os.unlink(bam)
    return {"records": n, "bam_bytes": size, "bytes_per_record_uncompressed": "~300 (name, 150 bases + qualities, NH HI AS nM XS jM jI tags)",
            "reference_s": t_ref, "gpu_tool_s": min(t_gpu, t_gpu2), "gpu_tool_first_run_s": t_gpu,
            "reference_records_per_s": n / t_ref, "gpu_tool_records_per_s": n / min(t_gpu, t_gpu2),
            "speedup": t_ref / min(t_gpu, t_gpu2), "host_cores": os.cpu_count(), "reference_threads": 1,
            "identical": res["identical"] and res2["identical"], "sha256": res["sha256"], "differences": res.get("differences"),
            "generate_s": gen_s, "how": "wall clock of `intronProspector -c -j -n file.bam`, BGZF level 1, file in tmpfs"}


# ---------------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--records", type=int, default=100_000_000, help="records per GPU (the named workload is 100M)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--ref-sample", type=int, default=4_000_000, help="records per reference process per step")
    ap.add_argument("--cpu-sample", type=int, default=40_000_000, help="records of the cpu_baseline / parity sample")
    ap.add_argument("--cli-records", type=int, default=20_000_000, help="records of the tool-level end-to-end run")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--quick", action="store_true", help="headline measurement only (profiling runs)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)

    import numpy as np
    import torch
    import torch.distributed as dist

    from intronprospector_b200 import _native, sharded
    from intronprospector_b200.synth import LONG, SHORT_STRANDED, SHORT_UNSTRANDED, Synth

    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product has no CPU path)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    extras = rank == 0 and world == 1 and not args.quick and not args.no_cpu_baseline and os.path.exists(REF_BIN)
    peak, peak_src = measured_hbm_peak()

    # ---- synthetic shard of this rank, in pinned host memory ----
    t_gen = time.perf_counter()
    syn = Synth(args.records, kind=SHORT_UNSTRANDED, seed=20261018 + rank, model_seed=20261018)
    to_torch = {np.int32: torch.int32, np.uint16: torch.uint16, np.uint8: torch.uint8, np.uint32: torch.uint32}
    host = syn.columns(lambda shape, dtype: torch.empty(shape, dtype=to_torch[dtype], pin_memory=True))
    names = syn.names
    n, n_cigar = syn.n, syn.n_cigar
    gen_s = time.perf_counter() - t_gen
    log("generated %d records" % n)
    tmpdir = tempfile.mkdtemp(prefix="ipx_bench_", dir=SHM) if extras else None
    sample = min(args.cpu_sample, n)
    sample_bam = None
    if extras:
        sample_bam = os.path.join(tmpdir, "sample.bam")
        syn.write_raw_bam(sample_bam, 0, sample)
    syn.close()

    # the library launches on its own stream; events are recorded on that stream (torch.cuda.Event only sees the
    # stream it is recorded on)
    eng = _native.Engine(device=local, profile=True)
    stream = torch.cuda.ExternalStream(eng.stream(), device=dev)
    eng.set_targets(names)
    devc = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
    torch.cuda.synchronize()
    h2d_bytes = sum(v.numel() * v.element_size() for v in host.values())

    # N > 1: junctions are owned by key range; what a rank does not own is stored straight into its owner's mailbox over
    # NVLink and folded there.  IPX_NO_MAILBOX=1 falls back to packed buffers through NCCL all-to-alls.
    use_mailboxes = world > 1 and os.environ.get("IPX_NO_MAILBOX") is None
    splitters = None
    if use_mailboxes:
        sharded.connect_mailboxes_dist(eng, dev, junction_rows_per_source=1 << 20, hist_rows_per_source=1 << 24)
    if world > 1:
        # load-balanced key ranges from a sample of every rank's junction keys (one untimed pass; real runs do the same
        # after their first batch)
        eng.clear()
        eng.push_device(devc)
        splitters = sharded.balanced_splitters_dist(eng, names, dev)
        eng.clear()

    def exchange():
        return sharded.exchange_dist(eng, names, dev, splitters=splitters)

    def device_step():
        eng.clear()
        eng.push_device(devc)
        if world > 1:
            exchange()
        eng.finalize_device()

    # the end-to-end path sends the compact form of the batch (what a record packer produces directly: operation counts
    # and target runs instead of offsets and a target column); the device expands it
    compact = _native.Engine.compact_columns({k: (v.numpy() if hasattr(v, "numpy") else v) for k, v in host.items()},
                                             alloc=lambda shape, dtype: torch.empty(shape, dtype=to_torch[dtype], pin_memory=True))
    compact_bytes = sum(int(np.asarray(v.numpy() if hasattr(v, "numpy") else v).nbytes) for v in compact.values())

    def e2e_step():
        eng.clear()
        eng.push_compact(compact)
        if world > 1:
            exchange()
            if use_mailboxes:
                sharded.settle_exchange_dist(eng, names, dev, splitters=splitters)
        return eng.flush()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing ----
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()  # samples cover the warm-up steps and the timed region (both run the same load)
        time.sleep(0.3)
    for _ in range(args.warmup):
        device_step()
    barrier()
    eng.reset_timings()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    for a, b in ev:
        a.record(stream)
        device_step()
        b.record(stream)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    step_ms = [a.elapsed_time(b) for a, b in ev]
    tm = eng.timings()
    ms_per_step = sum(step_ms) / len(step_ms)
    log("device-timed: %.3f ms per step" % ms_per_step)

    # ---- what the host link alone costs: the same pinned columns copied to the device, nothing else ----
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        for k, v in host.items():
            devc[k].copy_(v, non_blocking=True)
        torch.cuda.synchronize()
    h2d_floor_ms = 1000.0 * (time.perf_counter() - t0) / 3

    # ---- end to end through the C ABI with host buffers ----
    for _ in range(max(2, args.warmup)):  # also fills the library's pool of pinned result blocks (two alternate below)
        table = e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        table = e2e_step()
    barrier()
    e2e_ms = 1000.0 * (time.perf_counter() - t0) / args.steps
    d2h_bytes = int(table.n) * (4 * 6 + 1 + 8 + 24 + 8 + 8) + int(table.hist_cnt.shape[0]) * 4
    n_junctions = int(table.n)
    n_hist_runs = int(table.n_hist)
    log("end to end: %.2f ms per step" % e2e_ms)

    phases = None
    if world > 1:
        red = torch.tensor([ms_per_step, e2e_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(red, op=dist.ReduceOp.MAX)
        ms_per_step, e2e_ms = float(red[0]), float(red[1])
        tot = torch.tensor([n_junctions], dtype=torch.int64, device=dev)
        dist.all_reduce(tot)
        n_junctions = int(tot[0])
        # where the step goes on every rank (library CUDA events per stage; import includes the device-side wait for the
        # slowest exporter): max over ranks per phase
        ph = torch.tensor([tm["extract_ms"], tm["aggregate_ms"], tm["export_ms"], tm["import_ms"], tm["finalize_ms"], tm["motif_ms"]],
                          dtype=torch.float64, device=dev) / args.steps
        ph_max, ph_min = ph.clone(), ph.clone()
        dist.all_reduce(ph_max, op=dist.ReduceOp.MAX)
        dist.all_reduce(ph_min, op=dist.ReduceOp.MIN)
        keys = ["extract", "aggregate", "export", "wait_and_import", "finalize_and_clear", "resolve"]
        phases = {"max_over_ranks": dict(zip(keys, [float(x) for x in ph_max])), "min_over_ranks": dict(zip(keys, [float(x) for x in ph_min]))}

    # ---- N = 1 extras: reference on the prefix, parity, the other configs, the tools end to end ----
    cpu_line = parity = other = cli = None
    extras_errors = []
    ok = True

    def attempt(what, fn):
        """An extra that fails is reported in the line (and fails the run) instead of taking the headline with it."""
        nonlocal ok
        try:
            return fn()
        except Exception as e:  # noqa: BLE001 -- whatever went wrong is recorded
            import traceback
            traceback.print_exc()
            extras_errors.append("%s: %s: %s" % (what, type(e).__name__, str(e)[-400:]))
            ok = False
            return None

    if extras:
        try:
            del devc, compact
            table.close()
            torch.cuda.empty_cache()
            eng.close()
            eng = None

            def reference_and_parity():
                nonlocal cpu_line, parity, ok
                t_ref, want, _ = run_tool(REF_BIN, os.path.join(tmpdir, "ref_sample"), [sample_bam])
                dec_s = time_decode(sample_bam)
                cpu_line = {"value": sample / t_ref, "unit": UNIT, "cores": 1, "kind": "reference",
                            "sample": "first %d records of the workload as uncompressed BAM in tmpfs; whole CLI run %.2f s of which "
                                      "record decode alone %.2f s" % (sample, t_ref, dec_s),
                            "host_cores_available": os.cpu_count()}
                log("reference on %d records: %.2f s" % (sample, t_ref))
                # the GPU path over the same records through the C ABI
                cut = int(host["cigar_off"][sample])
                part = {k: host[k][:sample] for k in ("tid", "pos", "flag", "nh_class", "xs")}
                part["cigar_off"] = host["cigar_off"][:sample + 1]
                part["cigar"] = host["cigar"][:cut]
                pc = _native.Engine.compact_columns({k: v.numpy() for k, v in part.items()})
                pe = _native.Engine(device=local)  # a fresh handle: sjN numbering continues across the flushes of one
                pe.set_targets(names)
                pe.push_compact(pc)
                t = pe.flush()
                got = {"introns.tsv": t.tsv_header() + t.text_by_target(_native.FMT_INTRON_TSV, 1.0),
                       "juncs.bed": t.text_by_target(_native.FMT_JUNCTION_BED, 1.0),
                       "introns.bed": t.text_by_target(_native.FMT_INTRON_BED9, 1.0)}
                parity = compare_outputs(got, want)
                parity.update({"records": sample, "junctions": int(t.n), "how": "ipx_push_batch_compact + ipx_flush + ipx_format vs the "
                               "files the reference binary wrote for the same records (-c -j -n, -C 1.0)"})
                t.close()
                pe.close()
                ok = ok and parity["identical"]
                log("parity on %d records: %s" % (sample, parity["identical"]))

            attempt("reference + parity", reference_and_parity)
            if os.path.exists(sample_bam):
                os.unlink(sample_bam)
            del host
            other = []
            # the merge's reference run takes ~30 s on one core: start it first, collect it last
            merge = attempt("merge", lambda: merge_workload(tmpdir))
            if merge:
                log("merge: %.1f ms per call" % merge[0]["call_ms_host_rows_in_table_out"])

            def add(what, fn):
                w = attempt(what, fn)
                if w:
                    other.append(w)
                    log("%s: %.3f ms per step" % (what, w["ms_per_step"]))

            add("stranded", lambda: extraction_workload(
                "stranded RF + 3 Gb genome in HBM + canon filter, --min-anchor-length=12, 50 M records (BASELINE configs[2], a quarter of "
                "its records)", 50_000_000, SHORT_STRANDED, tmpdir, torch, dev, peak, 20_000_000, with_genome=True,
                cli_sets=[("canon", ["--strandness=RF", "--min-anchor-length=12", "--genome-fasta=" + os.path.join(tmpdir, "g.fa")]),
                          ("all", ["--strandness=RF", "--min-anchor-length=12", "--sj-filter=all",
                                   "--genome-fasta=" + os.path.join(tmpdir, "g.fa")])],
                strandness=_native.STRAND_RF, min_anchor_length=12, junction_filter=_native.FILTER_CANON))
            add("long reads", lambda: extraction_workload(
                "long reads (~10 kb, ~800 CIGAR operations), 1 M reads, --allow-anchor-indels -C 0.0 (BASELINE configs[3], a fifth of "
                "its reads)", 1_000_000, LONG, tmpdir, torch, dev, peak, 300_000,
                cli_sets=[("indels", ["--allow-anchor-indels", "-C", "0.0"]), ("defaults", [])], max_anchor_indel_size=36))
            add("long reads, defaults", lambda: extraction_workload(
                "the same long reads with default options (no anchor indels)", 1_000_000, LONG, tmpdir, torch, dev, peak, 0, cli_sets=[]))
            cli = attempt("tools end to end", lambda: cli_end_to_end(tmpdir, args.cli_records))
            if cli:
                ok = ok and cli["identical"]
                log("tools end to end: reference %.1f s, GPU tool %.1f s" % (cli["reference_s"], cli["gpu_tool_s"]))
            if merge:
                merge_line, merge_ref, merge_got, merge_err = merge

                def merge_parity():
                    want_m = merge_ref.collect()
                    res = compare_outputs(merge_got, want_m)
                    res["stderr_identical"] = merge_err == merge_ref.stderr
                    res["identical"] = res["identical"] and res["stderr_identical"]
                    res.update({"reference_s": merge_ref.seconds, "how": "GPU intronProspectorMerge vs reference binary on the same 64 "
                                "files: three output files and the strand-conflict lines on stderr"})
                    merge_line["parity"] = [res]
                    merge_line["cpu_baseline"] = {"value": merge_line["rows"] / merge_ref.seconds, "unit": "rows/s", "cores": 1,
                                                  "kind": "reference"}

                attempt("merge parity", merge_parity)
                other.append(merge_line)
            for w in other:
                for p in w.get("parity", []):
                    ok = ok and p["identical"]
        finally:
            shutil.rmtree(tmpdir, ignore_errors=True)

    if rank == 0:
        total_records = n * world
        value = total_records / (ms_per_step / 1000.0)
        steps = args.steps
        # K1 runs once per step: the pipelined kernel over the whole tiles plus a one-block launch for the last partial
        # tile; the roofline line times both together
        roof = roofline_of(n, tm, steps, peak)
        traffic = ncu_traffic()
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int64",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "records_per_gpu": n, "options": OPTIONS, "cigar_words_per_gpu": n_cigar,
                       "live_cigar_words": roof["live_cigar_words"], "junction_events": roof["junction_events"], "junctions": n_junctions,
                       "l2": "inputs (%.1f GB per GPU) larger than the 126 MB L2; no flush needed" % (h2d_bytes / 1e9),
                       "timing": "CUDA events on the launching stream around clear + extract/aggregate + "
                                 + (("export into the owners' mailboxes over NVLink + device-side wait + fold + " if use_mailboxes else
                                      "export + NCCL all-to-all + fold + ") if world > 1 else "") + "device finalize, max over ranks",
                       "generate_s": gen_s},
            "e2e": {"value": total_records / (e2e_ms / 1000.0), "unit": UNIT, "h2d_bytes_per_step": compact_bytes,
                    "d2h_bytes_per_step": d2h_bytes, "ms_per_step": e2e_ms,
                    "h2d_copy_alone_ms": h2d_floor_ms,
                    "note": "ipx_push_batch_compact + ipx_flush from pinned host columns (10 B/record + 4 B/CIGAR word); "
                            "host-link bound: the plain 16 B/record columns alone take h2d_copy_alone_ms to reach the device"},
            "gpu_launches": int(tm["extract_launches"] + tm["aggregate_launches"] + tm["finalize_launches"] + tm["motif_launches"]
                                + tm["exchange_launches"]),
            "roofline": {"bound": "hbm", "kernel": "k_extract_tma (K1: TMA-staged CIGAR walk -> junction events; + k_extract_simple on the last partial tile)",
                         "achieved": roof["achieved"], "peak": peak, "unit": "GB/s", "frac": roof["frac"],
                         "traffic": (traffic["k_extract_dram_bytes_per_record"] * n) if traffic and traffic.get("k_extract_dram_bytes_per_record") else None,
                         "traffic_source": traffic.get("source") if traffic else None,
                         "algorithmic_bytes_per_launch": roof["algorithmic_bytes_per_launch"],
                         "kernel_ms": roof["kernel_ms"], "peak_source": peak_src,
                         "stage_ms_per_step": {"extract": tm["extract_ms"] / steps, "aggregate": tm["aggregate_ms"] / steps,
                                               "finalize_and_clear": tm["finalize_ms"] / steps},
                         "aggregate_kernel": {"kernel": "k_fold<MODE_SINGLE> (K2)", "algorithmic_bytes_per_launch": 32.0 * roof["junction_events"],
                                              "kernel_ms": tm["aggregate_ms"] / steps},
                         "whole_step": {"algorithmic_bytes": roof["algorithmic_bytes_per_launch"] + 32.0 * roof["junction_events"]
                                        + 48.0 * n_junctions / world + 4.0 * n_hist_runs,
                                        "note": "SURVEY 8(d): extract + reduce bytes over the whole device-timed step"}},
            "cpu_baseline": cpu_line,
            "clocks": clocks,
        }
        ws = line["roofline"]["whole_step"]
        ws["achieved"] = ws["algorithmic_bytes"] / (ms_per_step / 1000.0) / 1e9
        ws["frac"] = ws["achieved"] / peak if peak else None
        if phases:
            line["phases_ms"] = phases
        if parity is not None:
            line["parity"] = parity
        if other is not None:
            line["other_workloads"] = other
        if cli is not None:
            line["cli_e2e"] = cli
        if extras_errors:
            line["extras_errors"] = extras_errors
        emit(line)
    if eng is not None:
        eng.close()
    if world > 1:
        dist.destroy_process_group()
    return 0 if ok else 3


if __name__ == "__main__":
    sys.exit(main())
