#!/usr/bin/env python3
"""bench.py -- headline benchmark of the junction-extraction hot path (BASELINE.json: aligned reads/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--records R] [--impl ours|reference]

Workload (config.workload): BASELINE.json configs[1], a synthetic coordinate-sorted 2x150 short-read unstranded
RNA-Seq stream of 100 M alignment records over a 25-contig 3 Gb genome (seeded generator, csrc/synth).  One "step" is
one pass of the hot path over that stream: clear the junction table, extract + aggregate every record, finish the
table on the device (key order, sjN ranking, ordered offset histograms).

  value        records/s with the record columns already resident in HBM, device-timed with CUDA events on the
               launching stream (max over ranks), whole job (all ranks' records / slowest rank's time)
  e2e          the same pass through the public C ABI with HOST (pinned) columns: ipx_push_batch (H2D copies inside
               the timed region) + ipx_flush (table D2H + float32 confidences on the host)
  roofline     the extract kernel K1: algorithmic bytes (16 B per record + 4 B per live CIGAR word + 32 B per junction
               event) / its CUDA-event time, against the measured HBM copy bandwidth in MEASURED_PEAKS.json
  cpu_baseline the unmodified reference binary (oracle/_ref/intronProspector, 1 thread) on a bounded prefix of the
               same stream written as an uncompressed BAM, timed on this box's host cores

With N > 1 (torchrun, one process per GPU) every rank extracts its own 100 M-record shard (weak scaling; shards are
consecutive in the global stream, like `intronProspector a.bam b.bam ...`), then partial junction tables are
exchanged by key range with one NCCL all-to-all and folded by the owning rank.

`--impl reference` times the reference's own CPU implementation instead (all host cores, one process per shard).
"""
import argparse
import ctypes as C
import json
import os
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
REF_BIN = os.path.join(REPO, "oracle", "_ref", "intronProspector")
HTS_TOOL = os.path.join(REPO, "oracle", "_ref", "hts_tool")


# Libraries (NCCL, torchrun banners) may print to stdout; the contract is ONE JSON line there.  Everything this process
# prints goes to stderr from here on, and emit() alone writes to the real stdout.
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(line):
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


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
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture, if any."""
    p = os.path.join(REPO, "profiles", "traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get("k_extract_dram_bytes_per_record")
        except ValueError:
            return None
    return None


def time_reference(bam_paths, tmpdir):
    """Run the unmodified reference over every file concurrently (it is single-threaded: one process per file)."""
    procs = []
    t0 = time.perf_counter()
    for i, bam in enumerate(bam_paths):
        out = os.path.join(tmpdir, "ref%d" % i)
        procs.append(subprocess.Popen([REF_BIN, "--intron-calls=" + out + ".tsv", "--junction-bed=" + out + ".juncs.bed",
                                       "--intron-bed=" + out + ".introns.bed", bam], stdout=subprocess.DEVNULL, stderr=subprocess.PIPE))
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
    tmpdir = tempfile.mkdtemp(prefix="ipx_ref_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
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
        for f in os.listdir(tmpdir):
            os.unlink(os.path.join(tmpdir, f))
        os.rmdir(tmpdir)
    total = per * cores
    ms = 1000.0 * sum(times) / len(times)
    value = total / (ms / 1000.0)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": "%d processes x first %d records of a shard, uncompressed BAM in tmpfs" % (cores, per),
                   "decode_only_s_per_shard": decode_s},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "reference",
                         "sample": "%d concurrent single-threaded reference processes, %d records each" % (cores, per)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--records", type=int, default=100_000_000, help="records per GPU (the named workload is 100M)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--ref-sample", type=int, default=4_000_000, help="records per reference process per step")
    ap.add_argument("--cpu-sample", type=int, default=40_000_000, help="records of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)

    import numpy as np
    import torch
    import torch.distributed as dist

    from intronprospector_b200 import _native, sharded
    from intronprospector_b200.synth import SHORT_UNSTRANDED, Synth

    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product has no CPU path)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    # ---- synthetic shard of this rank, in pinned host memory ----
    t_gen = time.perf_counter()
    syn = Synth(args.records, kind=SHORT_UNSTRANDED, seed=20261018 + rank, model_seed=20261018)
    to_torch = {np.int32: torch.int32, np.uint16: torch.uint16, np.uint8: torch.uint8, np.uint32: torch.uint32}
    host = syn.columns(lambda shape, dtype: torch.empty(shape, dtype=to_torch[dtype], pin_memory=True))
    names = syn.names
    n, n_cigar = syn.n, syn.n_cigar
    gen_s = time.perf_counter() - t_gen
    cpu_line = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and os.path.exists(REF_BIN):
        tmpdir = tempfile.mkdtemp(prefix="ipx_cpu_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
        try:
            sample = min(args.cpu_sample, n)
            bam = os.path.join(tmpdir, "sample.bam")
            syn.write_raw_bam(bam, 0, sample)
            ref_s = time_reference([bam], tmpdir)
            dec_s = time_decode(bam)
            cpu_line = {"value": sample / ref_s, "unit": UNIT, "cores": 1, "kind": "reference",
                        "sample": "first %d records of the workload as uncompressed BAM in tmpfs; whole CLI run %.2f s of which "
                                  "record decode alone %.2f s" % (sample, ref_s, dec_s),
                        "host_cores_available": os.cpu_count()}
        finally:
            for f in os.listdir(tmpdir):
                os.unlink(os.path.join(tmpdir, f))
            os.rmdir(tmpdir)
    syn.close()

    # the library launches on its own stream; events are recorded on that stream (torch.cuda.Event only sees the
    # stream it is recorded on)
    eng = _native.Engine(device=local, profile=True)
    stream = torch.cuda.ExternalStream(eng.stream(), device=dev)
    eng.set_targets(names)
    devc = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
    torch.cuda.synchronize()
    h2d_bytes = sum(v.numel() * v.element_size() for v in host.values())

    # N > 1: the partial tables are exchanged by key range (contig-aligned splitters) and folded by their owners.  The
    # exporting kernels store the rows straight into the owners' mailboxes over NVLink; IPX_NO_MAILBOX=1 falls back to
    # packed buffers through NCCL all-to-alls.
    use_mailboxes = world > 1 and os.environ.get("IPX_NO_MAILBOX") is None
    if use_mailboxes:
        sharded.connect_mailboxes_dist(eng, dev, junction_rows_per_source=1 << 20, hist_rows_per_source=1 << 24)

    def exchange():
        return sharded.exchange_dist(eng, names, dev)

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

    if world > 1:
        red = torch.tensor([ms_per_step, e2e_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(red, op=dist.ReduceOp.MAX)
        ms_per_step, e2e_ms = float(red[0]), float(red[1])
        tot = torch.tensor([n_junctions], dtype=torch.int64, device=dev)
        dist.all_reduce(tot)
        n_junctions = int(tot[0])

    if rank == 0:
        total_records = n * world
        value = total_records / (ms_per_step / 1000.0)
        steps = args.steps
        # K1 runs once per step: the pipelined kernel over the whole tiles plus a one-block launch for the last partial
        # tile; the roofline line times both together
        k_ms = tm["extract_ms"] / steps
        live = tm["cigar_live"] / steps
        events = tm["events"] / steps
        alg_bytes = 16.0 * n + 4.0 * live + 32.0 * events
        peak, peak_src = measured_hbm_peak()
        achieved = alg_bytes / (k_ms / 1000.0) / 1e9 if k_ms > 0 else 0.0
        traffic = ncu_traffic()
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int64",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "records_per_gpu": n, "cigar_words_per_gpu": n_cigar, "live_cigar_words": live,
                       "junction_events": events, "junctions": n_junctions, "options": "defaults (-C 1.0), -c -j -n",
                       "l2": "inputs (%.1f GB per GPU) larger than the 126 MB L2; no flush needed" % (h2d_bytes / 1e9),
                       "timing": "CUDA events on the launching stream around clear + extract/aggregate + "
                                 + (("export into the owners' mailboxes over NVLink + barrier + fold + " if use_mailboxes else
                                      "export + NCCL all-to-all + fold + ") if world > 1 else "") + "device finalize, max over ranks",
                       "generate_s": gen_s},
            "e2e": {"value": total_records / (e2e_ms / 1000.0), "unit": UNIT, "h2d_bytes_per_step": compact_bytes,
                    "d2h_bytes_per_step": d2h_bytes, "ms_per_step": e2e_ms,
                    "h2d_copy_alone_ms": h2d_floor_ms,
                    "note": "ipx_push_batch_compact + ipx_flush from pinned host columns (10 B/record + 4 B/CIGAR word); "
                            "host-link bound: the plain 16 B/record columns alone take h2d_copy_alone_ms to reach the device"},
            "gpu_launches": int(tm["extract_launches"] + tm["aggregate_launches"] + tm["finalize_launches"] + tm["motif_launches"]),
            "roofline": {"bound": "hbm", "kernel": "k_extract_tma (K1: TMA-staged CIGAR walk -> junction events; + k_extract_simple on the last partial tile)",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak if peak else None,
                         "traffic": (traffic * n) if traffic else None, "algorithmic_bytes_per_launch": alg_bytes,
                         "kernel_ms": k_ms, "peak_source": peak_src,
                         "stage_ms_per_step": {"extract": tm["extract_ms"] / steps, "aggregate": tm["aggregate_ms"] / steps,
                                               "finalize_and_clear": tm["finalize_ms"] / steps},
                         "aggregate_kernel": {"kernel": "k_fold<MODE_SINGLE> (K2)", "algorithmic_bytes_per_launch": 32.0 * events,
                                              "kernel_ms": tm["aggregate_ms"] / steps}},
            "cpu_baseline": cpu_line,
            "clocks": clocks,
        }
        emit(line)
    eng.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
