"""One process, one synthetic 100 M-record shard on the device, many library configurations (environment switches the
library reads per handle / per launch): per-stage CUDA-event times of the device step for each.
Usage: python profiles/tune_r2.py [records] 'ENV=V,ENV2=V2' 'ENV=V' ...   (an empty string = defaults)"""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.getcwd())
from intronprospector_b200 import _native
if os.environ.get("IPX_LIBIPX"):
    _native.LIBIPX = os.environ["IPX_LIBIPX"]  # a library built with other compile-time constants
from intronprospector_b200.synth import SHORT_UNSTRANDED, Synth

args = sys.argv[1:]
n_rec = int(args.pop(0)) if args and args[0].isdigit() else 100_000_000
configs = args or [""]
dev = torch.device("cuda", 0)
syn = Synth(n_rec, kind=SHORT_UNSTRANDED, seed=20261018, model_seed=20261018)
to_torch = {np.int32: torch.int32, np.uint16: torch.uint16, np.uint8: torch.uint8, np.uint32: torch.uint32}
host = syn.columns(lambda shape, dtype: torch.empty(shape, dtype=to_torch[dtype], pin_memory=True))
names = syn.names
syn.close()
devc = {k: v.to(dev) for k, v in host.items()}
del host
ref = None
for cfg in configs:
    pairs = [kv.split("=", 1) for kv in cfg.split(",") if kv]
    for k, v in pairs:
        os.environ[k] = v
    eng = _native.Engine(device=0, profile=True)
    eng.set_targets(names)
    stream = torch.cuda.ExternalStream(eng.stream(), device=dev)
    def step():
        eng.clear(); eng.push_device(devc); eng.finalize_device()
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    eng.reset_timings()
    steps = 8
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for a, b in ev:
        a.record(stream); step(); b.record(stream)
    torch.cuda.synchronize()
    ms = sum(a.elapsed_time(b) for a, b in ev) / steps
    tm = eng.timings()
    eng.clear(); eng.push_device(devc)
    t = eng.flush()
    sig = (int(t.n), int(t.n_hist), int(np.asarray(t.hist_cnt, dtype=np.uint64).sum()))
    if ref is None:
        ref = sig
    print(json.dumps({"config": cfg or "defaults", "step_ms": round(ms, 4), "extract": round(tm["extract_ms"] / steps, 4),
                      "aggregate": round(tm["aggregate_ms"] / steps, 4), "finalize_clear": round(tm["finalize_ms"] / steps, 4),
                      "same_table": sig == ref, "junctions": sig[0], "runs": sig[1]}), flush=True)
    t.close(); eng.close()
    for k, _ in pairs:
        os.environ.pop(k, None)
