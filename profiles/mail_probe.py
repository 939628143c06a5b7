"""Cost of the mailbox export / import kernels on one GPU: a single handle whose mailbox has one slot (itself), so
every row of the 100 M-read table goes through export -> slot -> import.  Phases are timed with events on the
library's stream.  Usage: python profiles/mail_probe.py"""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.getcwd())
from intronprospector_b200 import _native, sharded
from intronprospector_b200.synth import SHORT_UNSTRANDED, Synth

dev = torch.device("cuda", 0)
syn = Synth(int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000, kind=SHORT_UNSTRANDED, seed=20261018, model_seed=20261018)
to_torch = {np.int32: torch.int32, np.uint16: torch.uint16, np.uint8: torch.uint8, np.uint32: torch.uint32}
host = syn.columns(lambda shape, dtype: torch.empty(shape, dtype=to_torch[dtype], pin_memory=True))
names = syn.names
syn.close()
eng = _native.Engine(device=0)
eng.set_targets(names)
devc = {k: v.to(dev) for k, v in host.items()}
sharded.connect_mailboxes_local([eng], 1 << 20, 1 << 24)
stream = torch.cuda.ExternalStream(eng.stream(), device=dev)
acc = {}
for it in range(8):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
    eng.clear()
    ev[0].record(stream)
    eng.push_device(devc)
    ev[1].record(stream)
    dbg = os.environ.get("PROBE_SYNC") is not None
    if dbg: torch.cuda.synchronize(); print("pushed", flush=True)
    eng.partial_export_mail([], [])
    if dbg: torch.cuda.synchronize(); print("exported", flush=True)
    ev[2].record(stream)
    eng.partial_import_mail()
    if dbg: torch.cuda.synchronize(); print("imported", flush=True)
    ev[3].record(stream)
    eng.finalize_device()
    ev[4].record(stream)
    torch.cuda.synchronize()
    if dbg: print("finalized", flush=True)
    if it >= 3:
        for k, a, b in (("K1+K2", 0, 1), ("export_mail (+clear)", 1, 2), ("import_mail", 2, 3), ("finalize", 3, 4)):
            acc[k] = acc.get(k, 0.0) + ev[a].elapsed_time(ev[b]) / 5
t = eng.flush()
print(json.dumps({k: round(v, 4) for k, v in acc.items()}), "junctions", t.n)
