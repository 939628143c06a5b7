import os, sys, time, json
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.getcwd())
from intronprospector_b200 import _native, sharded
from intronprospector_b200.synth import SHORT_UNSTRANDED, Synth
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
syn = Synth(100_000_000, kind=SHORT_UNSTRANDED, seed=20261018 + rank, model_seed=20261018)
to_torch = {np.int32: torch.int32, np.uint16: torch.uint16, np.uint8: torch.uint8, np.uint32: torch.uint32}
host = syn.columns(lambda shape, dtype: torch.empty(shape, dtype=to_torch[dtype], pin_memory=True))
names = syn.names; syn.close()
eng = _native.Engine(device=local, profile=True); eng.set_targets(names)
devc = {k: v.to(dev) for k, v in host.items()}
def T():
    torch.cuda.synchronize(); return time.perf_counter()
acc = {}
def add(k, dt): acc[k] = acc.get(k, 0.0) + dt
for it in range(8):
    dist.barrier(); t0 = T()
    eng.clear(); eng.push_device(devc); t1 = T()
    (sj, cj), (sh, ch) = sharded.export_partials(eng, names, world, dev); t2 = T()
    nj, nh = sharded.exchange_counts([cj, ch], dev); rj = sharded.all_to_all_bytes(sj, cj, nj); rh = sharded.all_to_all_bytes(sh, ch, nh); t3 = T()
    sharded.import_partials(eng, rj, nj, rh, nh); t4 = T()
    eng.finalize_device(); t5 = T()
    if it >= 3:
        add("local K1+K2", t1 - t0); add("export", t2 - t1); add("all_to_all x2", t3 - t2); add("import", t4 - t3); add("finalize", t5 - t4)
if rank == 0:
    print(json.dumps({k: round(1e3 * v / 5, 3) for k, v in acc.items()}), "junction bytes per dest", cj, "offset-row bytes per dest", ch)
dist.destroy_process_group()
