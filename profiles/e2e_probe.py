import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.getcwd())
from intronprospector_b200 import _native
from intronprospector_b200.synth import SHORT_UNSTRANDED, Synth
syn = Synth(100_000_000, kind=SHORT_UNSTRANDED, seed=20261018, model_seed=20261018)
to_torch = {np.int32: torch.int32, np.uint16: torch.uint16, np.uint8: torch.uint8, np.uint32: torch.uint32}
host = syn.columns(lambda shape, dtype: torch.empty(shape, dtype=to_torch[dtype], pin_memory=True))
names = syn.names; syn.close()
eng = _native.Engine(device=0); eng.set_targets(names)
compact = _native.Engine.compact_columns({k: v.numpy() for k, v in host.items()},
                                         alloc=lambda shape, dtype: torch.empty(shape, dtype=to_torch[dtype], pin_memory=True))
use_compact = os.environ.get("PLAIN") is None
for it in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    eng.clear(); t1 = time.perf_counter()
    if use_compact:
        eng.push_compact(compact)
    else:
        eng.push(host)
    t2 = time.perf_counter()
    torch.cuda.synchronize(); t3 = time.perf_counter()
    import ctypes as C
    tp = C.POINTER(_native.TableStruct)()
    eng._check(eng.lib.ipx_flush(eng.h, C.byref(tp))); t4 = time.perf_counter()
    tab = _native.Table(eng.lib, tp, eng.names, eng.lib.ipx_table_free); t5 = time.perf_counter()
    print("clear %.2f push(enqueue) %.2f wait %.2f flush(C) %.2f Table(py) %.2f total %.2f ms" % tuple(1e3 * x for x in (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4, t5 - t0)), file=sys.stderr)
