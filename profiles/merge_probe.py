"""The merge fold (K4) alone, for ncu: a few synthetic tables parsed by the host, folded twice through the C ABI.
Usage: [ncu ... -k regex:k_merge_fold] python profiles/merge_probe.py [tables] [universe]"""
import ctypes as C, os, sys, tempfile, shutil
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import merge_synth
from intronprospector_b200 import _native
n_tables = int(sys.argv[1]) if len(sys.argv) > 1 else 8
universe = int(sys.argv[2]) if len(sys.argv) > 2 else 400_000
d = tempfile.mkdtemp(dir="/dev/shm")
paths, _ = merge_synth.write_tables(d, n_tables, universe, n_contigs=25)
io = _native.load_io_library()
arr = (C.c_char_p * len(paths))(*[p.encode() for p in paths])
calls = C.c_void_p()
assert io.ipxio_calls_load(arr, len(paths), _native.FILTER_NULL, C.byref(calls)) == 0
rows = io.ipxio_calls_rows(calls)
names = [io.ipxio_calls_chroms(calls)[i].decode() for i in range(io.ipxio_calls_n_chroms(calls))]
eng = _native.Engine(device=0, profile=True)
eng.set_targets(names)
for _ in range(2):
    t, _log = eng.merge_rows(rows)
    n = t.n
    t.close()
tm = eng.timings()
print("rows", int(rows.contents.n_rows), "introns", n, "merge_ms per call", tm["merge_ms"] / 2)
eng.close(); io.ipxio_calls_free(calls); shutil.rmtree(d)
