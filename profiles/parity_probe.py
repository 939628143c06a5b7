#!/usr/bin/env python3
"""GPU table against the CPU oracle's at growing sizes (diagnostic: which column differs first, and where)."""
import sys
import os
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from intronprospector_b200 import _native, selftest  # noqa: E402
from intronprospector_b200.synth import SHORT_UNSTRANDED, Synth  # noqa: E402

sizes = [int(x) for x in sys.argv[1:]] or [5_000_000, 20_000_000, 40_000_000]
syn = Synth(max(sizes), kind=SHORT_UNSTRANDED, seed=20261018, model_seed=20261018)
cols = syn.columns()
for n in sizes:
    cut = int(cols["cigar_off"][n])
    part = {k: cols[k][:n] for k in ("tid", "pos", "flag", "nh_class", "xs")}
    part["cigar_off"] = cols["cigar_off"][:n + 1]
    part["cigar"] = cols["cigar"][:cut]
    gpu = _native.Engine(device=0)
    gpu.set_targets(syn.names)
    gpu.push_compact(_native.Engine.compact_columns(part))
    got = gpu.flush()
    ora = selftest.oracle_engine()
    ora.set_targets(syn.names)
    ora.push(part)
    want = ora.flush()
    diffs = selftest.compare_tables(got, want)
    print(n, "junctions", got.n, want.n, "diffs", diffs, flush=True)
    if diffs and got.n == want.n:
        a, b = got.columns(), want.columns()
        for k in ("intron_start", "anchor_start", "anchor_end", "read_counts", "ijunc", "hist_off", "confidence"):
            bad = np.flatnonzero(np.any((a[k] != b[k]).reshape(len(a[k]), -1), axis=1))
            if len(bad):
                i = int(bad[0])
                print("  ", k, "first bad row", i, "of", len(bad), "got", a[k][i], "want", b[k][i], "key", a["chrom"][i] if i < got.n else None,
                      a["intron_start"][i] if i < got.n else None, a["intron_end"][i] if i < got.n else None)
        bad = np.flatnonzero(a["hist_off"] != b["hist_off"])
        if len(bad):
            r = int(bad[0]) - 1
            print("   hist of row", r, "got", a["hist_cnt"][int(a["hist_off"][r]):int(a["hist_off"][r + 1])][:40], "want",
                  b["hist_cnt"][int(b["hist_off"][r]):int(b["hist_off"][r + 1])][:40])
        elif not np.array_equal(a["hist_cnt"], b["hist_cnt"]):
            i = int(np.flatnonzero(a["hist_cnt"] != b["hist_cnt"])[0])
            r = int(np.searchsorted(a["hist_off"], i, side="right") - 1)
            print("   hist_cnt differs at", i, "row", r, "got", a["hist_cnt"][int(a["hist_off"][r]):int(a["hist_off"][r + 1])], "want",
                  b["hist_cnt"][int(b["hist_off"][r]):int(b["hist_off"][r + 1])])
    gpu.close()
    ora.close()
