#!/bin/bash
# where the merge tool's wall clock goes: 32 synthetic tables, phase marks of the tool (IPX_TRACE) and of the library call
cd "$(dirname "$0")/.."
d=$(mktemp -d -p /dev/shm)
python - "$d" <<'PY'
import sys
sys.path.insert(0, "tests")
import merge_synth
merge_synth.write_tables(sys.argv[1], n_samples=32, n_universe=400000, n_contigs=25)
PY
for v in "" 1 dev; do
echo "== ${v:-default (host parser until the device is up, then the device)}${v:+ (1: host parser only, dev: device only)}"
( [ "$v" = 1 ] && export IPX_TSV_HOST=1; [ "$v" = dev ] && export IPX_TSV_DEVICE=1; export IPX_TRACE=1; time intronprospector_b200/csrc/build/intronProspectorMerge --intron-calls=$d/out.tsv --junction-bed=$d/out.bed --intron-bed=$d/out9.bed $d/sample*.tsv 2>&1 | grep -v "different strand" ) 2>&1
done
rm -rf $d
