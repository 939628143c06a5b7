#!/bin/bash
# a merge of synthetic tables through the GPU tool with phase marks, under a short timeout (hang hunting)
cd "$(dirname "$0")/.."
d=$(mktemp -d)
python - "$d" <<'PY'
import sys
sys.path.insert(0, "tests")
import merge_synth
merge_synth.write_tables(sys.argv[1], n_samples=24, n_universe=30000)
PY
for t in "" 1; do
echo "== upload threads: ${t:-default}"
IPX_MERGE_UPLOAD_THREADS=$t IPX_TRACE=1 timeout 20 intronprospector_b200/csrc/build/intronProspectorMerge --intron-calls=$d/out.tsv $d/sample*.tsv; echo "rc=$?"
done
rm -rf $d
