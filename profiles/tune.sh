#!/bin/bash
# Runs the bench once per setting of an environment variable and prints the stage times.
# Usage: profiles/tune.sh VAR v1 v2 ...    (first the parity tests with the default)
var=$1; shift
mkdir -p gpurun_out
python -m pytest tests/test_parity_gpu.py -m gpu -x -q > gpurun_out/pytest_tune.log 2>&1; tail -1 gpurun_out/pytest_tune.log
for v in "$@"; do
  env $var=$v python bench.py --no-cpu-baseline --steps 5 --warmup 3 > gpurun_out/tune_${var}_$v.json 2> gpurun_out/tune_${var}_$v.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/tune_${var}_$v.json"))
    print("$var=$v", "step %.3f ms" % d["ms_per_step"], d["roofline"]["stage_ms_per_step"], "junctions", d["config"]["junctions"])
except Exception as e:
    print("$var=$v failed", e)
PY
done
