#!/bin/bash
# profiles/tune2.sh "ENV1=a ENV2=b" "ENV1=c" ... : one bench run per quoted environment setting
mkdir -p gpurun_out
python -m pytest tests/test_parity_gpu.py -m gpu -x -q > gpurun_out/pytest_tune.log 2>&1; tail -1 gpurun_out/pytest_tune.log
i=0
for setting in "$@"; do
  i=$((i+1))
  env $setting python bench.py --no-cpu-baseline --steps 5 --warmup 3 > gpurun_out/tune2_$i.json 2> gpurun_out/tune2_$i.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/tune2_$i.json"))
    print("$setting", "step %.3f ms" % d["ms_per_step"], d["roofline"]["stage_ms_per_step"], "junctions", d["config"]["junctions"])
except Exception as e:
    print("$setting failed", e)
PY
done
