#!/bin/bash
# One GPU cycle under gpurun: parity tests, the 100M-read bench, the ncu launch list and one full-set capture of
# the two hot kernels on a shorter run.  Usage: profiles/gpu_cycle.sh <tag> [pytest-args]
tag=${1:-x}; shift
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q "$@" > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_$tag.log
tail -3 gpurun_out/pytest_$tag.log
python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/bench_$tag.json
SMALL="python bench.py --records 20000000 --steps 2 --warmup 1 --no-cpu-baseline"
$SMALL > gpurun_out/plain_$tag.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv $SMALL > gpurun_out/ncu1_$tag.log 2>&1
SMALL1="python bench.py --records 20000000 --steps 1 --warmup 1 --no-cpu-baseline"
$SMALL1 > gpurun_out/plain1_$tag.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_extract_tma|k_extract_long|k_fold' -s 2 -c 2 -f -o gpurun_out/prof_$tag $SMALL1 > gpurun_out/ncu2_$tag.log 2>&1
echo done
# optional extra comparison runs: EXTRA="VAR=1 VAR2=x" profiles/gpu_cycle.sh tag
if [ -n "$EXTRA" ]; then
  env $EXTRA python bench.py --no-cpu-baseline --steps 5 --warmup 3 > gpurun_out/bench_${tag}_extra.json 2> gpurun_out/bench_${tag}_extra.err; echo "extra rc=$?"
  tail -c 1200 gpurun_out/bench_${tag}_extra.json
fi
