#!/bin/bash
# One GPU cycle under gpurun: parity tests, the 100M-read bench, the ncu launch list and one full-set capture of
# the hot kernels on a shorter run.  Usage: profiles/gpu_cycle.sh <tag> [pytest-args]
tag=${1:-x}; shift
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout=300 "$@" > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_$tag.log
tail -15 gpurun_out/pytest_$tag.log
if [ -z "$NO_BENCH" ]; then
timeout 600 python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/bench_$tag.json; tail -5 gpurun_out/bench_$tag.err
fi
if [ -z "$NO_NCU" ]; then
SMALL="python bench.py --records 20000000 --steps 2 --warmup 1 --no-cpu-baseline --quick"
timeout 300 $SMALL > gpurun_out/plain_$tag.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv $SMALL > gpurun_out/ncu1_$tag.log 2>&1
SMALL1="python bench.py --records 20000000 --steps 1 --warmup 1 --no-cpu-baseline --quick"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_extract_tma|k_fold|k_fin_small|k_fin_write' -s 4 -c 4 -f -o gpurun_out/prof_$tag $SMALL1 > gpurun_out/ncu2_$tag.log 2>&1
fi
echo done
