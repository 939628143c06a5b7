for v in 0 1 2 3 4; do
  IPX_K2_DBG=$v python bench.py --no-cpu-baseline --steps 5 --warmup 3 > gpurun_out/dbg_$v.json 2>/dev/null
  python -c "
import json; d=json.load(open('gpurun_out/dbg_$v.json')); print('dbg $v', d['roofline']['stage_ms_per_step'])"
done
