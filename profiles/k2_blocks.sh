#!/bin/bash
# K2 occupancy sweep: libipx.so built with -DIPX_K2_BLOCKS=3/4/6 (csrc/build/variants) against the default 5.
B=intronprospector_b200/csrc/build
cp $B/libipx.so $B/libipx_default.so
for v in default k2b3 k2b4 k2b6; do
  if [ $v = default ]; then cp $B/libipx_default.so $B/libipx.so; else cp $B/variants/libipx_$v.so $B/libipx.so; fi
  python bench.py --no-cpu-baseline --steps 5 --warmup 3 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$v', round(d['ms_per_step'],4), {k: round(x,4) for k,x in d['roofline']['stage_ms_per_step'].items()})"
done
cp $B/libipx_default.so $B/libipx.so
