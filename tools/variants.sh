#!/bin/bash
# Build kernel variants ON the GPU box and time them (eager single-step launches and the persistent rollout).
# usage: tools/variants.sh "name1:-DFLAG=.. -DFLAG2=.." "name2:..."
mkdir -p gpurun_out
for spec in "$@"; do
  name="${spec%%:*}"; flags="${spec#*:}"
  PYMGRID_B200_NVCC_EXTRA="$flags" python -c "from pymgrid_b200 import build; build.build(force=True)" > gpurun_out/build_$name.log 2>&1
  for p in graph rollout; do
    timeout 300 python bench.py --path $p --steps 200 --warmup 10 --no-cpu 2> gpurun_out/var_${name}_$p.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$name', '$p', round(d['ms_per_step']*1e3,2), 'us/step', round(d['value']/1e9,3), 'G/s', 'frac', round(d['roofline']['frac'],3))" 
  done
done
