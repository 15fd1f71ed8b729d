#!/bin/bash
# Build kernel variants ON the GPU box and time them.  usage: tools/variants.sh "name1:-DFLAG=.." "name2:..."
mkdir -p gpurun_out
for spec in "$@"; do
  name="${spec%%:*}"; flags="${spec#*:}"
  PYMGRID_B200_NVCC_EXTRA="$flags" python -c "from pymgrid_b200 import build; build.build(force=True)" > gpurun_out/build_$name.log 2>&1
  timeout 300 python bench.py --no-cpu ${BENCH_ARGS} 2> gpurun_out/var_${name}.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$name', d['config']['path'], round(d['ms_per_step']*1e3,2), 'us/step', round(d['value']/1e9,3), 'G/s', {k: round(v['us_per_step'],2) for k,v in d['other_paths'].items()})"
done
