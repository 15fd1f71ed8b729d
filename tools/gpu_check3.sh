#!/bin/bash
mkdir -p gpurun_out
date +%s > gpurun_out/t0
timeout 200 python -m pytest tests/test_gpu_parity.py -x -q -k "generator" > gpurun_out/t_new3.log 2>&1
echo "generator tests rc=$? t=$(( $(date +%s) - $(cat gpurun_out/t0) ))" | tee gpurun_out/summary3.txt
tail -4 gpurun_out/t_new3.log
timeout 150 python bench.py --workload generator --steps 300 --warmup 5 --no-cpu --single-path > gpurun_out/bench_gen_v3.json 2> gpurun_out/bench_gen_v3.err
echo "bench generator rc=$? t=$(( $(date +%s) - $(cat gpurun_out/t0) ))" | tee -a gpurun_out/summary3.txt
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_gen_v3.json").read().strip().splitlines()[-1])
    print("ring v3: value %.4g" % d["value"], "us/step %.2f" % (1e3 * d["ms_per_step"]), "frac %.3f" % d["roofline"]["frac"], "e2e %.4g" % d["e2e"]["value"])
except Exception as ex:
    print("no bench line:", ex)
PY
timeout 150 ncu --section SourceCounters --section SpeedOfLight --metrics smsp__inst_executed.sum,gpu__time_duration.sum --import-source on --clock-control none -k regex:mg_rollout_kernel -s 1 -c 1 -f -o gpurun_out/prof_gen_ring_v3 \
  python bench.py --workload generator --steps 24 --warmup 3 --no-cpu --single-path --preheat 0 > gpurun_out/prof_gen_ring_v3.log 2>&1
echo "ncu rc=$?"
