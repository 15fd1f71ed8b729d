#!/bin/bash
mkdir -p gpurun_out
date +%s > gpurun_out/t0
timeout 200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_dropin.py -x -q -k "generator or rule_based" > gpurun_out/t_new2.log 2>&1
echo "new tests rc=$? t=$(( $(date +%s) - $(cat gpurun_out/t0) ))" | tee gpurun_out/summary2.txt
tail -15 gpurun_out/t_new2.log
for flag in "" "--no-ring"; do
  timeout 150 python bench.py --workload generator --steps 300 --warmup 5 --no-cpu --single-path $flag > gpurun_out/bench_gen${flag}.json 2> gpurun_out/bench_gen${flag}.err
  echo "bench generator $flag rc=$? t=$(( $(date +%s) - $(cat gpurun_out/t0) ))" | tee -a gpurun_out/summary2.txt
  python - "$flag" <<'PY'
import json, sys
f = "gpurun_out/bench_gen%s.json" % sys.argv[1]
try:
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(sys.argv[1] or "ring", "value %.4g" % d["value"], "us/step %.2f" % (1e3 * d["ms_per_step"]), "frac %.3f" % d["roofline"]["frac"], "e2e %.4g" % d["e2e"]["value"])
except Exception as ex:
    print("no bench line:", ex)
PY
done
