#!/bin/bash
# One bounded GPU session: new host-buffer paths first, then the bench, then the whole GPU suite. Logs -> gpurun_out/.
mkdir -p gpurun_out
date +%s > gpurun_out/t0
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader > gpurun_out/gpu.txt 2>&1
timeout 170 python -m pytest tests/test_gpu_dropin.py -x -q -k "host_rollout or modules or trajectory" > gpurun_out/t_new.log 2>&1
echo "new tests rc=$? t=$(( $(date +%s) - $(cat gpurun_out/t0) ))" | tee -a gpurun_out/summary.txt
tail -3 gpurun_out/t_new.log
timeout 120 python bench.py --steps 1000 --warmup 5 --no-cpu --single-path > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err
echo "bench quick rc=$? t=$(( $(date +%s) - $(cat gpurun_out/t0) ))" | tee -a gpurun_out/summary.txt
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_quick.json").read().strip().splitlines()[-1])
    print("value", d["value"], "e2e", json.dumps(d["e2e"])[:900])
except Exception as ex:
    print("no bench line:", ex)
PY
timeout 240 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
echo "bench default rc=$? t=$(( $(date +%s) - $(cat gpurun_out/t0) ))" | tee -a gpurun_out/summary.txt
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/t_gpu.log 2>&1
echo "gpu suite rc=$? t=$(( $(date +%s) - $(cat gpurun_out/t0) ))" | tee -a gpurun_out/summary.txt
tail -5 gpurun_out/t_gpu.log
