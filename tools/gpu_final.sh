#!/bin/bash
# Final bounded validation: whole GPU suite, smoke, default bench, generator bench, then sanitizer passes on the new paths.
mkdir -p gpurun_out
date +%s > gpurun_out/t0
el() { echo $(( $(date +%s) - $(cat gpurun_out/t0) )); }
timeout 240 python -m pytest tests -m gpu -x -q > gpurun_out/final_gpu_suite.log 2>&1
echo "gpu suite rc=$? t=$(el)" | tee gpurun_out/final_summary.txt; tail -2 gpurun_out/final_gpu_suite.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1
echo "smoke rc=$? t=$(el)" | tee -a gpurun_out/final_summary.txt; tail -1 gpurun_out/final_smoke.log
timeout 240 python bench.py > gpurun_out/final_bench_default.json 2> gpurun_out/final_bench_default.err
echo "bench default rc=$? t=$(el)" | tee -a gpurun_out/final_summary.txt
timeout 120 python bench.py --workload generator --steps 500 --warmup 5 --no-cpu > gpurun_out/final_bench_generator.json 2> gpurun_out/final_bench_generator.err
echo "bench generator rc=$? t=$(el)" | tee -a gpurun_out/final_summary.txt
timeout 100 compute-sanitizer --tool racecheck --print-limit 5 python -m pytest tests/test_gpu_parity.py -x -q -k "sliding_windows" > gpurun_out/final_racecheck_ring.log 2>&1
echo "racecheck ring rc=$? t=$(el)" | tee -a gpurun_out/final_summary.txt; grep -E "RACECHECK SUMMARY|passed|failed|ERROR SUMMARY" gpurun_out/final_racecheck_ring.log | tail -3
timeout 80 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_dropin.py -x -q -k "host_rollout and native" > gpurun_out/final_memcheck_host_rollout.log 2>&1
echo "memcheck host rollout rc=$? t=$(el)" | tee -a gpurun_out/final_summary.txt; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/final_memcheck_host_rollout.log | tail -3
