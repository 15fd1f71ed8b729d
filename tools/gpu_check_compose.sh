#!/bin/bash
# First GPU session of the composed path (written after round 1's GPU budget was spent):
#   gpurun --timeout 900 -- 'bash tools/gpu_check_compose.sh'
# GPU parity suite, compute-sanitizer on the composed kernel, a launch list + one full ncu capture of mgc_kernel.
mkdir -p gpurun_out
python -m pytest tests/test_zz_gpu_compose.py -q > gpurun_out/compose_gpu_suite.log 2>&1; echo "compose suite rc=$?" >> gpurun_out/compose_summary.txt
python -m pytest tests/test_zz_gpu_dropin_more.py -q > gpurun_out/dropin_more_gpu_suite.log 2>&1; echo "drop-in (more) suite rc=$?" >> gpurun_out/compose_summary.txt
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_zz_gpu_compose.py -x -q -k "batch_matches or discrete" > gpurun_out/compose_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/compose_summary.txt
timeout 600 compute-sanitizer --tool racecheck python -m pytest tests/test_zz_gpu_compose.py -x -q -k "batch_matches" > gpurun_out/compose_racecheck.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/compose_summary.txt
python bench.py --workload composed --steps 200 --warmup 5 > gpurun_out/bench_composed.json 2> gpurun_out/bench_composed.err; echo "bench composed rc=$?" >> gpurun_out/compose_summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mgc_kernel -c 2 -o gpurun_out/prof_compose python bench.py --workload composed --steps 20 --warmup 2 > gpurun_out/prof_compose.log 2>&1
cat gpurun_out/compose_summary.txt
