#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mg_rollout_kernel -s 1 -c 1 -f -o gpurun_out/prof_gen_ring \
  python bench.py --workload generator --steps 24 --warmup 3 --no-cpu --single-path --preheat 0 > gpurun_out/prof_gen_ring.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/prof_gen_ring.log; ls -la gpurun_out/*.ncu-rep
