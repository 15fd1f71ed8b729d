#!/bin/bash
# Last bounded GPU session of round 1 (~100 s of box time left): the new fuzz / overfull-battery / soc parity tests and the
# rest of the GPU suite side by side, smoke(), then the default bench.  Logs -> gpurun_out/.
mkdir -p gpurun_out
date +%s > gpurun_out/t0
el() { echo $(( $(date +%s) - $(cat gpurun_out/t0) )); }
(timeout 75 python -m pytest tests/test_gpu_parity.py tests/test_gpu_dropin.py -q --tb=short -p no:cacheprovider -k "fuzz or overfull or soc_before" > gpurun_out/t_fuzz.log 2>&1; echo "fuzz rc=$? t=$(el)" >> gpurun_out/summary_last.txt) &
(timeout 75 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -k "not fuzz and not overfull and not soc_before" > gpurun_out/t_rest.log 2>&1; echo "rest rc=$? t=$(el)" >> gpurun_out/summary_last.txt) &
(timeout 60 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$? t=$(el)" >> gpurun_out/summary_last.txt) &
wait
tail -n 4 gpurun_out/t_fuzz.log gpurun_out/t_rest.log gpurun_out/smoke.log
timeout 40 python bench.py > gpurun_out/bench_last.json 2> gpurun_out/bench_last.err
echo "bench rc=$? t=$(el)" >> gpurun_out/summary_last.txt
cat gpurun_out/summary_last.txt
