#!/bin/bash
# One gpurun session: everything lands in gpurun_out/<tag>_*.  Usage: tools/gpu_session.sh <tag> <stage...>
tag=$1; shift
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/${tag}_gpu.txt 2>&1
for stage in "$@"; do
  case $stage in
    microbench)
      timeout 120 ./tools/microbench_store 131072 200 > $out/${tag}_microbench.txt 2>&1
      timeout 120 ./tools/microbench_store 65536 400 > $out/${tag}_microbench_65536.txt 2>&1 ;;
    composed)
      timeout 300 python bench.py --workload composed --steps 512 --warmup 5 > $out/${tag}_bench_composed.json 2> $out/${tag}_bench_composed.err
      timeout 300 ncu --set full --clock-control none --import-source on -k regex:mgc_kernel -s 2 -c 1 -o $out/${tag}_mgc \
        python bench.py --workload composed --steps 8 --warmup 3 --no-cpu > $out/${tag}_ncu_composed.log 2>&1 ;;
    tests)
      timeout 900 python -m pytest tests -m gpu -x -q > $out/${tag}_gpu_tests.log 2>&1 ;;
    bench)
      timeout 600 python bench.py > $out/${tag}_bench_default.json 2> $out/${tag}_bench_default.err ;;
    bench256)
      timeout 600 python bench.py --steps 256 --warmup 5 --no-cpu --no-configs > $out/${tag}_bench_steps256.json 2> $out/${tag}_bench_steps256.err ;;
    bench20)
      timeout 600 python bench.py --steps 20 --warmup 3 > $out/${tag}_bench_steps20.json 2> $out/${tag}_bench_steps20.err ;;
    generator)
      timeout 300 python bench.py --workload generator --steps 500 --warmup 5 --single-path --no-cpu > $out/${tag}_bench_generator.json 2> $out/${tag}_bench_generator.err ;;
    ragged)
      timeout 300 python bench.py --ragged --steps 500 --warmup 5 --single-path --no-cpu > $out/${tag}_bench_ragged.json 2> $out/${tag}_bench_ragged.err ;;
    emit_tests)
      timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --maxfail=6 -k "row_emitters or rollout_kernel_equals or generator_grids_rollout or vectorised_generator or automatic_emitter or set_trajectories or trajectory_windows or overlapped_launches" > $out/${tag}_emit_tests.log 2>&1 ;;
    tune)
      timeout 1200 python tools/tune_emitters.py --steps 400 --step-path > $out/${tag}_tune.jsonl 2> $out/${tag}_tune.err ;;
    tune_roles)
      timeout 900 python tools/tune_emitters.py --steps 400 --variants image_ws --workloads pymgrid25,ragged,generator > $out/${tag}_tune.jsonl 2> $out/${tag}_tune.err ;;
    tune_composed)
      timeout 600 python tools/tune_composed.py > $out/${tag}_tune_composed.txt 2> $out/${tag}_tune_composed.err ;;
    compose_tests)
      timeout 900 python -m pytest tests/test_zz_gpu_compose.py -x -q -m gpu > $out/${tag}_compose_tests.log 2>&1 ;;
    tune_store_hint)      # same-box ABAB of the row-store cache hint (default .L1::no_allocate vs st.global.cs, -DMG_STORE_CS)
      cp pymgrid_b200/_lib/libpymgrid_b200.so /tmp/lib_na.so
      PYMGRID_B200_NVCC_EXTRA=-DMG_STORE_CS timeout 300 python -m pymgrid_b200.build > $out/${tag}_build_cs.log 2>&1
      cp pymgrid_b200/_lib/libpymgrid_b200.so /tmp/lib_cs.so
      for round in 1 2 3; do for v in cs na; do
        cp /tmp/lib_$v.so pymgrid_b200/_lib/libpymgrid_b200.so
        timeout 300 python tools/tune_emitters.py --steps 400 --variants lsu_ws --workloads pymgrid25,discrete > $out/${tag}_tune_${v}_$round.jsonl 2> $out/${tag}_tune_${v}_$round.err
      done; done
      cp /tmp/lib_na.so pymgrid_b200/_lib/libpymgrid_b200.so ;;
    tune_const)
      MG_DEBUG_CONST_ACTIONS=1 timeout 900 python tools/tune_emitters.py --steps 400 --variants default --workloads pymgrid25,ragged,generator > $out/${tag}_tune_const.jsonl 2> $out/${tag}_tune_const.err ;;
    tune_gen)
      timeout 900 python tools/tune_emitters.py --steps 400 --variants image_ws --workloads generator > $out/${tag}_tune.jsonl 2> $out/${tag}_tune.err ;;
    tune_steps)
      timeout 900 python tools/tune_emitters.py --steps 256 --variants lsu_ws --step-path --workloads pymgrid25,ragged,replicas,discrete,generator > $out/${tag}_tune.jsonl 2> $out/${tag}_tune.err ;;
    tune_default)
      timeout 900 python tools/tune_emitters.py --steps 400 --variants default > $out/${tag}_tune.jsonl 2> $out/${tag}_tune.err ;;
    ncu_gen)
      timeout 600 ncu --set full --clock-control none --import-source on -k regex:mg_rollout -s 1 -c 1 -o $out/${tag}_gen \
        python bench.py --workload generator --steps 24 --warmup 3 --preheat 0 --single-path --no-cpu $BENCH_EXTRA > $out/${tag}_ncu_gen.log 2>&1 ;;
    ncu_gen_src)
      timeout 600 ncu --section SourceCounters --section InstructionStats --section WarpStateStats --section LaunchStats --section Occupancy --clock-control none --import-source on -k regex:mg_rollout -s 1 -c 1 -o $out/${tag}_gen \
        python bench.py --workload generator --steps 24 --warmup 3 --preheat 0 --single-path --no-cpu $BENCH_EXTRA > $out/${tag}_ncu_gen.log 2>&1 ;;
    ncu_ragged)
      timeout 600 ncu --set full --clock-control none --import-source on -k regex:mg_rollout -s 1 -c 1 -o $out/${tag}_ragged \
        python bench.py --ragged --steps 64 --warmup 3 --preheat 0 --single-path --no-cpu $BENCH_EXTRA > $out/${tag}_ncu_ragged.log 2>&1 ;;
    launch_list)
      timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches.csv \
        python bench.py --steps 20 --warmup 3 --preheat 0 --no-cpu --min-timed-ms 1 > $out/${tag}_launch_list_bench.log 2>&1 ;;
    sanitizer)
      timeout 600 compute-sanitizer --tool memcheck python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_memcheck_smoke.log 2>&1
      timeout 900 compute-sanitizer --tool racecheck python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_racecheck_smoke.log 2>&1
      timeout 900 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "generator_grids_rollout or automatic_emitter" > $out/${tag}_racecheck_generator.log 2>&1 ;;
    bench_discrete)
      timeout 300 python bench.py --workload discrete --steps 500 --warmup 5 --no-cpu > $out/${tag}_bench_discrete.json 2> $out/${tag}_bench_discrete.err ;;
    final_benches)
      timeout 600 python bench.py > $out/${tag}_bench_default.json 2> $out/${tag}_bench_default.err
      timeout 600 python bench.py --steps 20 --warmup 5 > $out/${tag}_bench_steps20.json 2> $out/${tag}_bench_steps20.err
      timeout 300 python bench.py --workload generator --steps 500 --warmup 5 --no-cpu > $out/${tag}_bench_generator.json 2> $out/${tag}_bench_generator.err
      timeout 300 python bench.py --ragged --steps 500 --warmup 5 --no-cpu --no-configs > $out/${tag}_bench_ragged.json 2> $out/${tag}_bench_ragged.err
      timeout 300 python bench.py --workload discrete --steps 500 --warmup 5 --no-cpu > $out/${tag}_bench_discrete.json 2> $out/${tag}_bench_discrete.err
      timeout 300 python bench.py --workload replicas --steps 500 --warmup 5 --no-cpu > $out/${tag}_bench_replicas.json 2> $out/${tag}_bench_replicas.err
      timeout 300 python bench.py --impl reference --steps 5 --warmup 3 > $out/${tag}_bench_reference.json 2> $out/${tag}_bench_reference.err ;;
    smoke)
      timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1 ;;
    ncu_default)
      timeout 600 ncu --set full --clock-control none --import-source on -k regex:mg_rollout -s 1 -c 1 -o $out/${tag}_default \
        python bench.py --steps 64 --warmup 3 --preheat 0 --single-path --no-cpu $BENCH_EXTRA > $out/${tag}_ncu_default.log 2>&1 ;;
    scale*)
      # scale2 / scale4 / scale8 / scale1248: the default bench (headline + configs block incl. the config-5 shard) on N GPUs of this box
      for n in $(echo ${stage#scale} | sed 's/1248/1 2 4 8/'); do
        if [ "$n" = "1" ]; then
          timeout 600 python bench.py --gpus 1 --steps 256 --warmup 5 --no-cpu > $out/${tag}_scale_n1.json 2> $out/${tag}_scale_n1.err
        else
          timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) \
            bench.py --gpus $n --steps 256 --warmup 5 > $out/${tag}_scale_n$n.json 2> $out/${tag}_scale_n$n.err
        fi
      done
      nvidia-smi topo -m > $out/${tag}_topo.txt 2>&1 ;;
    *) echo "unknown stage $stage" ;;
  esac
done
ls -la $out | tail -30
