#!/bin/bash
# Round-2 GPU pass (1 GPU, under gpurun).  Stages selected by $STAGES (default: all): tests smoke bench ncu_attn ncu_dec launches
mkdir -p gpurun_out
STAGES=${STAGES:-"tests smoke bench ncu_attn ncu_dec launches"}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
for st in $STAGES; do
  case $st in
    tests)
      for f in ${TEST_FILES:-tests/test_gpu_kernels.py tests/test_gpu_engine.py tests/test_gpu_bench_config.py}; do
        n=$(basename $f .py)
        timeout ${TEST_TIMEOUT:-1500} python -m pytest $f -m gpu -q -x -s ${PYTEST_ARGS} 2>&1 | tail -80 > gpurun_out/$n.log
        echo "== $n"; tail -12 gpurun_out/$n.log
      done ;;
    smoke)
      timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "== smoke rc=$?"; tail -3 gpurun_out/smoke.log ;;
    bench)
      timeout 900 python bench.py --steps ${BENCH_STEPS:-3} --warmup 3 ${BENCH_ARGS} > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "== bench rc=$?"; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err ;;
    ncu_attn)
      # the decode attention kernel at position ~512 of the benchmarked batch (64 rows, groups of 4): skip 256 steps x 24 launches
      timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_grouped -s 6144 -c 2 -f -o gpurun_out/attn_grouped_r2 \
        python bench.py --steps 1 --warmup 0 --ar-steps 258 --no-e2e --no-cpu-baseline --graph off --no-roofline > gpurun_out/p_attn.log 2>&1; echo "== ncu_attn rc=$?" ;;
    ncu_dec)
      timeout 600 ncu --set full --clock-control none --import-source on -k regex:decoder_points_tc -c 1 -f -o gpurun_out/decoder_tc_r2 \
        python bench.py --steps 1 --warmup 0 --ar-steps 4 --no-e2e --no-cpu-baseline --graph off --no-roofline > gpurun_out/p_dec.log 2>&1; echo "== ncu_dec rc=$?" ;;
    launches)
      timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r2.csv \
        python bench.py --steps 1 --warmup 0 --ar-steps 6 --no-e2e --no-cpu-baseline --graph off --no-roofline > gpurun_out/p_launch.log 2>&1; echo "== launches rc=$?" ;;
  esac
done
ls -la gpurun_out | tail -12
