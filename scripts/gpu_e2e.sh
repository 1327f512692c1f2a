#!/bin/bash
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests/test_gpu_bench_config.py -m gpu -q -x -s -k "large_batch or dropin" 2>&1 | tail -8 ) > gpurun_out/e2e_tests.log 2>&1
cat gpurun_out/e2e_tests.log
timeout 1400 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_r512.json 2> gpurun_out/bench_r512.err; echo "bench rc=$?"
python -c "
import json
d=json.load(open('gpurun_out/bench_r512.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e'])" ; tail -2 gpurun_out/bench_r512.err
