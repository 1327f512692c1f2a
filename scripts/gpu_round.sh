#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "conv or linear_big or attn" 2>&1 | tail -4
  timeout 600 python -m pytest tests/test_gpu_engine.py -m gpu -q -x -s 2>&1 | grep -E "passed|failed|tc conv" | tail -4
  timeout 900 python -m pytest tests/test_gpu_bench_config.py -m gpu -q -x -s -k "large_batch or dropin" 2>&1 | grep -E "passed|failed|bit-exact" | tail -4
  timeout 300 python scripts/conv_time.py 2>&1 | grep -E "^tc|^cudnn|sum conv|blocks" | tail -12 ) > gpurun_out/round.log 2>&1
cat gpurun_out/round.log
ROWS=512 bash scripts/gpu_bench_rows.sh
