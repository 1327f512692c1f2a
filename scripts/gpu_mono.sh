#!/bin/bash
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "attn" 2>&1 | tail -4
  timeout 600 python -m pytest tests/test_gpu_bench_config.py -m gpu -q -x -s -k "large_batch" 2>&1 | grep -E "passed|failed|bit-exact" | tail -3 ) > gpurun_out/mono.log 2>&1
cat gpurun_out/mono.log
ROWS=512 bash scripts/gpu_bench_rows.sh
