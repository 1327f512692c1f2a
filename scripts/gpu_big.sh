#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "linear_big or layernorm" 2>&1 | tail -8
  timeout 300 python scripts/gemm_time.py 256 512 2>&1 | grep -v " tc:" | tail -12
  timeout 900 python -m pytest tests/test_gpu_engine.py -m gpu -q -x 2>&1 | tail -4
  timeout 300 python scripts/step_time.py 64 128 2>&1 | tail -1
  timeout 300 python scripts/step_time.py 256 128 2>&1 | tail -1
  timeout 300 python scripts/step_time.py 512 128 2>&1 | tail -1
  timeout 1200 python -m pytest tests/test_gpu_bench_config.py -m gpu -q -x -s -k "bench_batch" 2>&1 | tail -4 ) > gpurun_out/big.log 2>&1
cat gpurun_out/big.log
