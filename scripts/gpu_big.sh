#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "linear_big or layernorm" 2>&1 | tail -8
  timeout 300 python scripts/gemm_time.py 2>&1 | tail -30 ) > gpurun_out/big.log 2>&1
cat gpurun_out/big.log
