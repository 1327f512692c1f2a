#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "attn" 2>&1 | tail -5
  timeout 300 python scripts/attn_time.py 2>&1 | tail -4
  timeout 900 python -m pytest tests/test_gpu_engine.py -m gpu -q -x 2>&1 | tail -3
  timeout 300 python scripts/step_time.py 64 128 2>&1 | tail -1 ) > gpurun_out/attn.log 2>&1
cat gpurun_out/attn.log
