#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -s -k "conv3d_tc or conv_prep" 2>&1 | tail -15
  timeout 600 python -m pytest tests/test_gpu_engine.py -m gpu -q -x -s -k "conv_prologue_tc" 2>&1 | tail -12
  timeout 300 python scripts/conv_time.py 2>&1 | tail -45 ) > gpurun_out/conv.log 2>&1
cat gpurun_out/conv.log
