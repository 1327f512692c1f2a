#!/bin/bash
mkdir -p gpurun_out
( timeout 150 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "attn_prefill" 2>&1 | tail -12 ) > gpurun_out/prefill.log 2>&1
cat gpurun_out/prefill.log
if grep -q "passed" gpurun_out/prefill.log && ! grep -q "failed" gpurun_out/prefill.log; then
  ( timeout 600 python -m pytest tests/test_gpu_engine.py -m gpu -q -x 2>&1 | tail -3
    timeout 900 python -m pytest tests/test_gpu_bench_config.py -m gpu -q -x -s -k "large_batch or bench_batch" 2>&1 | grep -E "passed|failed|bit-exact" | tail -4
    for v in 1 0; do SFB200_PREFILL_TC=$v timeout 300 python scripts/step_time.py 512 8 2>&1 | tail -1; done ) >> gpurun_out/prefill.log 2>&1
  tail -8 gpurun_out/prefill.log
fi
