#!/bin/bash
# Development GPU run (under gpurun): unit-kernel tests, then engine tests, each file in its own process so a sticky CUDA
# error in one does not hide the others.  Logs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
for f in tests/test_gpu_kernels.py tests/test_gpu_engine.py; do
  n=$(basename $f .py)
  CUDA_LAUNCH_BLOCKING=${BLOCKING:-0} timeout 900 python -m pytest $f -m gpu -q -x ${PYTEST_ARGS} 2>&1 | tail -60 > gpurun_out/$n.log
  echo "== $n"; tail -15 gpurun_out/$n.log
done
