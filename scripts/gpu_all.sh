#!/bin/bash
# full GPU test suite, each file in its own process
mkdir -p gpurun_out
for f in tests/test_gpu_kernels.py tests/test_gpu_engine.py tests/test_gpu_encoder.py tests/test_gpu_mesh.py tests/test_gpu_bench_config.py; do
  n=$(basename $f .py)
  timeout 1500 python -m pytest $f -m gpu -q -x -s ${PYTEST_ARGS} 2>&1 | tail -40 > gpurun_out/$n.log
  echo "== $n"; grep -E "passed|failed|error|max \|d|bit-exact|tc conv" gpurun_out/$n.log | tail -12
done
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "== smoke rc=$?"; tail -2 gpurun_out/smoke.log
