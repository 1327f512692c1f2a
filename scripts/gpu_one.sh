#!/bin/bash
# run one pytest selection under a hard timeout (tcgen05 bring-up: a wrong barrier hangs the kernel)
mkdir -p gpurun_out
timeout ${T:-180} python -m pytest "$@" -q -x 2>&1 | tail -40 > gpurun_out/one.log
echo "rc=$?" >> gpurun_out/one.log
cat gpurun_out/one.log
