#!/bin/bash
mkdir -p gpurun_out
( for c in 1 0; do for g in 1 0; do SFB200_CHAIN=$c SFB200_ATTN_GROUPED=$g timeout 300 python scripts/step_time.py 64 128 2>&1 | tail -1; done; done
  timeout 300 python scripts/step_time.py 16 128 2>&1 | tail -1
  timeout 300 python scripts/step_time.py 1 128 2>&1 | tail -1
  SFB200_LIB=shapeformer_b200/lib/libsfb200_probe.so timeout 300 python scripts/chain_timeline.py 64 16 2>&1 | tail -40 ) > gpurun_out/probe.log 2>&1
cat gpurun_out/probe.log
