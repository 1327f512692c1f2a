#!/bin/bash
# Round profile pass (1 GPU, under gpurun): ncu captures of the three hot kernels + launch list of a shortened bench step.
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 0 --ar-steps 4 --no-e2e --no-cpu-baseline --graph off"
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_round.csv $B > gpurun_out/p_launch.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:decoder_points_tc -c 1 -o gpurun_out/decoder_tc $B > gpurun_out/p_dec.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:tc_linear_kernel -s 300 -c 4 -o gpurun_out/tc_linear_decode $B > gpurun_out/p_lin.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:tc_linear_kernel -s 8 -c 2 -o gpurun_out/tc_linear_prefill $B > gpurun_out/p_linp.log 2>&1
ls -la gpurun_out | tail -12
