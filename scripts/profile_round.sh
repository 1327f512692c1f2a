#!/bin/bash
# Round profile pass (1 GPU, under gpurun): launch list of a shortened bench step + ncu captures of the hot kernels.
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 0 --ar-steps 6 --no-e2e --no-cpu-baseline --graph off"
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_final.csv $B > gpurun_out/p_launch.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:decoder_points_tc -c 1 -o gpurun_out/decoder_tc_final $B > gpurun_out/p_dec.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc_linear_ps -s 100 -c 4 -o gpurun_out/tc_linear_ps_final $B > gpurun_out/p_lin.log 2>&1
ls -la gpurun_out | tail -8
