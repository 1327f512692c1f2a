#!/bin/bash
# Round-2 profile pass at the default bench batch (512 rows per GPU): ncu --set full of the hot kernels + the launch list.
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline --graph off --no-roofline"
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 900 $NCU -k regex:attn_grouped -s 1560 -c 2 -o gpurun_out/r2_attn_grouped $B --ar-steps 70 > gpurun_out/p_attn.log 2>&1; echo "attn rc=$?"
timeout 600 $NCU -k regex:tc_big_linear -s 1200 -c 6 -o gpurun_out/r2_tc_big $B --ar-steps 4 > gpurun_out/p_big.log 2>&1; echo "big rc=$?"
timeout 600 $NCU -k regex:conv3d_tc -c 17 -o gpurun_out/r2_conv3d $B --ar-steps 2 > gpurun_out/p_conv.log 2>&1; echo "conv rc=$?"
timeout 600 $NCU -k regex:decoder_points_tc -c 1 -o gpurun_out/r2_decoder_points $B --ar-steps 2 > gpurun_out/p_dec.log 2>&1; echo "dec rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r2_rows512.csv $B --ar-steps 6 > gpurun_out/p_launch.log 2>&1; echo "launches rc=$?"
ls -la gpurun_out | tail -12
