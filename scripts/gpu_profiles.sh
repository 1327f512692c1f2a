#!/bin/bash
# Round-2 profile pass at the default bench batch (512 rows per GPU): ncu --set full of the hot kernels, summarised ON THE BOX
# (scripts/ncu_summary.py / ncu_hot.py) so that only small files travel back; + the launch list.
mkdir -p gpurun_out /tmp/ncu
B="python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline --graph off --no-roofline"
NCU="ncu --set full --clock-control none --import-source on -f"
cap() {  # name regex skip count extra-args...
  local name=$1 re=$2 skip=$3 cnt=$4; shift 4
  timeout 900 $NCU -k regex:$re -s $skip -c $cnt -o /tmp/ncu/$name $B "$@" > gpurun_out/p_$name.log 2>&1; echo "$name rc=$?"
  if [ -f /tmp/ncu/$name.ncu-rep ]; then
    python scripts/ncu_summary.py /tmp/ncu/$name.ncu-rep gpurun_out/r2_ncu_$name.json > /dev/null 2>&1
    python scripts/ncu_hot.py /tmp/ncu/$name.ncu-rep 2.0 > gpurun_out/r2_ncu_${name}_hot.txt 2>&1
    ls -la /tmp/ncu/$name.ncu-rep
  fi
}
cap attn_grouped attn_grouped 792 2 --ar-steps 36
cp /tmp/ncu/attn_grouped.ncu-rep gpurun_out/r2_attn_grouped.ncu-rep 2>/dev/null
cap tc_big tc_big_linear 1000 8 --ar-steps 3
cap conv3d conv3d_tc 0 17 --ar-steps 2
cap decoder_points decoder_points_tc 0 1 --ar-steps 2
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file /tmp/ncu/launches.csv $B --ar-steps 2 > gpurun_out/p_launch.log 2>&1; echo "launches rc=$?"
python scripts/summarize_launches.py /tmp/ncu/launches.csv 40 > gpurun_out/r2_launches_rows512.md 2>&1
du -sh gpurun_out
