#!/bin/bash
mkdir -p gpurun_out
for r in ${ROWS:-256 512}; do
  timeout 900 python bench.py --steps 2 --warmup 1 --rows $r --no-e2e --no-cpu-baseline ${EXTRA} > gpurun_out/bench_rows$r.json 2> gpurun_out/bench_rows$r.err
  echo "== rows $r rc=$?"; python -c "
import json,sys
d=json.load(open('gpurun_out/bench_rows$r.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','breakdown_ms')}); print(d['roofline'] and {k:d['roofline'][k] for k in ('achieved','frac','avg_launch_us','share_of_ar_pass')}); print(d['roofline_decoder'] and {k:d['roofline_decoder'][k] for k in ('launch_ms','frac','executed_frac')})" 2>&1 | tail -4; tail -3 gpurun_out/bench_rows$r.err
done
