#!/bin/bash
# round-end rehearsal: the full GPU suite, smoke, the default bench (both arms), and the 64-row / cfg 2 lines
bash scripts/gpu_all.sh
timeout 1400 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"
python -c "
import json
d=json.load(open('gpurun_out/bench_final.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','breakdown_ms')}); print('e2e', d['e2e']['value']); print('roof', d['roofline']['frac'], d['roofline']['traffic'], 'dec', d['roofline_decoder']['executed_frac'], d['roofline_decoder']['tensor_pipe_pct_ncu']); print('cpu', d['cpu_baseline']['value'], d['cpu_baseline']['spread_max_over_min'], d['clocks'])"; tail -2 gpurun_out/bench_final.err
if [ -n "$EXTRA_LINES" ]; then
  timeout 600 python bench.py --steps 3 --warmup 3 --rows 64 --no-cpu-baseline > gpurun_out/bench_rows64.json 2> gpurun_out/bench_rows64.err; echo "rows64 rc=$?"
  timeout 600 python bench.py --steps 2 --warmup 1 --cfg2 --no-cpu-baseline --cloud-points 0 > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err; echo "cfg2 rc=$?"
  timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference_arm.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
  python -c "
import json
for n in ('rows64','cfg2','reference_arm'):
    d=json.load(open(f'gpurun_out/bench_{n}.json')); print(n, d['value'], d.get('ms_per_step'), (d.get('e2e') or {}).get('value'))"
fi
