#!/bin/bash
mkdir -p gpurun_out
for w in cfg3_train cfg3_eval cfg5_eval cfg4; do
  timeout 900 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; echo "$w rc=$?"; tail -2 gpurun_out/bench_$w.err
  python - <<PY
import json
try:
    d = json.load(open('gpurun_out/bench_$w.json'))
    print('$w', 'ms/step', round(d['ms_per_step'],3), 'Mpts/s', round(d['value']/1e6,2), 'e2e Mpts/s', round(d['e2e']['value']/1e6,2), 'frac', round(d['step_frac_of_bf16_sustained'],3), 'clk', d['clocks'])
    if d.get('roofline'): print('   roofline', d['roofline']['kernel'], round(d['roofline']['frac'],3))
except Exception as e:
    print('$w failed', e)
PY
done
