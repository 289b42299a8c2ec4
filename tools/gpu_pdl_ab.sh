#!/bin/bash
mkdir -p gpurun_out
for v in 1 0; do
  for w in cfg2 cfg2_eval; do
    PCSEG_PDL=$v timeout 300 python bench.py --workload $w --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/pdl_${v}_$w.json 2> gpurun_out/pdl.err || tail -3 gpurun_out/pdl.err
    python -c "
import json; d=json.load(open('gpurun_out/pdl_${v}_$w.json')); print('PDL=$v $w ms/step', round(d['ms_per_step'],4), 'Mpts/s', round(d['value']/1e6,2), 'e2e', round(d['e2e']['value']/1e6,2))"
  done
done
