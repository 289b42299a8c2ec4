#!/bin/bash
mkdir -p gpurun_out
echo "=== eval tests (fused head)"; timeout 300 python -m pytest tests/test_eval_gpu.py tests/test_properties_gpu.py -q -m gpu -x 2>&1 | tail -8
for v in 1 0; do
  PCSEG_EVAL_CHAIN=$v timeout 300 python bench.py --workload cfg2_eval --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/ev_$v.json 2> gpurun_out/ev_$v.err; echo "chain=$v rc=$?"; tail -2 gpurun_out/ev_$v.err
  python -c "
import json; d=json.load(open('gpurun_out/ev_$v.json')); print('EVAL chain=$v ms/step', d['ms_per_step'], 'Mpts/s', d['value']/1e6, 'e2e', d['e2e']['value']/1e6, 'frac', d['step_frac_of_bf16_sustained'])"
done
