#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -4
timeout 600 python bench.py --workload cfg2_eval --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_cfg2_eval.json 2> gpurun_out/bench_cfg2_eval.err; echo "rc=$?"; tail -2 gpurun_out/bench_cfg2_eval.err
python -c "
import json; d=json.load(open('gpurun_out/bench_cfg2_eval.json')); print('EVAL ms/step', d['ms_per_step'], 'pts/s', d['value'], 'e2e', d['e2e']['value'], 'frac', d['step_frac_of_bf16_sustained'])"
