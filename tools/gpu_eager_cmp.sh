#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_cfg2_full.json 2> gpurun_out/bench_cfg2_full.err; echo "rc=$?"; tail -3 gpurun_out/bench_cfg2_full.err
python -c "
import json; d=json.load(open('gpurun_out/bench_cfg2_full.json')); print('TRAIN', d['value'], 'e2e', d['e2e']['value'], 'cpu', d['cpu_baseline']['value'], 'eager', d['torch_eager_same_gpu'])"
timeout 900 python bench.py --workload cfg2_eval --steps 20 --warmup 5 > gpurun_out/bench_cfg2_eval_full.json 2> gpurun_out/bench_cfg2_eval_full.err; echo "rc=$?"; tail -3 gpurun_out/bench_cfg2_eval_full.err
python -c "
import json; d=json.load(open('gpurun_out/bench_cfg2_eval_full.json')); print('EVAL', d['value'], 'e2e', d['e2e']['value'], 'cpu', d['cpu_baseline']['value'], 'eager', d['torch_eager_same_gpu'])"
