#!/bin/bash
mkdir -p gpurun_out
echo "--- PDL off"; timeout 300 python tools/ragged_sweep.py --fracs 0.5,0.25 --steps 30 2>&1 >/dev/null | tail -2
echo "--- PDL on"; PCSEG_PDL=1 timeout 300 python tools/ragged_sweep.py --fracs 0.5,0.25 --steps 30 2>&1 >/dev/null | tail -2
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err; echo "rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_cfg2.json')); print('TRAIN ms/step', d['ms_per_step'], 'pts/s', d['value'], 'e2e', d['e2e']['value'])"
