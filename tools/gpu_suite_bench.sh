#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -3
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err; echo "rc=$?"
python - <<'PY'
import json
d = json.load(open('gpurun_out/bench_cfg2.json'))
g = d['gemm_kernels']
print("TRAIN ms/step", d["ms_per_step"], "pts/s", d["value"], "e2e", d["e2e"]["value"], {k: round(g[k]['ms_per_launch']*1e3,1) for k in ['5','21','37','4','6']}, d['roofline']['frac'])
PY
