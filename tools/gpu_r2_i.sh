#!/bin/bash
# validation: whole GPU test-suite + smoke + bench with stamps
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -m gpu -x 2>&1 | tail -12 | cut -c1-400
echo "=== smoke"; timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -3
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_cfg2.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/bench_cfg2.json'))
g = d['gemm_kernels']
print("TRAIN ms/step", round(d["ms_per_step"],4), "Mpts/s", round(d["value"]/1e6,2), "e2e", round(d["e2e"]["value"]/1e6,2), "clk", d["clocks"], "launches", d["gpu_launches"])
for k in sorted(g, key=int):
    v = g[k]; per_step = v['ms_per_launch'] * v['launches'] / d['steps'] * 1e3
    print(f"   tag {k:>3s} {v.get('kernel',''):42s} {v['ms_per_launch']*1e3:8.1f} us x {v['launches']//d['steps']:2d} = {per_step:8.1f} us/step")
f = d.get("fwd")
if f: print("FWD ms/step", round(f["ms_per_step"],4), "Mpts/s", round(f["value"]/1e6,2), "e2e", round(f["e2e"]["value"]/1e6,2))
print("roofline", d["roofline"]["kernel"][:50], round(d["roofline"]["frac"],3))
PY
