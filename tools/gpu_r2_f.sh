#!/bin/bash
# round 2, run F: layerwise (incl. the 256-tile case) + parity tests, then the bench line with per-kernel stamps
mkdir -p gpurun_out
for f in test_layerwise_gpu test_parity_baseline_sizes_gpu; do
  echo "=== $f"; timeout 1500 python -m pytest tests/$f.py -q -m gpu -x 2>&1 | tail -${TAILN:-12}
done
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_cfg2.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/bench_cfg2.json'))
g = d['gemm_kernels']
print("TRAIN ms/step", round(d["ms_per_step"],4), "Mpts/s", round(d["value"]/1e6,2), "e2e", round(d["e2e"]["value"]/1e6,2), "clk", d["clocks"], "launches", d["gpu_launches"])
tot = 0
for k in sorted(g, key=int):
    v = g[k]; per_step = v['ms_per_launch'] * v['launches'] / d['steps'] * 1e3; tot += per_step
    print(f"   tag {k:>3s} {v.get('kernel',''):42s} {v['ms_per_launch']*1e3:8.1f} us x {v['launches']//d['steps']:2d} = {per_step:8.1f} us/step")
print("   sum of stamped kernels us/step", round(tot,1))
f = d.get("fwd")
if f: print("FWD ms/step", round(f["ms_per_step"],4), "Mpts/s", round(f["value"]/1e6,2), "e2e", round(f["e2e"]["value"]/1e6,2))
PY
