#!/bin/bash
# round 2, run G: folded seg_conv1: layerwise + train + properties tests, A/B bench (PCSEG_FOLD6=1/0) with kernel stamps
mkdir -p gpurun_out
for f in test_layerwise_gpu test_train_gpu test_properties_gpu; do
  echo "=== $f"; timeout 1500 python -m pytest tests/$f.py -q -m gpu -x 2>&1 | tail -${TAILN:-14} | cut -c1-400
done
for f6 in 1 0; do
PCSEG_FOLD6=$f6 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-fwd > gpurun_out/bench_cfg2_f6$f6.json 2> gpurun_out/bench_cfg2_f6$f6.err; echo "bench fold6=$f6 rc=$?"; tail -3 gpurun_out/bench_cfg2_f6$f6.err
python - <<PY
import json
d = json.load(open('gpurun_out/bench_cfg2_f6$f6.json'))
g = d['gemm_kernels']
print("fold6=$f6 TRAIN ms/step", round(d["ms_per_step"],4), "Mpts/s", round(d["value"]/1e6,2), "e2e", round(d["e2e"]["value"]/1e6,2), "clk", d["clocks"], "launches", d["gpu_launches"])
tot = 0
for k in sorted(g, key=int):
    v = g[k]; per_step = v['ms_per_launch'] * v['launches'] / d['steps'] * 1e3; tot += per_step
    print(f"   tag {k:>3s} {v.get('kernel',''):42s} {v['ms_per_launch']*1e3:8.1f} us x {v['launches']//d['steps']:2d} = {per_step:8.1f} us/step")
print("   sum of stamped kernels us/step", round(tot,1))
PY
done
