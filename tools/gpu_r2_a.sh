#!/bin/bash
# round 2, run A: folded conv5 (Gram-predicted BN + folded BN backward): parity tests, then A/B bench against the legacy step
mkdir -p gpurun_out
for f in test_gemm_gpu test_layerwise_gpu test_train_gpu; do
  echo "=== $f"; timeout 900 python -m pytest tests/$f.py -q -m gpu -x 2>&1 | tail -${TAILN:-25}
done
for fold in 1 0; do
  PCSEG_FOLDED=$fold timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_cfg2_fold$fold.json 2> gpurun_out/bench_cfg2_fold$fold.err; echo "fold=$fold rc=$?"; tail -2 gpurun_out/bench_cfg2_fold$fold.err
  python - <<PY
import json
d = json.load(open('gpurun_out/bench_cfg2_fold$fold.json'))
g = d['gemm_kernels']
print("fold=$fold TRAIN ms/step", round(d["ms_per_step"],4), "Mpts/s", round(d["value"]/1e6,2), "e2e", round(d["e2e"]["value"]/1e6,2), "clk", d["clocks"])
print("   ", {k: round(g[k]['ms_per_launch']*1e3,1) for k in sorted(g, key=int)})
PY
done
