#!/bin/bash
# round 2, run B: folded conv5 after the k_fold_bwd restructure: tests, parity report, A/B bench, ncu launch list
mkdir -p gpurun_out
for f in test_layerwise_gpu test_train_gpu; do
  echo "=== $f"; timeout 900 python -m pytest tests/$f.py -q -m gpu -x 2>&1 | tail -${TAILN:-25}
done
timeout 600 python tools/train_parity_report.py 8 2048 5 2>&1 | tail -40
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
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 340 -c 200 --csv --log-file gpurun_out/launches_fold1.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/ncu.log
python tools/launch_summary.py gpurun_out/launches_fold1.csv 2>&1 | head -45
