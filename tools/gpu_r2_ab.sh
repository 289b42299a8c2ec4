#!/bin/bash
# round 2 A/B on one box: tests of the default build, then the cfg2 bench line of every variant with per-kernel stamps.
# VARIANTS entries: name[:lib-suffix[:ENV=VAL]]   (lib/libpcseg_b200<suffix>.so)
mkdir -p gpurun_out
for f in ${TESTS:-test_layerwise_gpu test_gemm_gpu}; do
  echo "=== $f"; timeout 900 python -m pytest tests/$f.py -q -m gpu -x 2>&1 | tail -${TAILN:-6} | cut -c1-300
done
for spec in ${VARIANTS:-base}; do
  IFS=: read -r v sfx envs <<< "$spec"
  env PCSEG_LIB_SUFFIX=$sfx $envs timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-fwd > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err || { echo "variant '$v' FAILED"; tail -5 gpurun_out/ab_$v.err; }
  V=$v python - <<'PY'
import json, os
v = os.environ["V"]
try:
    d = json.load(open(f'gpurun_out/ab_{v}.json'))
except Exception as e:
    print("variant", v, "no json", e); raise SystemExit
g = d['gemm_kernels']
tags = ["4", "5", "6", "7", "20", "21", "23", "24", "36", "53"]
print(f"variant '{v}': ms/step {d['ms_per_step']:.4f}  Mpts/s {d['value']/1e6:.2f}  clk {d['clocks']['sm_mhz']} {d['clocks']['reasons']}  " +
      "  ".join(f"t{t}={g[t]['ms_per_launch']*1e3:.1f}x{g[t]['launches']//d['steps']}" for t in tags if t in g))
PY
done
