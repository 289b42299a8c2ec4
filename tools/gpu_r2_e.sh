#!/bin/bash
# round 2, run E: x3 inference tests + suites touched by the deterministic Gram, then --set full capture of one step's GEMMs
mkdir -p gpurun_out
for f in test_eval_x3_gpu test_eval_gpu test_layerwise_gpu test_properties_gpu test_train_gpu test_ragged_gpu; do
  echo "=== $f"; timeout 900 python -m pytest tests/$f.py -q -m gpu -x 2>&1 | tail -${TAILN:-12}
done
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_cfg2.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/bench_cfg2.json'))
g = d['gemm_kernels']
print("TRAIN ms/step", round(d["ms_per_step"],4), "Mpts/s", round(d["value"]/1e6,2), "e2e", round(d["e2e"]["value"]/1e6,2), "clk", d["clocks"], "launches", d["gpu_launches"])
print("   ", {k: round(g[k]['ms_per_launch']*1e3,1) for k in sorted(g, key=int)})
f = d.get("fwd")
if f: print("FWD ms/step", round(f["ms_per_step"],4), "Mpts/s", round(f["value"]/1e6,2), "e2e", round(f["e2e"]["value"]/1e6,2))
PY
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-fwd"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 81 -c 27 -f -o /tmp/prof_gemm $CMD > gpurun_out/ncu_gemm.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/ncu_gemm.log
python tools/ncu_summary.py /tmp/prof_gemm.ncu-rep > gpurun_out/ncu_full_gemm.txt 2>&1; cat gpurun_out/ncu_full_gemm.txt
ls -la /tmp/prof_gemm.ncu-rep; cp /tmp/prof_gemm.ncu-rep gpurun_out/ 2>/dev/null
