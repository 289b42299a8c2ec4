#!/bin/bash
# round 2, run D: whole GPU test-suite, smoke, the default bench line, source-level ncu capture of the two big GEMMs
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -25
echo "=== smoke"; timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -4
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_cfg2.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/bench_cfg2.json'))
g = d['gemm_kernels']
print("TRAIN ms/step", round(d["ms_per_step"],4), "Mpts/s", round(d["value"]/1e6,2), "e2e", round(d["e2e"]["value"]/1e6,2), "clk", d["clocks"], "launches", d["gpu_launches"])
print("   ", {k: round(g[k]['ms_per_launch']*1e3,1) for k in sorted(g, key=int)})
print("   roofline", d["roofline"]["kernel"], round(d["roofline"]["frac"],3), "cpu", d["cpu_baseline"], "eager", d["torch_eager_same_gpu"])
f = d.get("fwd")
if f: print("FWD ms/step", round(f["ms_per_step"],4), "Mpts/s", round(f["value"]/1e6,2), "e2e", round(f["e2e"]["value"]/1e6,2), f.get("roofline"))
PY
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-fwd"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"gemm_kernel<256, 6|gemm_kernel<256, 8" -s 3 -c 2 -o gpurun_out/prof_big -f $CMD > gpurun_out/ncu_k.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/ncu_k.log; ls -la gpurun_out/*.ncu-rep
