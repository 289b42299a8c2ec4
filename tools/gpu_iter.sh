#!/bin/bash
# quick iteration: tests (quiet) + train/eval bench lines
mkdir -p gpurun_out
TAILN=6 bash tools/gpu_tests.sh
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err; echo "rc=$?"; tail -2 gpurun_out/bench_cfg2.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/bench_cfg2.json'))
print("TRAIN ms/step", d["ms_per_step"], "pts/s", d["value"], "e2e", d["e2e"]["value"], "frac", d["step_frac_of_bf16_sustained"], "launches", d["gpu_launches"])
print("roofline", d["roofline"])
names = {0: "fwd", 16: "dgrad", 32: "wgrad"}
tot = 0
for k, v in sorted(d["gemm_kernels"].items(), key=lambda kv: int(kv[0])):
    k = int(k); tot += v["ms_per_launch"]
    print(f"   {names[k - k % 16]:6s} conv{k % 16}: {v['ms_per_launch']*1e3:8.1f} us")
print("   GEMM total us", tot * 1e3)
PY
timeout 600 python bench.py --workload cfg2_eval --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_cfg2_eval.json 2> gpurun_out/bench_cfg2_eval.err; echo "rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_cfg2_eval.json')); print('EVAL ms/step', d['ms_per_step'], 'pts/s', d['value'], 'e2e', d['e2e']['value'], 'frac', d['step_frac_of_bf16_sustained'])"
