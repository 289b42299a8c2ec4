#!/bin/bash
# final round measurements: all workloads + launch lists (train and eval)
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err; echo "cfg2 rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_cfg2_reference.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
for w in cfg2_eval cfg3_train cfg3_eval cfg5_eval cfg4; do
  timeout 900 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; echo "$w rc=$?"
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/bench_*.json')):
    try:
        d = json.load(open(f))
        print(f.split('/')[-1], 'ms', round(d['ms_per_step'],3), 'Mpts/s', round(d['value']/1e6,3), 'e2e', round(d['e2e']['value']/1e6,3), 'frac', round(d.get('step_frac_of_bf16_sustained',0),3), 'clk', d.get('clocks'))
    except Exception as e:
        print(f, 'ERR', e)
PY
bash tools/gpu_list_only.sh; cp gpurun_out/launches.csv gpurun_out/launches_train.csv
WL=cfg2_eval SKIP=24 CNT=32 bash tools/gpu_list_only.sh; cp gpurun_out/launches.csv gpurun_out/launches_eval.csv
