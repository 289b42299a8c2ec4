#!/bin/bash
# end-of-round measurements: smoke, every bench workload, reference arm, ragged sweeps
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err; echo "cfg2 rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_cfg2_reference.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
timeout 900 python bench.py --workload cfg2_eval > gpurun_out/bench_cfg2_eval.json 2> gpurun_out/bench_cfg2_eval.err; echo "cfg2_eval rc=$?"
for w in cfg3_train cfg3_eval cfg5_eval cfg5_eval_sharded cfg4; do
  timeout 900 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; echo "$w rc=$?"
done
timeout 600 python tools/ragged_sweep.py > gpurun_out/ragged_sweep_8x16k.json 2> gpurun_out/ragged_sweep_8x16k.err; tail -5 gpurun_out/ragged_sweep_8x16k.err
timeout 600 python tools/ragged_sweep.py --B 8 --N 65536 --fracs 1.0,0.75,0.5,0.25 --steps 10 > gpurun_out/ragged_sweep_8x64k.json 2> gpurun_out/ragged_sweep_8x64k.err; tail -4 gpurun_out/ragged_sweep_8x64k.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/bench_cfg*.json')):
    try:
        d = [json.loads(l) for l in open(f) if l.startswith('{')][-1]
        print(f.split('/')[-1], 'ms', round(d['ms_per_step'],3), 'Mpts/s', round(d['value']/1e6,3), 'e2e', round(d['e2e']['value']/1e6,3), 'frac', round(d.get('step_frac_of_bf16_sustained',0),3), 'clk', d.get('clocks'), 'cpu', (d.get('cpu_baseline') or {}).get('value'), 'eager', {k: round(v/1e6,2) for k, v in (d.get('torch_eager_same_gpu') or {}).items() if isinstance(v, float)})
    except Exception as e:
        print(f, 'ERR', e)
PY
