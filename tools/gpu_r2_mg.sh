#!/bin/bash
# round 2 multi-GPU run: usage NG=2 bash tools/gpu_r2_mg.sh   (hardware parity tests, weak cfg2 and strong cfg4 lines)
NG=${NG:-2}
mkdir -p gpurun_out
if [ "${SKIP_TESTS:-0}" != "1" ]; then
  timeout 900 python -m pytest tests/test_multigpu_gpu.py -q -m gpu -s 2>&1 | grep -E "DDP-EQUIV|bit-identical|passed|failed|skipped|Error" | cut -c1-700
fi
run() {  # name, extra args
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port $3 bench.py --gpus $NG --steps 30 --warmup 8 --no-cpu-baseline $2 > gpurun_out/bench_$1_${NG}gpu.json 2> gpurun_out/bench_$1_${NG}gpu.err; echo "$1 rc=$?"; tail -2 gpurun_out/bench_$1_${NG}gpu.err
  python - <<PY
import json
for l in open('gpurun_out/bench_$1_${NG}gpu.json'):
    if l.startswith('{'):
        d=json.loads(l); print('$1 ${NG}GPU ms', round(d['ms_per_step'],4), 'Mpts/s', round(d['value']/1e6,2), 'e2e', round(d['e2e']['value']/1e6,2), d['scaling'], d['config'].get('global_batch'), 'graph', d['config']['cuda_graph'], d['clocks'])
        f=d.get('fwd')
        if f: print('   fwd Mpts/s', round(f['value']/1e6,2), 'e2e', round(f['e2e']['value']/1e6,2))
PY
}
run cfg2 "" 29533
run cfg4_strong "--workload cfg4_strong --no-fwd" 29534
if [ "$NG" = "1" ]; then exit 0; fi
