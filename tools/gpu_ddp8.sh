#!/bin/bash
mkdir -p gpurun_out
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/bench_cfg2_8gpu.json 2> gpurun_out/bench_cfg2_8gpu.err; echo "rc=$?"; tail -2 gpurun_out/bench_cfg2_8gpu.err
python -c "
import json
for l in open('gpurun_out/bench_cfg2_8gpu.json'):
    if l.startswith('{'):
        d=json.loads(l); print('8GPU', d['ms_per_step'], d['value'], d['e2e']['value'], d['config']['cuda_graph'], d['clocks'])"
