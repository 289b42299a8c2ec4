#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_train_gpu.py tests/test_properties_gpu.py -x -q 2>&1 | tail -3
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 50 --warmup 10 > gpurun_out/bench_cfg2_2gpu.json 2> gpurun_out/bench_cfg2_2gpu.err; echo "rc=$?"; tail -2 gpurun_out/bench_cfg2_2gpu.err
python -c "
import json
for l in open('gpurun_out/bench_cfg2_2gpu.json'):
    if l.startswith('{'):
        d=json.loads(l); print('2GPU', d['ms_per_step'], d['value'], d['e2e']['value'], d['config']['cuda_graph'], d['clocks'])"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535 tools/ddp_equiv_check.py 2>&1 | tail -3
