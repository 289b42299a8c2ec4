#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/sharded_eval_check.py 2>&1 | tail -5
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload cfg5_eval_sharded --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_cfg5_sharded_2gpu.json 2> gpurun_out/bench_cfg5_sharded_2gpu.err; echo "rc=$?"; tail -3 gpurun_out/bench_cfg5_sharded_2gpu.err
python -c "
import json; d=json.load(open('gpurun_out/bench_cfg5_sharded_2gpu.json')); print('SHARDED2', d['ms_per_step'], d['value'], d['e2e']['value'], d['scaling'])"
timeout 300 python bench.py --workload cfg5_eval_sharded --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_cfg5_sharded_1gpu.json 2> gpurun_out/bench_cfg5_sharded_1gpu.err; echo "rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_cfg5_sharded_1gpu.json')); print('SHARDED1', d['ms_per_step'], d['value'], d['e2e']['value'], d['scaling'])"
timeout 300 python -m pytest tests/test_properties_gpu.py -q -k sharded 2>&1 | tail -3
