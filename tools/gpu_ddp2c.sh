#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535 tools/ddp_equiv_check.py 2>&1 | tail -2
timeout 400 python -m pytest tests/test_train_gpu.py tests/test_properties_gpu.py -q 2>&1 | grep -E "passed|failed|FAILED|Error" | head -8
