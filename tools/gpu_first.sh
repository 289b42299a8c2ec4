#!/bin/bash
# first GPU bring-up: each step under its own timeout so a hang cannot eat the box
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv | tee gpurun_out/gpu.txt
echo "=== debug k" ; timeout 120 python tools/debug_gemm.py k 2>&1 | tail -40
echo "=== debug mn"; timeout 120 python tools/debug_gemm.py mn 2>&1 | tail -40
echo "=== pytest gemm"; timeout 600 python -m pytest tests/test_gemm_gpu.py -q -m gpu -x 2>&1 | tail -25
echo "=== pytest eval"; timeout 600 python -m pytest tests/test_eval_gpu.py -q -m gpu -x 2>&1 | tail -25
echo "=== pytest train"; timeout 900 python -m pytest tests/test_train_gpu.py -q -m gpu -x 2>&1 | tail -40
