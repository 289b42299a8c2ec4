#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 340 -c 260 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu.log 2>&1; echo "ncu rc=$?"
timeout 300 python -m pytest tests/test_train_gpu.py -q -m gpu -x 2>&1 | grep -E "Error|assert|bad|FAILED|passed|failed" | head -20
