#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err; echo "rc=$?"; cat gpurun_out/bench_cfg2.json; tail -3 gpurun_out/bench_cfg2.err
timeout 600 python bench.py --workload cfg2_eval --steps 20 --warmup 5 > gpurun_out/bench_cfg2_eval.json 2> gpurun_out/bench_cfg2_eval.err; echo "rc=$?"; cat gpurun_out/bench_cfg2_eval.json; tail -3 gpurun_out/bench_cfg2_eval.err
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 340 -c 260 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/ncu.log
