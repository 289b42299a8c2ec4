#!/bin/bash
# launch list + full captures of one training step's kernels (run after the plain command exited 0)
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
timeout 600 $CMD > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 290 -c 200 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "list rc=$?"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 69 -c 23 -o gpurun_out/prof_gemm $CMD > gpurun_out/ncu_gemm.log 2>&1
echo "gemm rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"k_bn_bwd_apply|k_bn_relu|k_head|k_ingest" -s 60 -c 20 -o gpurun_out/prof_ew $CMD > gpurun_out/ncu_ew.log 2>&1
echo "ew rc=$?"
ls -la gpurun_out/*.ncu-rep
