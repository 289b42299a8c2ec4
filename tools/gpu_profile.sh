#!/bin/bash
# launch list + full captures of one training step's kernels (run after the plain command exited 0).
# Only text summaries (and the GEMM report) are brought back: gpurun merges at most 64 MiB.
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
timeout 600 $CMD > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; exit 1; }
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 69 -c 23 -f -o /tmp/prof_gemm $CMD > gpurun_out/ncu_gemm.log 2>&1
echo "gemm rc=$?"
python tools/ncu_summary.py /tmp/prof_gemm.ncu-rep > gpurun_out/ncu_full_gemm.txt 2>&1
timeout 1200 ncu --set full --clock-control none -k regex:"k_bn_bwd_apply|k_bn_relu|k_head|k_ingest|k_convert|k_adam|k_cloud|k_maxpool" -s 81 -c 27 -f -o /tmp/prof_ew $CMD > gpurun_out/ncu_ew.log 2>&1
echo "ew rc=$?"
python tools/ncu_summary.py /tmp/prof_ew.ncu-rep > gpurun_out/ncu_full_pointwise.txt 2>&1
cp /tmp/prof_gemm.ncu-rep gpurun_out/ 2>/dev/null
ls -la gpurun_out/
