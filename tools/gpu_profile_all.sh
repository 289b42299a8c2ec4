#!/bin/bash
# end-of-round evidence: full ncu captures (GEMMs, pointwise kernels) + launch lists (train, eval, ragged train step)
bash tools/gpu_profile.sh
rm -f gpurun_out/prof_gemm.ncu-rep
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 210 -c 140 --csv --log-file gpurun_out/launches_train.csv $CMD > gpurun_out/ncu_list_train.log 2>&1; echo "list train rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_eval.csv $CMD --workload cfg2_eval > gpurun_out/ncu_list_eval.log 2>&1; echo "list eval rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_ragged.csv python tools/ragged_step.py --steps 1 --frac 0.5 > gpurun_out/ncu_ragged.log 2>&1; echo "list ragged rc=$?"
python tools/launch_summary.py gpurun_out/launches_train.csv > gpurun_out/launch_summary_train.txt
python tools/launch_summary.py gpurun_out/launches_eval.csv > gpurun_out/launch_summary_eval.txt
python tools/launch_summary.py gpurun_out/launches_ragged.csv > gpurun_out/launch_summary_ragged.txt
head -12 gpurun_out/launch_summary_train.txt; head -8 gpurun_out/launch_summary_ragged.txt
du -sh gpurun_out
