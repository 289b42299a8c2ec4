#!/bin/bash
# usage: KREGEX=... SKIP=.. CNT=.. OUT=name bash tools/gpu_prof_kernel.sh
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline ${WL:+--workload $WL}"
timeout 600 $CMD > gpurun_out/plain.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"$KREGEX" -s ${SKIP:-3} -c ${CNT:-2} -o gpurun_out/${OUT:-prof_k} -f $CMD > gpurun_out/ncu_k.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/ncu_k.log
