#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/ragged_step.py --steps 20
timeout 300 python tools/ragged_step.py --steps 20 --frac 1.0
timeout 300 python tools/ragged_step.py --steps 20 --frac 1.0 --dense
timeout 300 python tools/ragged_step.py --steps 20 --frac 0.9
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_ragged.csv python tools/ragged_step.py --steps 1 > gpurun_out/ncu_ragged.log 2>&1
echo "rc=$?"
python tools/launch_summary.py gpurun_out/launches_ragged.csv | head -40
