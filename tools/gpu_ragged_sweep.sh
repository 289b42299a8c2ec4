#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_ragged_gpu.py -q 2>&1 | tail -5
timeout 600 python tools/ragged_sweep.py > gpurun_out/ragged_sweep_8x16k.json 2> gpurun_out/ragged_sweep_8x16k.err; echo "rc=$?"; tail -8 gpurun_out/ragged_sweep_8x16k.err
timeout 600 python tools/ragged_sweep.py --B 8 --N 65536 --fracs 1.0,0.75,0.5,0.25 --steps 10 > gpurun_out/ragged_sweep_8x64k.json 2> gpurun_out/ragged_sweep_8x64k.err; echo "rc=$?"; tail -8 gpurun_out/ragged_sweep_8x64k.err
