#!/bin/bash
# ragged-execution tests first, then the whole GPU suite
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ragged_gpu.py -q 2>&1 | tail -40
timeout 900 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -5
