#!/bin/bash
mkdir -p gpurun_out
for f in test_gemm_gpu test_eval_gpu test_layerwise_gpu test_train_gpu test_properties_gpu; do
  echo "=== $f"; timeout 900 python -m pytest tests/$f.py -q -m gpu -x 2>&1 | tail -${TAILN:-30}
done
echo "=== smoke"; timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -5
