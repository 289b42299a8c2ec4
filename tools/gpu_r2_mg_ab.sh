#!/bin/bash
# A/B of the data-parallel knobs at NG GPUs: SM reservation, NCCL CTA cap, overlap on/off
NG=${NG:-2}
mkdir -p gpurun_out
ab() {  # label, env...
  label=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29540 bench.py --gpus $NG --steps 40 --warmup 8 --no-cpu-baseline --no-fwd > gpurun_out/ab_$label.json 2> gpurun_out/ab_$label.err
  python - <<PY
import json
for l in open('gpurun_out/ab_$label.json'):
    if l.startswith('{'):
        d=json.loads(l); print('$label'.ljust(28), 'ms', round(d['ms_per_step'],4), 'Mpts/s', round(d['value']/1e6,2), d['clocks']['sm_mhz'], d['clocks']['reasons'])
PY
}
ab reserve8 PCSEG_SM_RESERVE=8
ab reserve0 PCSEG_SM_RESERVE=0
ab reserve8_ctas4 PCSEG_SM_RESERVE=8 NCCL_MAX_CTAS=4
ab reserve4_ctas4 PCSEG_SM_RESERVE=4 NCCL_MAX_CTAS=4
ab reserve16 PCSEG_SM_RESERVE=16
ab nooverlap PCSEG_SM_RESERVE=0 PCSEG_DDP_OVERLAP=0
ab reserve8_b PCSEG_SM_RESERVE=8
