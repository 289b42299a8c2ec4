#!/bin/bash
# same-box comparison: 1 GPU, then NG GPUs with the peer-memory all-reduce and with NCCL
NG=${NG:-2}
mkdir -p gpurun_out
show() { python - <<PY
import json
for l in open('$1'):
    if l.startswith('{'):
        d=json.loads(l); print('$2'.ljust(12), 'ms', round(d['ms_per_step'],4), 'Mpts/s', round(d['value']/1e6,2), 'e2e', round(d['e2e']['value']/1e6,2), d['clocks']['sm_mhz'], d['clocks']['reasons'])
PY
}
timeout 300 python bench.py --steps 40 --warmup 8 --no-cpu-baseline --no-fwd > gpurun_out/same_1gpu.json 2>/dev/null; show gpurun_out/same_1gpu.json 1gpu
for mode in peer skip nccl; do
  PCSEG_COMM=$( [ $mode = skip ] && echo peer || echo $mode ) PCSEG_PEER_SKIP=$( [ $mode = skip ] && echo 1 || echo 0 ) timeout -k 10 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus $NG --steps 40 --warmup 8 --no-cpu-baseline --no-fwd > gpurun_out/same_${mode}_${NG}gpu.json 2>/dev/null; show gpurun_out/same_${mode}_${NG}gpu.json ${mode}_${NG}
done
timeout 300 python bench.py --steps 40 --warmup 8 --no-cpu-baseline --no-fwd > gpurun_out/same_1gpu_b.json 2>/dev/null; show gpurun_out/same_1gpu_b.json 1gpu_again
python - <<'PY'
import json
for l in open('gpurun_out/same_peer_2gpu.json'):
    if l.startswith('{'):
        d=json.loads(l); print('peer kernel stamp', d['gemm_kernels'].get('94'))
PY
