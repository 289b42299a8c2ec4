#!/bin/bash
# experiment: whole data-parallel step (NCCL included) captured as one CUDA graph; hard timeouts (a hang must not outlive them)
NG=${NG:-2}
mkdir -p gpurun_out
ab() {
  label=$1; shift
  env "$@" timeout -k 10 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $NG --steps 40 --warmup 8 --no-cpu-baseline --no-fwd > gpurun_out/og_$label.json 2> gpurun_out/og_$label.err
  echo "$label rc=$?"; tail -3 gpurun_out/og_$label.err | cut -c1-300
  python - <<PY
import json
for l in open('gpurun_out/og_$label.json'):
    if l.startswith('{'):
        d=json.loads(l); print('$label'.ljust(20), 'ms', round(d['ms_per_step'],4), 'Mpts/s', round(d['value']/1e6,2), 'e2e', round(d['e2e']['value']/1e6,2), 'graph', d['config']['cuda_graph'])
PY
}
ab onegraph PCSEG_DDP_ONE_GRAPH=1
ab segmented PCSEG_DDP_ONE_GRAPH=0
ab onegraph_noovl PCSEG_DDP_ONE_GRAPH=1 PCSEG_DDP_OVERLAP=0
nvidia-smi --query-gpu=index,utilization.gpu,memory.used --format=csv | head -5
