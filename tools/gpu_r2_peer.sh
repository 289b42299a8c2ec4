#!/bin/bash
# peer-memory all-reduce: correctness (DDP equivalence) then A/B against NCCL.  Hard timeouts: a spinning kernel traps after
# ~tens of seconds, the process is killed after 150 s at the latest.
NG=${NG:-2}
mkdir -p gpurun_out
PCSEG_COMM=peer timeout -k 10 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29611 tools/ddp_equiv_check.py > gpurun_out/ddp_peer.log 2>&1; echo "ddp_equiv peer rc=$?"; grep "^{" gpurun_out/ddp_peer.log | cut -c1-700; grep -i "error\|timed out\|Traceback" gpurun_out/ddp_peer.log | head -5
ab() {
  label=$1; shift
  env "$@" timeout -k 10 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $NG --steps 40 --warmup 8 --no-cpu-baseline --no-fwd > gpurun_out/peer_$label.json 2> gpurun_out/peer_$label.err
  echo "$label rc=$?"; grep -i "error\|timed out" gpurun_out/peer_$label.err | head -3 | cut -c1-300
  python - <<PY
import json
for l in open('gpurun_out/peer_$label.json'):
    if l.startswith('{'):
        d=json.loads(l); print('$label'.ljust(12), 'ms', round(d['ms_per_step'],4), 'Mpts/s', round(d['value']/1e6,2), 'e2e', round(d['e2e']['value']/1e6,2), 'graph', d['config']['cuda_graph'], d['clocks']['sm_mhz'], d['clocks']['reasons'])
PY
}
ab peer PCSEG_COMM=peer
ab nccl PCSEG_COMM=nccl
ab peer_b PCSEG_COMM=peer
