#!/bin/bash
# round 2, run H: tests for the fold kernels, bench with stamps, ncu of the small fold / predict kernels
mkdir -p gpurun_out
for f in test_layerwise_gpu; do
  echo "=== $f"; timeout 1500 python -m pytest tests/$f.py -q -m gpu -x 2>&1 | tail -${TAILN:-8} | cut -c1-300
done
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-fwd > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_cfg2.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/bench_cfg2.json'))
g = d['gemm_kernels']
print("TRAIN ms/step", round(d["ms_per_step"],4), "Mpts/s", round(d["value"]/1e6,2), "e2e", round(d["e2e"]["value"]/1e6,2), "clk", d["clocks"], "launches", d["gpu_launches"])
for k in sorted(g, key=int):
    v = g[k]; per_step = v['ms_per_launch'] * v['launches'] / d['steps'] * 1e3
    print(f"   tag {k:>3s} {v.get('kernel',''):42s} {v['ms_per_launch']*1e3:8.1f} us x {v['launches']//d['steps']:2d} = {per_step:8.1f} us/step")
PY
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-fwd"
timeout 900 ncu --set full --clock-control none -k regex:"k_predict|k_fold|k_gram|k_pool|k_convert" -s 36 -c 14 -f -o /tmp/prof_small $CMD > gpurun_out/ncu_small.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/ncu_small.log
python tools/ncu_summary.py /tmp/prof_small.ncu-rep 2>&1 | cut -c1-200
ncu -i /tmp/prof_small.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]; ix={n:i for i,n in enumerate(h)}
want=['smsp__average_warp_latency_issue_stalled_long_scoreboard_per_warp_active.pct','smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','launch__grid_size','launch__block_size']
for r in rows[2:]:
    print(r[ix['Kernel Name']][:40].ljust(40), [ (w.split('.')[0][-28:], r[ix[w]]) for w in want if w in ix])
"
