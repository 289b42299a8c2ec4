#!/bin/bash
# diagnostic: MMA-warp wait accounting (lib built with -DPCSEG_PROF_WAIT) for the K >= 1024 GEMMs of one eager training step
mkdir -p gpurun_out
PCSEG_LIB_SUFFIX=_prof timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-fwd 2>&1 | grep PROF > gpurun_out/prof_wait.txt
python - <<'PY'
import re, collections
agg = collections.defaultdict(list)
for l in open('gpurun_out/prof_wait.txt'):
    m = re.match(r"PROF (gemm<[^>]*>) cta (\d+) tiles (\d+): total (\d+) cyc, wait accumulator (\d+), wait operands (\d+), issue (\d+), k-blocks (\d+) ready (\d+)", l)
    if m: agg[(m.group(1), int(m.group(3)))].append(tuple(int(m.group(i)) for i in (4, 5, 6, 7, 8, 9)))
for k, v in sorted(agg.items()):
    n = len(v); t = sum(x[0] for x in v) / n; a = sum(x[1] for x in v) / n; f = sum(x[2] for x in v) / n
    iss = sum(x[3] for x in v) / n; kb = sum(x[4] for x in v) / n; rdy = sum(x[5] for x in v) / n
    print(f"{k[0]:18s} tiles/CTA {k[1]:3d}  n={n:3d}  total {t:9.0f} cyc  wait-accumulator {a:8.0f} ({100*a/t:4.1f}%)  wait-operands {f:8.0f} ({100*f/t:4.1f}%)  "
          f"issue {iss:8.0f} ({100*iss/t:4.1f}%)  k-blocks {kb:6.0f}, operands ready at first probe {100*rdy/max(kb,1):4.1f}%, wait per k-block {f/max(kb,1):5.0f} cyc, issue per k-block {iss/max(kb,1):5.0f} cyc")
PY
