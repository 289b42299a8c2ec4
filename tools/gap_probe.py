"""How much of the training step is GPU idle time between kernels?  (torch.profiler / CUPTI kernel timeline)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pcseg_b200
from torch.profiler import profile, ProfilerActivity

B, N, C = 8, 16384, 5
torch.manual_seed(0)
m = pcseg_b200.PointNetSegmentation(C).cuda().train()
tr = pcseg_b200.FusedTrainer(m, class_weights=torch.ones(C))
x = torch.rand(B, N, 4, device="cuda"); lab = torch.randint(0, C, (B, N), device="cuda")
for _ in range(5): tr.step(x, lab)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(5): tr.step(x, lab)
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
t0, t1 = evs[0].time_range.start, max(e.time_range.end for e in evs)
busy = sum(e.time_range.end - e.time_range.start for e in evs)
print(f"span {(t1-t0)/5:.1f} us/step, kernel-busy {busy/5:.1f} us/step, idle {(t1-t0-busy)/5:.1f} us/step, events/step {len(evs)/5:.0f}")
import collections
agg = collections.Counter(); cnt = collections.Counter()
for e in evs:
    agg[e.name[:60]] += e.time_range.end - e.time_range.start; cnt[e.name[:60]] += 1
for k, v in agg.most_common(14):
    print(f"  {v/5:8.1f} us/step  n={cnt[k]/5:4.1f}  {k}")
# CPU time per step
import time
torch.cuda.synchronize(); t=time.perf_counter()
for _ in range(20): tr.step(x, lab)
t_cpu=time.perf_counter()-t; torch.cuda.synchronize(); t_all=time.perf_counter()-t
print(f"CPU enqueue time per step {t_cpu/20*1e3:.3f} ms; total {t_all/20*1e3:.3f} ms")
