"""A few ragged training steps (for ncu launch lists / host-time probes): python tools/ragged_step.py [--frac 0.25 --steps 3]"""
import argparse
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import pcseg_b200  # noqa: E402
from ragged_sweep import lengths_for  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--B", type=int, default=8)
ap.add_argument("--N", type=int, default=16384)
ap.add_argument("--frac", type=float, default=0.25)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--dense", action="store_true")
a = ap.parse_args()
dev = torch.device("cuda", 0)
C = 5
lengths = lengths_for(a.B, a.N, a.frac, np.random.default_rng(0))
x = torch.rand(a.B, a.N, 4, device=dev)
y = torch.randint(0, C, (a.B, a.N), device=dev)
for b, L in enumerate(lengths):
    x[b, L:] = 0
    y[b, L:] = -1
torch.manual_seed(0)
m = pcseg_b200.PointNetSegmentation(C).to(dev).train()
tr = pcseg_b200.FusedTrainer(m, class_weights=torch.ones(C, device=dev), use_cuda_graph=False)
ls = None if a.dense else lengths
for _ in range(2):
    tr.step(x, y, lengths=ls)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
e0.record()
for _ in range(a.steps):
    tr.step(x, y, lengths=ls)
e1.record()
t_host = time.perf_counter() - t0
torch.cuda.synchronize()
print(f"lengths {lengths} valid {sum(lengths)} host {t_host / a.steps * 1e3:.3f} ms/step, device {e0.elapsed_time(e1) / a.steps:.3f} ms/step")
