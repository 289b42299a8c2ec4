"""device time of the fused training step with and without dropout (CUDA-graph replay): the difference is what the two
dropout layers cost (Philox regeneration in forward and backward + mask application)"""
import os, sys
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import pcseg_b200
B, N, C = 8, 16384, 5
dev = torch.device("cuda", 0)
x = torch.rand(B, N, 4, device=dev)
y = torch.randint(0, C, (B, N), device=dev)
for p in (0.3, 0.0, 0.3, 0.0):
    torch.manual_seed(0)
    m = pcseg_b200.PointNetSegmentation(C).to(dev).train()
    m.dropout.p = p
    tr = pcseg_b200.FusedTrainer(m, class_weights=torch.ones(C, device=dev))
    for _ in range(6):
        tr.step(x, y)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(40):
        tr.step(x, y)
    e1.record()
    torch.cuda.synchronize()
    print(f"dropout p={p}: {e0.elapsed_time(e1) / 40:.4f} ms/step (graph={tr._graph is not None})")
