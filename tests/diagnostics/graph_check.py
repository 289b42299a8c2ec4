import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import pcseg_b200
from oracle import pointnet_oracle as orc
C, B, N = 5, 4, 640
rng = np.random.default_rng(3)
x = torch.from_numpy(rng.random((B, N, 4), dtype=np.float32)).cuda()
labels = torch.from_numpy(rng.integers(0, C, (B, N)).astype(np.int64)).cuda()
def run(use_graph):
    m = pcseg_b200.PointNetSegmentation(C)
    m.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in orc.synth_state(C, 21).items()})
    m = m.cuda().train(); m.dropout.p = 0.0
    tr = pcseg_b200.FusedTrainer(m, class_weights=torch.ones(C), use_cuda_graph=use_graph)
    out = [float(tr.step(x, labels)["loss"].item()) for _ in range(8)]
    return out, tr.flat["params"].clone(), getattr(tr, "_capture_error", None), tr._graph is not None
for tag, g in (("eager A", False), ("eager B", False), ("graph  ", True)):
    l, p, err, has = run(g)
    print(tag, " ".join(f"{v:.5f}" for v in l), "graph" if has else "", err or "")
