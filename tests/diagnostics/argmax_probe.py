"""argmax agreement of the CUDA inference path with the fp32 oracle on ALL points (no near-tie exclusion)"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import pcseg_b200
from oracle import pointnet_oracle as orc
for (B, N, C, seed) in [(1, 16384, 5, 1), (2, 8192, 5, 2), (2, 8192, 3, 3), (4, 4096, 8, 4)]:
    sd = orc.synth_state(C, seed)
    m = pcseg_b200.PointNetSegmentation(C)
    m.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()})
    m = m.cuda().eval()
    x = np.random.default_rng(seed).random((B, N, 4), dtype=np.float32)
    with torch.no_grad():
        got = m(torch.from_numpy(x).cuda()).cpu().numpy()
    ref = orc.forward_eval(sd, x, dtype=np.float64)
    err = np.abs(got - ref).max() / np.abs(ref).max()
    agree = (got.argmax(-1) == ref.argmax(-1)).mean()
    top2 = np.sort(ref, axis=-1)[..., -2:]
    margin = (top2[..., 1] - top2[..., 0]) / np.abs(ref).max()
    print(f"B{B} N{N} C{C}: rel-to-max err {err:.2e}, argmax agreement on all points {agree:.5f}, median top-2 margin/max {np.median(margin):.3f}, "
          f"points with margin < 1e-2: {(margin < 1e-2).mean():.4f}")
