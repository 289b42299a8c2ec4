"""GPU-side diagnostic for the training path: CUDA logits/grads vs fp64 oracle vs a bf16-rounding emulation."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch

import pcseg_b200
from oracle import pointnet_oracle as orc


def bf(a):
    return torch.from_numpy(np.asarray(a, np.float32)).bfloat16().float().numpy().astype(np.float64)


def emulate(sd, x):
    """forward_train with y and a rounded to bf16 where the CUDA path stores bf16."""
    B, N, _ = x.shape
    P = B * N
    a = x.reshape(P, 4).astype(np.float64)
    outs = {}

    def block(a_in, conv, bn, first=False, extra=None):
        W = sd[f"{conv}.weight"][:, :, 0].astype(np.float64)
        if not first:
            W = bf(W)
        y = a_in @ W.T
        if extra is not None:
            y = y + extra
        y = bf(y)
        mean = y.mean(0)
        var = y.var(0)
        inv = 1 / np.sqrt(var + 1e-5)
        g = sd[f"{bn}.weight"].astype(np.float64)
        b = sd[f"{bn}.bias"].astype(np.float64)
        z = (y - mean) * inv * g + b
        outs[conv] = y
        return z

    a = bf(np.maximum(block(a, "conv1", "bn1", first=True), 0))
    pf = a = bf(np.maximum(block(a, "conv2", "bn2"), 0))
    a = bf(np.maximum(block(a, "conv3", "bn3"), 0))
    a = bf(np.maximum(block(a, "conv4", "bn4"), 0))
    a = bf(np.maximum(block(a, "conv5", "bn5"), 0))
    z = np.maximum(block(a, "global_feat", "bn_global"), 0)
    g = z.reshape(B, N, -1).max(1)
    Wg = sd["seg_conv1.weight"][:, 64:, 0].astype(np.float64)
    cb = g @ Wg.T
    sd2 = dict(sd)
    sd2["seg_conv1.weight"] = sd["seg_conv1.weight"][:, :64]
    a = bf(np.maximum(block(pf, "seg_conv1", "bn_seg1", extra=np.repeat(cb, N, axis=0)) if False else
                      _blk(sd, pf, np.repeat(cb, N, axis=0)), 0))
    a = bf(np.maximum(block(a, "seg_conv2", "bn_seg2"), 0))
    z = np.maximum(block(a, "seg_conv3", "bn_seg3"), 0)
    W4 = sd["seg_conv4.weight"][:, :, 0].astype(np.float64)
    return (z @ W4.T + sd["seg_conv4.bias"]).reshape(B, N, -1)


def _blk(sd, pf, extra):
    W = bf(sd["seg_conv1.weight"][:, :64, 0].astype(np.float64))
    y = bf(pf @ W.T + extra)
    mean, var = y.mean(0), y.var(0)
    return (y - mean) / np.sqrt(var + 1e-5) * sd["bn_seg1.weight"] + sd["bn_seg1.bias"]


def main():
    cases = [(5, 2, 96, 11), (5, 4, 512, 1), (5, 8, 2048, 2)]
    if len(sys.argv) > 1:                      # e.g. "24,4,1024,1 5,4,1024,1"
        cases = [tuple(int(v) for v in a.split(",")) for a in sys.argv[1:]]
    for (C, B, N, seed) in cases:
        sd = orc.synth_state(C, seed)
        rng = np.random.default_rng(seed + 1)
        x = rng.random((B, N, 4), dtype=np.float32)
        labels = rng.integers(0, C, (B, N)).astype(np.int64)
        cw = np.ones(C, np.float32)
        m = pcseg_b200.PointNetSegmentation(C)
        m.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()})
        m = m.cuda().train()
        m.dropout.p = 0.0
        xt = torch.from_numpy(x).cuda()
        logits = m(xt)
        crit = torch.nn.CrossEntropyLoss(ignore_index=-1, weight=torch.from_numpy(cw).cuda())
        loss = crit(logits.view(-1, C), torch.from_numpy(labels).cuda().view(-1))
        loss.backward()
        got = logits.detach().cpu().numpy()
        ref, cache, newbuf = orc.forward_train(sd, x)
        emu = emulate(sd, x)
        s = np.abs(ref).max()
        print(f"== C={C} B={B} N={N}: |ref|max {s:.4f}  cuda-vs-ref {np.abs(got-ref).max()/s:.4f}  emu-vs-ref {np.abs(emu-ref).max()/s:.4f}  cuda-vs-emu {np.abs(got-emu).max()/s:.4f}")
        rl, dlog = orc.weighted_ce(ref, labels, cw)
        print(f"   loss cuda {loss.item():.6f} ref {rl:.6f}")
        grads = orc.backward(cache, dlog)
        for name, p in m.named_parameters():
            g = p.grad.detach().cpu().numpy().astype(np.float64)
            r = np.asarray(grads[name]).reshape(g.shape)
            sc = np.abs(r).max()
            cos = (g * r).sum() / (np.linalg.norm(g) * np.linalg.norm(r) + 1e-30)
            print(f"   {name:22s} |ref|max {sc:.3e} relerr {np.abs(g-r).max()/(sc+1e-30):.3e} cos {cos:.5f}")
        for name, b in m.named_buffers():
            if "num_batches" in name:
                continue
            r = newbuf[name]
            print(f"   buf {name:26s} maxabs err {np.abs(b.cpu().numpy()-r).max():.3e}")


if __name__ == "__main__":
    main()
