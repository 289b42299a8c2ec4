"""End-to-end training parity report: CUDA step (folded and legacy) vs the fp64 oracle, per-tensor gradient cosine.
    python tools/train_parity_report.py [B N C]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import pointnet_oracle as orc  # noqa: E402


def run(folded, C, sd, x, labels, cw):
    os.environ["PCSEG_FOLDED"] = "1" if folded else "0"
    import pcseg_b200
    m = pcseg_b200.PointNetSegmentation(C)
    m.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()})
    m = m.cuda().train()
    m.dropout.p = 0.0
    crit = torch.nn.CrossEntropyLoss(ignore_index=-1, weight=torch.from_numpy(cw).cuda())
    logits = m(torch.from_numpy(x).cuda())
    loss = crit(logits.contiguous().view(-1, C), torch.from_numpy(labels).cuda().view(-1))
    loss.backward()
    return logits.detach().cpu().numpy(), float(loss), {n: p.grad.detach().cpu().numpy().astype(np.float64) for n, p in m.named_parameters()}


def main():
    B, N, C = (int(v) for v in sys.argv[1:4]) if len(sys.argv) >= 4 else (8, 2048, 5)
    sd = orc.synth_state(C, 7 * B + N)
    rng = np.random.default_rng(B * N)
    x = rng.random((B, N, 4), dtype=np.float32)
    labels = rng.integers(0, C, (B, N)).astype(np.int64)
    cw = (0.5 + rng.random(C)).astype(np.float32)
    ref_logits, cache, _ = orc.forward_train(sd, x)
    ref_loss, dlog = orc.weighted_ce(ref_logits, labels, cw)
    ref = orc.backward(cache, dlog)
    res = {f: run(f, C, sd, x, labels, cw) for f in (True, False)}
    s = np.abs(ref_logits).max()
    for f in (True, False):
        lg, loss, _ = res[f]
        d = np.abs(lg - ref_logits)
        print(f"folded={f}: loss {loss:.6f} (oracle {ref_loss:.6f}), logits max err {d.max() / s:.4f} rms {np.sqrt((d * d).mean()) / s:.4f} of max|logit|")
    print(f"{'tensor':24s} {'cos folded':>11s} {'cos legacy':>11s} {'cos f-vs-l':>11s}")
    for name, r in ref.items():
        if name.endswith(".bias") and name.split(".")[0] in orc.CONV_NAMES[:-1]:
            continue
        r = np.asarray(r, np.float64).reshape(-1)

        def cos(a, b):
            return float((a * b).sum() / (np.linalg.norm(a) * np.linalg.norm(b) + 1e-30))
        gf, gl = res[True][2][name].reshape(-1), res[False][2][name].reshape(-1)
        print(f"{name:24s} {cos(gf, r):11.5f} {cos(gl, r):11.5f} {cos(gf, gl):11.5f}")


if __name__ == "__main__":
    main()
