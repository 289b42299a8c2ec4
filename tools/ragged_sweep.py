"""Valid-point-fraction sweep of a zero-padded ragged batch (the substitute for BASELINE configs[4]'s
occupancy sweep): padded (dense) execution against ragged execution of the SAME batch, inference and training step.

    python tools/ragged_sweep.py [--B 8 --N 16384 --steps 20 --warmup 5] > gpurun_out/ragged_sweep.json

Prints one JSON object: for every fraction the device time per step (CUDA events) and valid / total points per second.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
os.environ.setdefault("PCSEG_RAGGED_MIN_PAD", "0")      # measure the packed step itself (the trainer routes mostly-full batches to the dense path)
import pcseg_b200  # noqa: E402


def lengths_for(B, N, frac, rng):
    """one cloud is full (collate_fn pads to the longest cloud, pcs.py:50), the others are drawn so that the batch holds
    about frac * B * N real points"""
    if frac >= 1.0:
        return [N] * B
    rest = (frac * B - 1.0) / (B - 1) if B > 1 else frac
    rest = min(max(rest, 1.0 / N), 1.0)
    lo, hi = max(rest - min(rest, 1 - rest) * 0.5, 0.0), min(rest + min(rest, 1 - rest) * 0.5, 1.0)
    ls = [N] + [max(1, int(round(N * rng.uniform(lo, hi)))) for _ in range(B - 1)]
    return ls


def timed(fn, steps, warmup):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=8)
    ap.add_argument("--N", type=int, default=16384)
    ap.add_argument("--C", type=int, default=5)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--fracs", default="1.0,0.9,0.75,0.5,0.25")
    args = ap.parse_args()
    B, N, C = args.B, args.N, args.C
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(0)
    torch.manual_seed(1234)
    rows = []
    for frac in [float(f) for f in args.fracs.split(",")]:
        lengths = lengths_for(B, N, frac, rng)
        valid = sum(lengths)
        x = torch.rand(B, N, 4, device=dev)
        y = torch.randint(0, C, (B, N), device=dev)
        for b, L in enumerate(lengths):
            x[b, L:] = 0
            y[b, L:] = -1
        rec = {"target_frac": frac, "valid_frac": valid / (B * N), "valid_points": valid, "padded_points": B * N}
        # ---- inference
        model = pcseg_b200.PointNetSegmentation(C).to(dev).eval()
        with torch.no_grad():
            same = bool(torch.equal(model(x), model(x, lengths=lengths)))
            ms_d = timed(lambda: model(x), args.steps, args.warmup)
            ms_r = timed(lambda: model(x, lengths=lengths), args.steps, args.warmup)
        rec["eval"] = {"bit_identical": same, "padded_ms": ms_d, "ragged_ms": ms_r, "speedup": ms_d / ms_r,
                       "padded_valid_pts_per_s": valid / (ms_d * 1e-3), "ragged_valid_pts_per_s": valid / (ms_r * 1e-3),
                       "ragged_total_pts_per_s": B * N / (ms_r * 1e-3)}
        # ---- training step (forward + weighted CE + backward + Adam); padded = CUDA-graph replay, ragged = eager launches
        res = {}
        for name, ls in (("padded", None), ("ragged", lengths)):
            torch.manual_seed(1234)
            m = pcseg_b200.PointNetSegmentation(C).to(dev).train()
            tr = pcseg_b200.FusedTrainer(m, class_weights=torch.ones(C, device=dev), lr=1e-3, weight_decay=1e-4, device=dev)
            ms = timed(lambda: tr.step(x, y, lengths=ls), args.steps, max(args.warmup, 4))
            res[name] = (ms, float(tr.step(x, y, lengths=ls)["loss"].item()))
            del tr, m
        rec["train"] = {"padded_ms": res["padded"][0], "ragged_ms": res["ragged"][0], "speedup": res["padded"][0] / res["ragged"][0],
                        "padded_valid_pts_per_s": valid / (res["padded"][0] * 1e-3),
                        "ragged_valid_pts_per_s": valid / (res["ragged"][0] * 1e-3),
                        "ragged_total_pts_per_s": B * N / (res["ragged"][0] * 1e-3),
                        "loss_after_steps": {"padded": res["padded"][1], "ragged": res["ragged"][1]}}
        rows.append(rec)
        print(f"frac {rec['valid_frac']:.3f}: eval {ms_d:.3f} -> {ms_r:.3f} ms ({ms_d / ms_r:.2f}x, identical={same}); "
              f"train {res['padded'][0]:.3f} -> {res['ragged'][0]:.3f} ms ({res['padded'][0] / res['ragged'][0]:.2f}x)", file=sys.stderr)
    print(json.dumps({"workload": f"zero-padded batch {B} x {N}, C={C}, dropout 0.3", "steps": args.steps, "sweep": rows}))


if __name__ == "__main__":
    main()
