"""Generate golden vectors by executing the UNMODIFIED reference module.

Run in the build container only (it reads /root/reference, which does not exist
on the GPU box):

    python tests/golden/make_golden.py

The reference imports h5py at module scope (pcs.py:6) but only touches it inside
PointCloudDataset.__init__ (pcs.py:22-23), which the hot path never calls, so an
empty stub module is enough.

Weights come from `oracle.pointnet_oracle.synth_state(seed)` (numpy-only, so tests
can regenerate them without torch RNG); only inputs, outputs and gradient
digests are stored to keep the fixtures small.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.modules.setdefault("h5py", types.ModuleType("h5py"))
sys.path.insert(0, "/root/reference")
import point_cloud_segmentation as pcs  # noqa: E402

from oracle import pointnet_oracle as orc  # noqa: E402


def grad_digest(g: np.ndarray):
    """Small fingerprint of a gradient tensor: sum, abs-sum, and a strided sample."""
    flat = g.reshape(-1).astype(np.float64)
    idx = np.linspace(0, flat.size - 1, num=min(flat.size, 64)).astype(np.int64)
    return np.concatenate([[flat.sum(), np.abs(flat).sum()], flat[idx]])


def make_case(tag, C, B, N, seed, lens=None, class_w=None):
    torch.manual_seed(seed)
    torch.set_num_threads(1)
    sd_np = orc.synth_state(C, seed)
    model = pcs.PointNetSegmentation(num_classes=C)
    model.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd_np.items()}, strict=True)

    rng = np.random.default_rng(seed + 1)
    if lens is None:
        x = rng.random((B, N, 4), dtype=np.float32)
        labels = rng.integers(0, C, (B, N)).astype(np.int64)
    else:  # ragged clouds through the reference collate_fn (pcs.py:44-63)
        batch = [(torch.from_numpy(rng.random((n, 4), dtype=np.float32)),
                  torch.from_numpy(rng.integers(0, C, (n,)).astype(np.int64))) for n in lens]
        xt, lt, _ = pcs.collate_fn(batch)
        x, labels = xt.numpy(), lt.numpy()
    cw = np.ones(C, np.float32) if class_w is None else np.asarray(class_w, np.float32)

    out = {"x": x, "labels": labels, "class_w": cw, "C": np.int64(C), "seed": np.int64(seed)}
    xt = torch.from_numpy(x)

    # eval forward exactly as pcs.py:450-452
    model.eval()
    with torch.no_grad():
        le = model(xt)
        out["eval_logits"] = le.numpy().copy()
        out["eval_argmax"] = torch.argmax(le, dim=2).numpy().copy()

    # train step pieces exactly as pcs.py:241-254, dropout forced to p=0 (RNG parity is not a goal)
    model.train()
    model.dropout.p = 0.0
    crit = torch.nn.CrossEntropyLoss(ignore_index=-1, weight=torch.from_numpy(cw))
    model.zero_grad()
    lt_ = model(xt)
    loss = crit(lt_.contiguous().view(-1, C), torch.from_numpy(labels).contiguous().view(-1))
    loss.backward()
    out["train_logits"] = lt_.detach().numpy().copy()
    out["loss"] = np.float64(loss.item())
    for name, p in model.named_parameters():
        g = p.grad.numpy()
        out["gd/" + name] = grad_digest(g)
        if g.size <= 8192:
            out["g/" + name] = g.copy()
    for name, b in model.named_buffers():
        out["buf/" + name] = b.numpy().copy()
    np.savez_compressed(os.path.join(HERE, f"{tag}.npz"), **out)
    print(tag, "loss", out["loss"], "eval logits absmax", np.abs(out["eval_logits"]).max())


if __name__ == "__main__":
    make_case("case_dense_c5", C=5, B=2, N=96, seed=11, class_w=[0.5, 1.0, 2.0, 0.75, 0.75])
    make_case("case_ragged_c3", C=3, B=3, N=None, seed=23, lens=[70, 33, 128], class_w=[1.0, 0.6, 1.4])
    make_case("case_single_c5", C=5, B=1, N=200, seed=5)
