"""CPU oracle for the point-cloud segmentation hot path.  TEST INFRASTRUCTURE ONLY.

This file is a numpy restatement of the algorithm in the reference
`point_cloud_segmentation.py` (abbreviated pcs.py below).  It exists so that the
CUDA path can be checked against something that is (a) independent of the CUDA
code and (b) pinned to the real reference: `tests/golden/make_golden.py` runs the
unmodified reference module in the build container and the resulting vectors are
replayed against this oracle by `tests/test_oracle_golden.py`.

Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference`
legs of `bench.py` may import this module.  The product package never does; it
fails loudly when the CUDA library is missing.

Parity status: PINNED against the reference module executed in the build
container (the reference ships no tests or golden vectors of its own).

Layout convention: activations are kept point-major, `(P, C)` with `P = B*N`
rows, i.e. the transpose of the reference's `(B, C, N)` Conv1d layout
(pcs.py:103).  A `Conv1d(kernel_size=1)` is then a plain matrix product.
"""
from __future__ import annotations

import numpy as np

BN_EPS = 1e-5          # nn.BatchNorm1d default, pcs.py:86-94
BN_MOMENTUM = 0.1      # nn.BatchNorm1d default, pcs.py:86-94
DROPOUT_P = 0.3        # pcs.py:96

# (conv name, bn name or None, Cin, Cout) in execution order, pcs.py:70-94,106-128
TRUNK = [
    ("conv1", "bn1", 4, 64),
    ("conv2", "bn2", 64, 64),
    ("conv3", "bn3", 64, 64),
    ("conv4", "bn4", 64, 128),
    ("conv5", "bn5", 128, 1024),
    ("global_feat", "bn_global", 1024, 1024),
]
HEAD = [
    ("seg_conv1", "bn_seg1", 1088, 512),
    ("seg_conv2", "bn_seg2", 512, 256),
    ("seg_conv3", "bn_seg3", 256, 128),
]
CONV_NAMES = [c for c, _, _, _ in TRUNK] + [c for c, _, _, _ in HEAD] + ["seg_conv4"]
BN_NAMES = [b for _, b, _, _ in TRUNK] + [b for _, b, _, _ in HEAD]


def state_dict_spec(num_classes: int, input_dim: int = 4):
    """Names, shapes and dtypes of the 65 state_dict entries in registration order
    (convs pcs.py:70-83, then batch norms pcs.py:86-94)."""
    spec = []
    shapes = {c: (co, ci) for c, _, ci, co in TRUNK + HEAD}
    shapes["conv1"] = (64, input_dim)
    shapes["seg_conv4"] = (num_classes, 128)
    for c in CONV_NAMES:
        co, ci = shapes[c]
        spec.append((f"{c}.weight", (co, ci, 1), np.float32))
        spec.append((f"{c}.bias", (co,), np.float32))
    for (_, b, _, co) in TRUNK + HEAD:
        spec.append((f"{b}.weight", (co,), np.float32))
        spec.append((f"{b}.bias", (co,), np.float32))
        spec.append((f"{b}.running_mean", (co,), np.float32))
        spec.append((f"{b}.running_var", (co,), np.float32))
        spec.append((f"{b}.num_batches_tracked", (), np.int64))
    return spec


def synth_state(num_classes: int, seed: int, input_dim: int = 4):
    """Deterministic, torch-free synthetic weights (best_model.pth is not shipped,
    /root/reference/.MISSING_LARGE_BLOBS).  BN affine/running stats are randomised
    with mixed-sign gamma so that folding and the max-pool sign handling are
    exercised."""
    rng = np.random.default_rng(seed)
    sd = {}
    for name, shape, dt in state_dict_spec(num_classes, input_dim):
        leaf = name.split(".")[-1]
        if leaf == "num_batches_tracked":
            sd[name] = np.array(7, dtype=np.int64)
        elif name.split(".")[0] in BN_NAMES:
            if leaf == "weight":
                sd[name] = rng.uniform(-1.5, 1.5, shape).astype(np.float32)
            elif leaf == "bias":
                sd[name] = (0.2 * rng.standard_normal(shape)).astype(np.float32)
            elif leaf == "running_mean":
                sd[name] = (0.5 * rng.standard_normal(shape)).astype(np.float32)
            else:
                sd[name] = rng.uniform(0.5, 2.0, shape).astype(np.float32)
        else:
            fan_in = shape[1] if leaf == "weight" else None
            if leaf == "weight":
                bound = 1.0 / np.sqrt(fan_in)
                sd[name] = rng.uniform(-bound, bound, shape).astype(np.float32)
            else:
                sd[name] = rng.uniform(-0.1, 0.1, shape).astype(np.float32)
    return sd


def _w(sd, conv, dt):
    return sd[f"{conv}.weight"][:, :, 0].astype(dt), sd[f"{conv}.bias"].astype(dt)


def _bn_eval(y, sd, bn, dt):
    """BatchNorm1d in eval mode: running statistics, pcs.py:277,432."""
    g = sd[f"{bn}.weight"].astype(dt)
    b = sd[f"{bn}.bias"].astype(dt)
    m = sd[f"{bn}.running_mean"].astype(dt)
    v = sd[f"{bn}.running_var"].astype(dt)
    return (y - m) / np.sqrt(v + BN_EPS) * g + b


def forward_eval(sd, x, dtype=np.float64, pooled=None, return_pooled=False):
    """Inference forward, pcs.py:98-133 under model.eval() (pcs.py:432,450-451).
    x: (B, N, Cin) -> logits (B, N, num_classes).
    pooled / return_pooled: the (B, 1024) max-pool result of pcs.py:114 can be handed in / out, which is all that
    point-sharded inference exchanges (tests of the multi-GPU protocol)."""
    B, N, cin = x.shape
    a = x.reshape(B * N, cin).astype(dtype)
    point_feat = None
    for conv, bn, _, _ in TRUNK:
        W, b = _w(sd, conv, dtype)
        a = np.maximum(_bn_eval(a @ W.T + b, sd, bn, dtype), 0)       # pcs.py:106-113
        if conv == "conv2":
            point_feat = a                                              # pcs.py:107
    g = a.reshape(B, N, -1).max(axis=1)                                # pcs.py:114
    g_local = g
    if pooled is not None:
        g = np.asarray(pooled, dtype)
    gexp = np.repeat(g[:, None, :], N, axis=1).reshape(B * N, -1)      # pcs.py:117
    a = np.concatenate([point_feat, gexp], axis=1)                     # pcs.py:120
    for conv, bn, _, _ in HEAD:
        W, b = _w(sd, conv, dtype)
        a = np.maximum(_bn_eval(a @ W.T + b, sd, bn, dtype), 0)       # pcs.py:123-127 (dropout = identity in eval)
    W, b = _w(sd, "seg_conv4", dtype)
    logits = (a @ W.T + b).reshape(B, N, -1)                            # pcs.py:128-131
    return (logits, g_local) if return_pooled else logits


def _bn_train(y, sd, bn, dt):
    """BatchNorm1d in train mode over all B*N rows, padded rows included
    (pcs.py:230 + SURVEY §8 row P).  Returns output, cache and updated buffers."""
    n = y.shape[0]
    mean = y.mean(axis=0)
    var_b = y.var(axis=0)                      # biased: used to normalise
    invstd = 1.0 / np.sqrt(var_b + BN_EPS)
    yhat = (y - mean) * invstd
    g = sd[f"{bn}.weight"].astype(dt)
    b = sd[f"{bn}.bias"].astype(dt)
    out = yhat * g + b
    var_u = var_b * n / max(n - 1, 1)          # unbiased: goes into running_var
    new = {
        f"{bn}.running_mean": (1 - BN_MOMENTUM) * sd[f"{bn}.running_mean"].astype(dt) + BN_MOMENTUM * mean,
        f"{bn}.running_var": (1 - BN_MOMENTUM) * sd[f"{bn}.running_var"].astype(dt) + BN_MOMENTUM * var_u,
        f"{bn}.num_batches_tracked": sd[f"{bn}.num_batches_tracked"] + 1,
    }
    return out, (yhat, invstd, g), new


def forward_train(sd, x, dropout_masks=None, dtype=np.float64):
    """Training forward, pcs.py:98-133 under model.train() (pcs.py:230).
    dropout_masks: None (p forced to 0) or dict {'seg1': (P,512) 0/1, 'seg2': (P,256) 0/1}
    keep-masks; kept activations are scaled by 1/(1-p) (pcs.py:96,124,126).
    Returns logits (B,N,C), cache for `backward`, dict of updated BN buffers."""
    B, N, cin = x.shape
    P = B * N
    a = x.reshape(P, cin).astype(dtype)
    cache = {"B": B, "N": N, "layers": {}}
    new_buffers = {}
    point_feat = None

    def block(a_in, conv, bn, drop=None):
        W, b = _w(sd, conv, dtype)
        y = a_in @ W.T + b
        z, bnc, nb = _bn_train(y, sd, bn, dtype)
        new_buffers.update(nb)
        out = np.maximum(z, 0)
        scale = None
        if drop is not None:
            scale = drop.astype(dtype) / (1.0 - DROPOUT_P)
            out = out * scale
        cache["layers"][conv] = dict(a_in=a_in, W=W, b=b, bn=bnc, relu=(z > 0), drop=scale)
        return out

    for conv, bn, _, _ in TRUNK:
        a = block(a, conv, bn)
        if conv == "conv2":
            point_feat = a
    a3 = a.reshape(B, N, -1)
    arg = a3.argmax(axis=1)                                             # (B, 1024) first max, pcs.py:114
    g = np.take_along_axis(a3, arg[:, None, :], axis=1)[:, 0, :]
    cache["argmax"] = arg
    gexp = np.repeat(g[:, None, :], N, axis=1).reshape(P, -1)
    a = np.concatenate([point_feat, gexp], axis=1)
    masks = dropout_masks or {}
    a = block(a, "seg_conv1", "bn_seg1", masks.get("seg1"))
    a = block(a, "seg_conv2", "bn_seg2", masks.get("seg2"))
    a = block(a, "seg_conv3", "bn_seg3")
    W, b = _w(sd, "seg_conv4", dtype)
    cache["layers"]["seg_conv4"] = dict(a_in=a, W=W)
    logits = (a @ W.T + b).reshape(B, N, -1)
    return logits, cache, new_buffers


def weighted_ce(logits, labels, class_w, dtype=np.float64):
    """nn.CrossEntropyLoss(ignore_index=-1, weight=class_w), mean reduction
    (pcs.py:216, applied at pcs.py:247-251).  Returns (loss, dlogits)."""
    C = logits.shape[-1]
    z = logits.reshape(-1, C).astype(dtype)
    y = labels.reshape(-1)
    valid = y != -1
    ys = np.where(valid, y, 0)
    zmax = z.max(axis=1, keepdims=True)
    lse = zmax[:, 0] + np.log(np.exp(z - zmax).sum(axis=1))
    logp_y = z[np.arange(z.shape[0]), ys] - lse
    w = np.where(valid, np.asarray(class_w, dtype)[ys], 0.0)
    wsum = w.sum()
    loss = -(w * logp_y).sum() / wsum
    soft = np.exp(z - lse[:, None])
    onehot = np.zeros_like(z)
    onehot[np.arange(z.shape[0]), ys] = 1.0
    dz = (soft - onehot) * (w / wsum)[:, None]
    return loss, dz.reshape(logits.shape)


def backward(cache, dlogits, taps=None):
    """Gradients of every parameter for `forward_train` (what autograd computes at
    pcs.py:254).  Returns dict name -> grad with state_dict shapes.  `taps` (optional dict) receives
    intermediate gradients: 'dg' (wrt the pooled feature) and 'dx:<conv>' (wrt the input of a conv)."""
    B, N = cache["B"], cache["N"]
    P = B * N
    L = cache["layers"]
    grads = {}

    def conv_bwd(conv, dy, need_dx=True):
        c = L[conv]
        grads[f"{conv}.weight"] = (dy.T @ c["a_in"])[:, :, None]
        grads[f"{conv}.bias"] = dy.sum(axis=0)
        dx = dy @ c["W"] if need_dx else None
        if taps is not None and need_dx:
            taps[f"dx:{conv}"] = dx
        return dx

    def act_bwd(conv, bn, da):
        c = L[conv]
        if c["drop"] is not None:
            da = da * c["drop"]
        dz = da * c["relu"]
        yhat, invstd, g = c["bn"]
        grads[f"{bn}.weight"] = (dz * yhat).sum(axis=0)
        grads[f"{bn}.bias"] = dz.sum(axis=0)
        return g * invstd * (dz - dz.mean(axis=0) - yhat * (dz * yhat).mean(axis=0))

    dz4 = dlogits.reshape(P, -1).astype(L["seg_conv4"]["W"].dtype)
    da = conv_bwd("seg_conv4", dz4)
    for conv, bn in (("seg_conv3", "bn_seg3"), ("seg_conv2", "bn_seg2"), ("seg_conv1", "bn_seg1")):
        da = conv_bwd(conv, act_bwd(conv, bn, da))
    d_point_feat = da[:, :64]                                           # cat split, pcs.py:120
    dg = da[:, 64:].reshape(B, N, -1).sum(axis=1)                       # repeat bwd, pcs.py:117
    if taps is not None:
        taps["dg"] = dg
    da6 = np.zeros((B, N, dg.shape[1]), dtype=da.dtype)                 # max bwd, pcs.py:114
    np.put_along_axis(da6, cache["argmax"][:, None, :], dg[:, None, :], axis=1)
    da = da6.reshape(P, -1)
    for conv, bn in (("global_feat", "bn_global"), ("conv5", "bn5"), ("conv4", "bn4"), ("conv3", "bn3")):
        da = conv_bwd(conv, act_bwd(conv, bn, da))
    da = da + d_point_feat                                              # skip join at point_feat, pcs.py:107
    da = conv_bwd("conv2", act_bwd("conv2", "bn2", da))
    conv_bwd("conv1", act_bwd("conv1", "bn1", da), need_dx=False)       # input has no grad
    return grads


def argmax_labels(logits):
    """torch.max(outputs, 2)[1] / torch.argmax(outputs, dim=2), pcs.py:261,452."""
    return logits.argmax(axis=-1).astype(np.int64)


def adam_step(params, grads, m, v, step, lr=1e-3, b1=0.9, b2=0.999, eps=1e-8, wd=1e-4):
    """torch.optim.Adam(lr=1e-3, weight_decay=1e-4) single step (pcs.py:217,255):
    L2-style decay added to the gradient, bias-corrected moments."""
    out = {}
    for k in params:
        g = grads[k] + wd * params[k]
        m[k] = b1 * m[k] + (1 - b1) * g
        v[k] = b2 * v[k] + (1 - b2) * g * g
        mh = m[k] / (1 - b1 ** step)
        vh = v[k] / (1 - b2 ** step)
        out[k] = params[k] - lr * mh / (np.sqrt(vh) + eps)
    return out


def collate(points_list, labels_list):
    """Zero-pad ragged clouds to max_points; labels padded with -1; bool mask
    (pcs.py:44-63)."""
    B = len(points_list)
    n = max(p.shape[0] for p in points_list)
    pts = np.zeros((B, n, 4), np.float32)
    lab = np.full((B, n), -1, np.int64)
    msk = np.zeros((B, n), bool)
    for i, (p, l) in enumerate(zip(points_list, labels_list)):
        pts[i, : p.shape[0]] = p
        lab[i, : p.shape[0]] = l
        msk[i, : p.shape[0]] = True
    return pts, lab, msk
