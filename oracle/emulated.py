"""The reference training forward (pcs.py:98-133 under model.train()) with the CUDA path's STORAGE ROUNDING emulated.
TEST INFRASTRUCTURE ONLY (same rules as pointnet_oracle.py).

Train-mode BatchNorm followed by the global arg-max amplifies rounding: against the exact (fp64) oracle the bf16 CUDA path
differs by ~0.1 x max|logit| and trunk-gradient cosine ~0.9 although every kernel is right to one rounding
(tests/test_layerwise_gpu.py).  This module rounds to bf16 exactly where the CUDA path does and is exact (fp64) everywhere
else, so that an END-TO-END comparison can be tight: what remains are isolated rounding flips between the GPU's fp32
accumulation and fp64.  The cache it returns is the one `pointnet_oracle.backward` consumes, i.e. gradients are the exact
autograd gradients of the emulated forward (backward-side rounding has no measurable effect).

Rounding points of the CUDA training forward:
  conv1..conv4, seg_conv1..3 : weights bf16 (conv1: fp32), pre-BN output y stored bf16, batch statistics of the stored y,
                               activation relu(bn(y)) stored bf16 (seg_conv3's activation stays fp32 inside the head kernel)
  conv5 (folded)             : y never stored: statistics of the fp32 accumulators (Gram-predicted), activation stored bf16
  global_feat                : statistics and arg-max of the bf16-rounded output, pooled value from the rounded extremum
  conv biases                : dropped (train-mode BatchNorm cancels them exactly)
"""
import numpy as np

from . import pointnet_oracle as orc
from .layerwise import bf16_round

FOLDED_LAYERS = ("conv5",)          # y not rounded before its statistics / BatchNorm


def forward_train_emulated(sd, x, folded=True):
    B, N, cin = x.shape
    P = B * N
    a = np.asarray(x, np.float64).reshape(P, cin)
    cache = {"B": B, "N": N, "layers": {}}
    r = bf16_round

    def block(a_in, conv, bn, first=False, extra=None, wcols=None, round_act=True):
        W = sd[f"{conv}.weight"][:, :, 0].astype(np.float64)
        if wcols is not None:
            W = W[:, wcols]
        Wm = W if first else r(W)
        y = a_in @ Wm.T
        if extra is not None:
            y = y + extra
        if not (folded and conv in FOLDED_LAYERS):
            y = r(y)
        mean, var = y.mean(0), y.var(0)
        invstd = 1.0 / np.sqrt(var + orc.BN_EPS)
        yhat = (y - mean) * invstd
        g = sd[f"{bn}.weight"].astype(np.float64)
        z = yhat * g + sd[f"{bn}.bias"].astype(np.float64)
        out = np.maximum(z, 0)
        if round_act:
            out = r(out)
        return out, dict(a_in=a_in, W=Wm, b=np.zeros(W.shape[0]), bn=(yhat, invstd, g), relu=(z > 0), drop=None)

    a, cache["layers"]["conv1"] = block(a, "conv1", "bn1", first=True)
    a, cache["layers"]["conv2"] = block(a, "conv2", "bn2")
    pf = a
    for conv, bn in (("conv3", "bn3"), ("conv4", "bn4"), ("conv5", "bn5")):
        a, cache["layers"][conv] = block(a, conv, bn)
    a6, c6 = block(a, "global_feat", "bn_global", round_act=False)
    cache["layers"]["global_feat"] = c6
    a3 = a6.reshape(B, N, -1)
    arg = a3.argmax(axis=1)                       # first maximum (pcs.py:114); monotone in the rounded pre-BN value
    g = np.take_along_axis(a3, arg[:, None, :], axis=1)[:, 0, :]
    cache["argmax"] = arg
    Wg = sd["seg_conv1.weight"][:, 64:, 0].astype(np.float64)        # fp32 in the kernel
    cb = g @ Wg.T
    a, c = block(pf, "seg_conv1", "bn_seg1", extra=np.repeat(cb, N, axis=0), wcols=slice(0, 64))
    # pointnet_oracle.backward expects the concatenated operand [point_feat | repeated pooled feature] and the full weight
    c["a_in"] = np.concatenate([pf, np.repeat(g, N, axis=0)], axis=1)
    c["W"] = np.concatenate([c["W"], Wg], axis=1)
    cache["layers"]["seg_conv1"] = c
    a, cache["layers"]["seg_conv2"] = block(a, "seg_conv2", "bn_seg2")
    a, cache["layers"]["seg_conv3"] = block(a, "seg_conv3", "bn_seg3", round_act=False)
    W4 = sd["seg_conv4.weight"][:, :, 0].astype(np.float64)
    cache["layers"]["seg_conv4"] = dict(a_in=a, W=W4)
    logits = (a @ W4.T + sd["seg_conv4.bias"].astype(np.float64)).reshape(B, N, -1)
    return logits, cache
