"""fp64 restatement of the RAGGED (un-padded) execution scheme of the CUDA path, used by the CPU tests to
show that it reproduces the reference's padded batch exactly.  TEST INFRASTRUCTURE ONLY — the product never imports it.

The reference pads every cloud of a batch with zero rows up to the longest one (collate_fn, pcs.py:44-63) and feeds the
pad rows through the network like any other row.  All pad rows of one cloud are identical, so the packed
scheme keeps the real rows plus ONE representative pad row per padded cloud and gives that row the multiplicity
`N - len`:
  * forward: every sum over points (BatchNorm batch statistics, pcs.py:86-94 in train mode) weights a row by its
    multiplicity and normalises by the logical row count B*N; the max-pool (pcs.py:114) sees the representative row once
    (it is the first pad row, so `first maximum wins` is unchanged);
  * backward: the gradient of a row is carried PRE-MULTIPLIED by its multiplicity (every backward operator is linear in
    it), so weight / bias / BN-parameter gradients and the per-cloud sums come out of the unmodified formulas; only the
    affine term of the BatchNorm backward is scaled by the multiplicity.
The filler rows of the CUDA layout (multiplicity 0, there only to align clouds to GEMM tiles) contribute nothing by
construction and are simply left out here.
"""
import numpy as np

from . import pointnet_oracle as orc


def pack(x, labels, lengths):
    """(B, N, 4) padded batch + per-cloud lengths -> packed rows, labels, multiplicities, cloud ids."""
    B, N, _ = x.shape
    rows, labs, mult, cloud = [], [], [], []
    for b, L in enumerate(lengths):
        rows.append(x[b, :L])
        labs.append(labels[b, :L])
        mult.append(np.ones(L))
        cloud.append(np.full(L, b))
        if L < N:                                   # one representative of the N - L identical pad rows
            rows.append(np.zeros((1, x.shape[2]), x.dtype))
            labs.append(np.array([-1]))
            mult.append(np.array([float(N - L)]))
            cloud.append(np.array([b]))
    return (np.concatenate(rows).astype(np.float64), np.concatenate(labs).astype(np.int64), np.concatenate(mult),
            np.concatenate(cloud).astype(np.int64))


def forward_train_packed(sd, xp, mult, cloud, B, N, dtype=np.float64):
    """Training forward (dropout off) on packed rows; returns logits (R, C) and the cache for `backward_packed`."""
    n = float(B * N)                                # BatchNorm normalises by the LOGICAL number of rows
    m = mult[:, None]
    cache = {"layers": {}, "mult": mult, "cloud": cloud, "B": B, "n": n}

    def block(a_in, conv, bn):
        W, b = orc._w(sd, conv, dtype)
        y = a_in @ W.T + b
        mean = (m * y).sum(axis=0) / n
        var = (m * (y - mean) ** 2).sum(axis=0) / n
        invstd = 1.0 / np.sqrt(var + orc.BN_EPS)
        yhat = (y - mean) * invstd
        g = sd[f"{bn}.weight"].astype(dtype)
        z = yhat * g + sd[f"{bn}.bias"].astype(dtype)
        cache["layers"][conv] = dict(a_in=a_in, W=W, bn=(yhat, invstd, g), relu=(z > 0))
        return np.maximum(z, 0)

    a = xp.astype(dtype)
    point_feat = None
    for conv, bn, _, _ in orc.TRUNK:
        a = block(a, conv, bn)
        if conv == "conv2":
            point_feat = a
    g = np.zeros((B, a.shape[1]), dtype)
    arg = np.zeros((B, a.shape[1]), np.int64)       # packed row index of the first maximum of every cloud / channel
    for b in range(B):
        idx = np.nonzero(cloud == b)[0]             # real rows first, then the representative pad row
        loc = a[idx].argmax(axis=0)
        arg[b] = idx[loc]
        g[b] = a[arg[b], np.arange(a.shape[1])]
    cache["argmax"] = arg
    a = np.concatenate([point_feat, g[cloud]], axis=1)
    for conv, bn, _, _ in orc.HEAD:
        a = block(a, conv, bn)
    W, b = orc._w(sd, "seg_conv4", dtype)
    cache["layers"]["seg_conv4"] = dict(a_in=a, W=W)
    return a @ W.T + b, cache


def backward_packed(cache, dlogits_packed):
    """Parameter gradients from the packed rows; `dlogits_packed` is the loss gradient per packed row (for a pad row: of ONE
    of the identical pad rows).  Row gradients are carried multiplied by the row multiplicity."""
    L = cache["layers"]
    m = cache["mult"][:, None]
    n = cache["n"]
    cloud, B = cache["cloud"], cache["B"]
    grads = {}

    def conv_bwd(conv, dy, need_dx=True):
        c = L[conv]
        grads[f"{conv}.weight"] = (dy.T @ c["a_in"])[:, :, None]
        grads[f"{conv}.bias"] = dy.sum(axis=0)
        return dy @ c["W"] if need_dx else None

    def act_bwd(conv, bn, da):
        c = L[conv]
        dz = da * c["relu"]
        yhat, invstd, g = c["bn"]
        s1, s2 = dz.sum(axis=0), (dz * yhat).sum(axis=0)
        grads[f"{bn}.weight"] = s2
        grads[f"{bn}.bias"] = s1
        return g * invstd * (dz - m * (s1 / n) - m * yhat * (s2 / n))     # affine term x multiplicity

    da = conv_bwd("seg_conv4", m * dlogits_packed)
    for conv, bn in (("seg_conv3", "bn_seg3"), ("seg_conv2", "bn_seg2"), ("seg_conv1", "bn_seg1")):
        da = conv_bwd(conv, act_bwd(conv, bn, da))
    d_point_feat = da[:, :64]
    dg = np.zeros((B, da.shape[1] - 64), da.dtype)
    np.add.at(dg, cloud, da[:, 64:])                                        # repeat backward: sum over the cloud's rows
    da6 = np.zeros((da.shape[0], dg.shape[1]), da.dtype)
    da6[cache["argmax"], np.arange(dg.shape[1])[None, :]] = dg              # max backward: to the first arg-max row
    da = da6
    for conv, bn in (("global_feat", "bn_global"), ("conv5", "bn5"), ("conv4", "bn4"), ("conv3", "bn3")):
        da = conv_bwd(conv, act_bwd(conv, bn, da))
    da = da + d_point_feat
    da = conv_bwd("conv2", act_bwd("conv2", "bn2", da))
    conv_bwd("conv1", act_bwd("conv1", "bn1", da), need_dx=False)
    return grads
