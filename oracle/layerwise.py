"""Layer-wise CPU oracle (TEST INFRASTRUCTURE ONLY, same rules as pointnet_oracle.py).

`pointnet_oracle.py` restates the reference end to end.  Train-mode BatchNorm followed by a
global arg-max makes the END-TO-END map ill-conditioned (tiny input-rounding differences flip
arg-max routes and ReLU masks; even TF32 — the reference's own CUDA default — only reaches
gradient cosine ~0.97 on random weights).  To prove the kernels right independently of that
amplification, every function here recomputes ONE step of the reference algorithm in fp64 from
the tensors the CUDA path actually produced for the previous step, so each kernel is compared on
identical inputs.  Each function cites the reference lines it follows (pcs.py =
point_cloud_segmentation.py).
"""
import numpy as np

BN_EPS = 1e-5


def bf16_round(x):
    """Round-to-nearest-even to bfloat16 precision, returned as float64."""
    f = np.ascontiguousarray(x, dtype=np.float32)
    u = f.view(np.uint32).astype(np.uint64)
    bias = ((u >> 16) & 1) + 0x7FFF
    u = ((u + bias) & 0xFFFF0000).astype(np.uint32)
    return u.view(np.float32).astype(np.float64).reshape(np.shape(x))


def conv_pre_bn(a_prev, W, cloud_bias=None, pts_per_cloud=None, weights_bf16=True):
    """Conv1d(k=1) without bias (train-mode BN cancels it), pcs.py:106-127: y = a_prev W^T (+ per-cloud term
    of the concat, pcs.py:117-123)."""
    Wm = bf16_round(W) if weights_bf16 else np.asarray(W, np.float64)
    y = np.asarray(a_prev, np.float64) @ Wm.T
    if cloud_bias is not None:
        y = y + np.repeat(np.asarray(cloud_bias, np.float64), pts_per_cloud, axis=0)
    return y


def bn_batch_stats(y):
    """Sum and sum of squares over all rows (pcs.py:86-94 in train mode; padded rows included)."""
    y = np.asarray(y, np.float64)
    return np.stack([y.sum(0), (y * y).sum(0)])


def bn_params(stats, n, gamma, beta):
    """{scale, shift, invstd, -mean*invstd} from batch sums: biased variance, eps 1e-5 (nn.BatchNorm1d)."""
    mean = stats[0] / n
    var = np.maximum(stats[1] / n - mean * mean, 0.0)
    invstd = 1.0 / np.sqrt(var + BN_EPS)
    g = np.asarray(gamma, np.float64)
    return np.stack([g * invstd, np.asarray(beta, np.float64) - mean * g * invstd, invstd, -mean * invstd], axis=1)


def running_stats(stats, n, conv_bias, rmean, rvar, momentum=0.1):
    """running_mean / running_var update with momentum 0.1 and unbiased variance (nn.BatchNorm1d)."""
    mean = stats[0] / n
    var = np.maximum(stats[1] / n - mean * mean, 0.0)
    unb = var * n / (n - 1) if n > 1 else var
    return ((1 - momentum) * rmean + momentum * (mean + conv_bias), (1 - momentum) * rvar + momentum * unb)


def bn_relu(y, bnp, keep=None, keep_scale=1.0):
    """relu(bn(y)) (pcs.py:106-127), optional dropout keep mask (pcs.py:124,126); fp32 fma like the kernel."""
    t = np.float32(bnp[:, 0]) * np.asarray(y, np.float32) + np.float32(bnp[:, 1])
    out = np.maximum(t.astype(np.float64), 0.0)
    if keep is not None:
        out = out * keep * keep_scale
    return out, (t > 0)


def maxpool(y6, bnp6, B, N):
    """torch.max over points of relu(bn(y6)) (pcs.py:114) via the per-channel monotonicity of BN:
    arg-extremum of y (max if scale >= 0 else min), first index on ties."""
    y = np.asarray(y6, np.float64).reshape(B, N, -1)
    sgn = np.where(bnp6[:, 0] >= 0, 1.0, -1.0)
    arg = (y * sgn).argmax(axis=1)                      # first occurrence
    ystar = np.take_along_axis(y, arg[:, None, :], 1)[:, 0, :]
    g = np.maximum(np.float32(bnp6[:, 0]) * ystar.astype(np.float32) + np.float32(bnp6[:, 1]), 0).astype(np.float64)
    return g, ystar, arg.astype(np.int32)


def head_logits(ys3, bnp, W4, b4):
    """seg_conv4(relu(bn_seg3(.))), pcs.py:127-128."""
    a, _ = bn_relu(ys3, bnp)
    return a @ np.asarray(W4, np.float64).T + np.asarray(b4, np.float64), a


def ce_grad(logits, labels, class_w, wsum):
    """d(weighted-mean CE)/dlogits with ignore_index=-1, pcs.py:216,251."""
    z = np.asarray(logits, np.float64)
    y = labels.reshape(-1)
    valid = y >= 0
    ys = np.where(valid, y, 0)
    zmax = z.max(1, keepdims=True)
    e = np.exp(z - zmax)
    soft = e / e.sum(1, keepdims=True)
    onehot = np.zeros_like(z)
    onehot[np.arange(z.shape[0]), ys] = 1
    w = np.where(valid, np.asarray(class_w, np.float64)[ys], 0.0)
    nll = -(np.log(soft[np.arange(z.shape[0]), ys]))
    return (soft - onehot) * (w / wsum)[:, None], float((w * nll).sum()), float(w.sum())


def bn_bwd_stats(dz, y, bnp):
    """sum dz and sum dz*yhat (the two reductions of BatchNorm backward)."""
    dz = np.asarray(dz, np.float64)
    yhat = np.asarray(y, np.float64) * bnp[:, 2] + bnp[:, 3]
    return np.stack([dz.sum(0), (dz * yhat).sum(0)])


def bn_bwd_coef(stats_b, n, bnp):
    """dy = A dz + Bc y + Cc with A = scale, Bc = -scale c2 invstd, Cc = -scale (c1 + c2 (-mean invstd))."""
    c1, c2 = stats_b[0] / n, stats_b[1] / n
    A = bnp[:, 0]
    return np.stack([A, -A * c2 * bnp[:, 2], -A * (c1 + c2 * bnp[:, 3])], axis=1)


def bn_bwd_apply(dz, y, coef):
    return coef[:, 0] * np.asarray(dz, np.float64) + coef[:, 1] * np.asarray(y, np.float64) + coef[:, 2]


def dgrad_masked(dy, W, y_prev, bnp_prev, keep=None, keep_scale=1.0):
    """(dy W) * relu'(bn(y_prev)) (* dropout): conv backward-data + ReLU backward of autograd (pcs.py:254)."""
    da = np.asarray(dy, np.float64) @ bf16_round(W)
    t = np.float32(bnp_prev[:, 0]) * np.asarray(y_prev, np.float32) + np.float32(bnp_prev[:, 1])
    dz = da * (t > 0)
    if keep is not None:
        dz = dz * keep * keep_scale
    return dz


def wgrad(dy, a_prev):
    """conv backward-weight: dW = dy^T a_prev."""
    return np.asarray(dy, np.float64).T @ np.asarray(a_prev, np.float64)


# ---------------------------------------------------------------------------------------------
# Conditioning of the train-mode forward under storage rounding: the same network with weights and the
# stored tensors (pre-BN outputs y, activations a) rounded to a given number of significand bits, everything else fp64.
# 8 bits = bf16 (what the CUDA path stores), 11 bits = TF32 (the reference's own cuDNN default), 24 bits = fp32.
# ---------------------------------------------------------------------------------------------
def round_to_bits(bits):
    def f(a):
        m, e = np.frexp(np.asarray(a, np.float64))
        return np.ldexp(np.round(m * 2.0 ** bits) / 2.0 ** bits, e)
    return f


def forward_rounded(sd, x, bits, train=True):
    """pcs.py:98-133 with storage rounding to `bits` significand bits (conv1 runs on fp32 FMA units: its weights are
    not rounded).  train=True: batch statistics (dropout off); train=False: running statistics."""
    from . import pointnet_oracle as orc
    r = round_to_bits(bits)
    B, N, _ = x.shape
    a = x.reshape(B * N, -1).astype(np.float64)

    def block(a_in, conv, bn, first=False, extra=None, wcols=None):
        W = sd[f"{conv}.weight"][:, :, 0].astype(np.float64)
        if wcols is not None:
            W = W[:, wcols]
        if not first:
            W = r(W)
        y = a_in @ W.T + sd[f"{conv}.bias"].astype(np.float64)
        if extra is not None:
            y = y + extra
        y = r(y)
        if train:
            mean, var = y.mean(0), y.var(0)
        else:
            mean, var = sd[f"{bn}.running_mean"].astype(np.float64), sd[f"{bn}.running_var"].astype(np.float64)
        return (y - mean) / np.sqrt(var + orc.BN_EPS) * sd[f"{bn}.weight"] + sd[f"{bn}.bias"]

    a = r(np.maximum(block(a, "conv1", "bn1", first=True), 0))
    pf = a = r(np.maximum(block(a, "conv2", "bn2"), 0))
    for conv, bn in (("conv3", "bn3"), ("conv4", "bn4"), ("conv5", "bn5")):
        a = r(np.maximum(block(a, conv, bn), 0))
    g = np.maximum(block(a, "global_feat", "bn_global"), 0).reshape(B, N, -1).max(1)
    cb = g @ sd["seg_conv1.weight"][:, 64:, 0].astype(np.float64).T
    a = r(np.maximum(block(pf, "seg_conv1", "bn_seg1", extra=np.repeat(cb, N, axis=0), wcols=slice(0, 64)), 0))
    a = r(np.maximum(block(a, "seg_conv2", "bn_seg2"), 0))
    z = np.maximum(block(a, "seg_conv3", "bn_seg3"), 0)
    W4 = sd["seg_conv4.weight"][:, :, 0].astype(np.float64)
    return (z @ W4.T + sd["seg_conv4.bias"]).reshape(B, N, -1)
