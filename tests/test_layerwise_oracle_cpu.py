"""Pins oracle/layerwise.py — the per-kernel fp64 restatements every CUDA kernel is checked against in
tests/test_layerwise_gpu.py — to the end-to-end oracle (which is pinned to the reference's golden vectors): chaining the
layer-wise functions in the order the CUDA path runs them (conv without bias, BN from batch sums, max-pool through the
sign of the BN scale, seg_conv1 split into point-feature GEMM + per-cloud term, BN backward as dy = A dz + B y + C)
reproduces the oracle's logits, loss and gradients.  CPU only."""
import numpy as np
import pytest

from oracle import layerwise as lw
from oracle import pointnet_oracle as orc

CONVS = [c for c, _, _, _ in orc.TRUNK + orc.HEAD]
BNS = [b for _, b, _, _ in orc.TRUNK + orc.HEAD]


@pytest.mark.parametrize("C,B,N,seed", [(5, 3, 200, 2), (3, 2, 129, 8)])
def test_layerwise_chain_equals_the_end_to_end_oracle(C, B, N, seed):
    sd = orc.synth_state(C, seed)
    for c in orc.CONV_NAMES:                                   # bf16-representable weights: the restatements round them
        sd[f"{c}.weight"] = lw.bf16_round(sd[f"{c}.weight"]).astype(np.float32)
    rng = np.random.default_rng(seed)
    x = rng.random((B, N, 4))
    labels = rng.integers(-1, C, (B, N))
    cw = 0.5 + rng.random(C)
    P = B * N
    W = {c: sd[f"{c}.weight"][:, :, 0].astype(np.float64) for c in orc.CONV_NAMES}

    # ---- forward, layer by layer
    y, a, bnp, relu = {}, {}, {}, {}
    a_prev = x.reshape(P, 4)
    for i, (conv, bn) in enumerate(zip(CONVS, BNS)):
        if i == 6:                                             # seg_conv1: point-feature columns + per-cloud term
            g, ystar, arg = lw.maxpool(y[5], bnp[5], B, N)
            cb = g @ W[conv][:, 64:].T
            y[i] = lw.conv_pre_bn(a[1], W[conv][:, :64], cloud_bias=cb, pts_per_cloud=N)
        else:
            y[i] = lw.conv_pre_bn(a_prev, W[conv])
        bnp[i] = lw.bn_params(lw.bn_batch_stats(y[i]), P, sd[f"{bn}.weight"], sd[f"{bn}.bias"])
        a[i], relu[i] = lw.bn_relu(y[i], bnp[i])
        a_prev = a[i]
    logits, a8 = lw.head_logits(y[8], bnp[8], W["seg_conv4"], sd["seg_conv4.bias"])

    ref_logits, cache, _ = orc.forward_train(sd, x)
    np.testing.assert_allclose(logits.reshape(B, N, C), ref_logits, rtol=0, atol=2e-5 * np.abs(ref_logits).max())
    live = g > 0          # (a channel that ReLU zeroes for every point of a cloud has no unique arg-max and no gradient)
    assert np.array_equal(arg[live], cache["argmax"][live])

    # ---- loss gradient and backward, layer by layer
    wsum = cw[np.where(labels >= 0, labels, 0)].reshape(-1)[labels.reshape(-1) >= 0].sum()
    dl, num, wsum2 = lw.ce_grad(logits, labels, cw, wsum)
    ref_loss, ref_dl = orc.weighted_ce(ref_logits, labels, cw)
    assert abs(num / wsum2 - ref_loss) < 1e-6 and abs(wsum - wsum2) < 1e-9
    ref = orc.backward(cache, ref_dl)

    grads = {"seg_conv4.weight": lw.wgrad(dl, a8), "seg_conv4.bias": dl.sum(0)}
    dz = {8: (dl @ W["seg_conv4"]) * relu[8]}
    dy = {}

    def bn_backward(i):
        sb = lw.bn_bwd_stats(dz[i], y[i], bnp[i])
        grads[f"{BNS[i]}.bias"], grads[f"{BNS[i]}.weight"] = sb[0], sb[1]
        dy[i] = lw.bn_bwd_apply(dz[i], y[i], lw.bn_bwd_coef(sb, P, bnp[i]))

    for i in (8, 7):
        bn_backward(i)
        grads[f"{CONVS[i]}.weight"] = lw.wgrad(dy[i], a[i - 1])
        dz[i - 1] = lw.dgrad_masked(dy[i], W[CONVS[i]], y[i - 1], bnp[i - 1])
    bn_backward(6)
    dcb = dy[6].reshape(B, N, -1).sum(1)                       # per-cloud sums: gradient of the broadcast term
    grads["seg_conv1.weight"] = np.concatenate([lw.wgrad(dy[6], a[1]), dcb.T @ g], axis=1)
    dg = (dcb @ W["seg_conv1"][:, 64:]) * (g > 0)              # repeat + cat + relu backward, per cloud
    dz5 = np.zeros((B, N, 1024))
    np.put_along_axis(dz5, arg[:, None, :].astype(np.int64), dg[:, None, :], axis=1)   # max backward: arg-max rows only
    dz[5] = dz5.reshape(P, -1)
    d_pf_from_head = dy[6] @ W["seg_conv1"][:, :64]            # gradient reaching point_feat through seg_conv1
    for i in (5, 4, 3, 2):
        bn_backward(i)
        grads[f"{CONVS[i]}.weight"] = lw.wgrad(dy[i], a[i - 1])
        if i == 2:                                             # skip join at point_feat (conv2's output)
            t = np.float32(bnp[1][:, 0]) * y[1].astype(np.float32) + np.float32(bnp[1][:, 1])
            dz[1] = (dy[2] @ W["conv3"] + d_pf_from_head) * (t > 0)
        else:
            dz[i - 1] = lw.dgrad_masked(dy[i], W[CONVS[i]], y[i - 1], bnp[i - 1])
    bn_backward(1)
    grads["conv2.weight"] = lw.wgrad(dy[1], a[0])
    dz[0] = lw.dgrad_masked(dy[1], W["conv2"], y[0], bnp[0])
    bn_backward(0)
    grads["conv1.weight"] = lw.wgrad(dy[0], x.reshape(P, 4))

    for name, got in grads.items():
        r = np.asarray(ref[name], np.float64).reshape(np.shape(got))
        np.testing.assert_allclose(got, r, rtol=0, atol=2e-4 * np.abs(r).max() + 1e-9, err_msg=name)
