"""Kernel-by-kernel parity of the training path: every CUDA step is recomputed in fp64 by
oracle/layerwise.py FROM THE TENSORS THE CUDA PATH PRODUCED for the previous step, so each kernel is
compared on identical inputs (see the module docstring of oracle/layerwise.py for why end-to-end
comparison alone cannot be tight in train mode).  Integer outputs (arg-max indices, counts) must be
bit-exact; bf16 tensors must match to one bf16 rounding; fp32/fp64 reductions to ~1e-4 relative."""
import numpy as np
import pytest
import torch

from oracle import layerwise as lw
from oracle import pointnet_oracle as orc

pytestmark = pytest.mark.gpu

CONVS = ["conv1", "conv2", "conv3", "conv4", "conv5", "global_feat", "seg_conv1", "seg_conv2", "seg_conv3", "seg_conv4"]
BNS = ["bn1", "bn2", "bn3", "bn4", "bn5", "bn_global", "bn_seg1", "bn_seg2", "bn_seg3"]


def _np(t):
    return t.detach().to(torch.float64).cpu().numpy() if t.dtype != torch.int32 else t.cpu().numpy()


def _close_bf16(got, ref, what):
    """got is a bf16 tensor produced by a kernel that rounded ref-like fp32 values: allow one rounding step."""
    got, ref = np.asarray(got, np.float64), np.asarray(ref, np.float64)
    rms = np.sqrt((ref * ref).mean()) + 1e-30
    tol = 2.0 ** -7 * np.abs(ref) + 2e-3 * rms
    bad = np.abs(got - ref) > tol
    assert bad.mean() < 1e-4, (what, float(bad.mean()), float(np.abs(got - ref).max()), float(rms))


def _close_rms(got, ref, what, rel):
    """two differently-rounded evaluations of the same quantity: relative RMS difference"""
    got, ref = np.asarray(got, np.float64), np.asarray(ref, np.float64)
    rms = np.sqrt((ref * ref).mean()) + 1e-30
    err = np.sqrt(((got - ref) ** 2).mean())
    assert err <= rel * rms, (what, float(err), float(rms))


def _close_red(got, ref, what, rel=2e-4):
    got, ref = np.asarray(got, np.float64), np.asarray(ref, np.float64)
    scale = np.abs(ref).max() + 1e-30
    assert np.abs(got - ref).max() <= rel * scale + 1e-9, (what, float(np.abs(got - ref).max()), float(scale))


def _diverse_clouds(B, N, rng):
    xs = []
    for _ in range(B):
        k = rng.integers(2, 6)
        cen, sc = rng.random((k, 4)), 0.02 + 0.25 * rng.random((k, 4))
        idx = rng.integers(0, k, N)
        p = cen[idx] + sc[idx] * rng.standard_normal((N, 4))
        p[:, 3] = np.abs(p[:, 3]) * (0.2 + 2 * rng.random())
        xs.append(np.clip(p, -0.5, 2.0))
    return np.stack(xs).astype(np.float32)


@pytest.mark.parametrize("B,N,C,p_drop,folded", [(3, 500, 5, 0.0, True), (4, 1024, 3, 0.3, True), (2, 200, 8, 0.0, True),
                                                 (2, 384, 12, 0.3, True), (2, 200, 32, 0.0, True),
                                                 # 256 row tiles: every persistent CTA walks several tiles (TMEM double buffering, per-CTA
                                                 # column accumulators, split-K over many k-blocks)
                                                 (2, 16384, 5, 0.3, True),
                                                 (3, 500, 5, 0.0, False), (4, 1024, 3, 0.3, False)])
def test_every_training_kernel_against_its_own_inputs(B, N, C, p_drop, folded, monkeypatch):
    """folded=True: conv5 runs with Gram-predicted BatchNorm statistics and the folded BatchNorm backward (its y / dy are
    never materialised, oracle/folded_bn_ref.py); folded=False: the legacy step that materialises them (also the path
    of ragged batches)."""
    import pcseg_b200
    from pcseg_b200.engine import debug_tensor
    from oracle import folded_bn_ref as fb

    monkeypatch.setenv("PCSEG_FOLDED", "1" if folded else "0")      # read when a context is created
    monkeypatch.setenv("PCSEG_STORE_Y5", "1")      # keep global_feat's pre-BN output so that the fused max-pool can be checked bit-exactly
    FOLDED = {4, 5} if folded else set()
    if folded and N % 128 == 0:
        FOLDED.add(6)            # seg_conv1: per-cloud Gram matrices; needs row tiles that never straddle clouds

    P = B * N
    sd = orc.synth_state(C, 1000 + N)
    rng = np.random.default_rng(N + C)
    x = _diverse_clouds(B, N, rng)
    labels = rng.integers(0, C, (B, N)).astype(np.int64)
    x[0, N - N // 5:] = 0.0                                # zero-padded tail with ignored labels (pcs.py:53-61)
    labels[0, N - N // 5:] = -1
    cw = (0.5 + rng.random(C)).astype(np.float32)

    m = pcseg_b200.PointNetSegmentation(C)
    m.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()})
    m = m.cuda().train()
    m.dropout.p = p_drop
    dev = torch.device("cuda", torch.cuda.current_device())
    f = m._ensure_flat(dev)
    eng = m._get_engine(dev)
    xt, lt, cwt = torch.from_numpy(x).cuda(), torch.from_numpy(labels).cuda(), torch.from_numpy(cw).cuda()
    ce = torch.zeros(32, dtype=torch.uint8, device=dev)
    logits = m._run_train_forward(xt, labels=lt, class_w=cwt, ce=ce)
    wsum_t = ce.view(torch.float64)[1:2].clone()
    eng.backward(xt, f["params"], f["grads"], logits=logits, labels=lt, class_w=cwt, wsum=wsum_t)
    torch.cuda.synchronize()
    grads = {n: g.detach().cpu().numpy().astype(np.float64) for (n, _), g in zip(m.named_parameters(), m.grad_views())}

    def T(kind, layer=0):
        return _np(debug_tensor(eng, B, N, kind, layer))

    W = {c: sd[f"{c}.weight"][:, :, 0].astype(np.float64) for c in CONVS}
    keep_scale = 1.0
    if p_drop > 0:
        thr = int(p_drop * 65536 + 0.5)
        keep_scale = 1.0 / (1.0 - thr / 65536.0)

    # ---------------------------------------------------------------- forward
    y, act, bnp = {}, {}, {}
    xin = x.reshape(P, 4).astype(np.float64)
    keeps = {}
    for i in range(9):
        bnp[i] = T("bnp", i)
        if i == 6 and 6 in FOLDED:
            # ---- seg_conv1 with predicted statistics: per-cloud Gram matrices / column sums of point_feat, per-cloud term cb
            cb = T("cb")
            a1c = act[1].reshape(B, N, 64)
            s1c = a1c.sum(1)
            _close_red(T("gram6").reshape(B, 64, 64), np.einsum("bnk,bnj->bkj", a1c, a1c) - np.einsum("bk,bj->bkj", s1c, s1c) / N,
                       "centred per-cloud gram[seg_conv1]", rel=1e-5)
            _close_red(T("colsum6"), a1c.sum(1), "per-cloud colsum[seg_conv1]", rel=1e-5)
            yref = lw.conv_pre_bn(act[1], W["seg_conv1"][:, :64], cloud_bias=cb, pts_per_cloud=N)
            st = lw.bn_batch_stats(yref)
            _close_red(T("stats_f", i), st, "predicted stats_f[bn_seg1]", rel=1e-4)
            exact = lw.bn_params(st, P, sd["bn_seg1.weight"], sd["bn_seg1.bias"])
            assert np.abs(bnp[i][:, 2] / exact[:, 2] - 1).max() < 2e-4, "predicted 1/std of bn_seg1 vs the fp64 batch statistics"
            st = T("stats_f", i)
            rm, rv = lw.running_stats(st, P, sd["seg_conv1.bias"].astype(np.float64), sd["bn_seg1.running_mean"].astype(np.float64),
                                      sd["bn_seg1.running_var"].astype(np.float64))
            _close_red(_np(m.bn_seg1.running_mean), rm, "bn_seg1.running_mean", rel=1e-5)
            _close_red(_np(m.bn_seg1.running_var), rv, "bn_seg1.running_var", rel=1e-5)
            assert int(m.bn_seg1.num_batches_tracked.item()) == int(sd["bn_seg1.num_batches_tracked"]) + 1
            act[i] = T("act", i)
            relu_out, _ = lw.bn_relu(yref, bnp[i])
            if p_drop > 0:
                keep = (act[i] != 0).astype(np.float64)
                live = relu_out > 1e-3
                assert abs(1.0 - keep[live].mean() - p_drop) < 0.01, "dropout fraction of seg_conv1"
                keeps[i] = keep
                _close_bf16(act[i], relu_out * keep * keep_scale, "act[seg_conv1] (BN + ReLU + dropout in the GEMM epilogue)")
            else:
                _close_bf16(act[i], relu_out, "act[seg_conv1] (BN + ReLU in the GEMM epilogue)")
            y[i] = yref
            continue
        if i in FOLDED and i != 5:
            # ---- Gram-predicted statistics (pcs.py:110): G and s of the input activation, {sum y, sum y^2} predicted from
            # them, BN + ReLU applied straight to the fp32 accumulators; y itself is never stored
            Ci = act[i - 1].shape[1]
            s_in = act[i - 1].sum(0)
            _close_red(T("gram", i), act[i - 1].T @ act[i - 1] - np.outer(s_in, s_in) / P, f"centred gram[{CONVS[i]}]", rel=1e-5)
            _close_red(T("colsum", i)[0], s_in, f"colsum[{CONVS[i]}]", rel=1e-5)
            yref = lw.conv_pre_bn(act[i - 1], W[CONVS[i]])                                  # fp64, bf16 weights
            st = lw.bn_batch_stats(yref)
            _close_red(T("stats_f", i), st, f"predicted stats_f[{BNS[i]}]", rel=1e-4)
            exact = lw.bn_params(st, P, sd[f"{BNS[i]}.weight"], sd[f"{BNS[i]}.bias"])
            assert np.abs(bnp[i][:, 2] / exact[:, 2] - 1).max() < 2e-4, "predicted 1/std vs the fp64 batch statistics, per channel"
            st = T("stats_f", i)
            _close_red(bnp[i], lw.bn_params(st, P, sd[f"{BNS[i]}.weight"], sd[f"{BNS[i]}.bias"]), f"bnp[{BNS[i]}]", rel=1e-4)
            rm, rv = lw.running_stats(st, P, sd[f"{CONVS[i]}.bias"].astype(np.float64), sd[f"{BNS[i]}.running_mean"].astype(np.float64),
                                      sd[f"{BNS[i]}.running_var"].astype(np.float64))
            bn_mod = getattr(m, BNS[i])
            _close_red(_np(bn_mod.running_mean), rm, f"{BNS[i]}.running_mean", rel=1e-5)
            _close_red(_np(bn_mod.running_var), rv, f"{BNS[i]}.running_var", rel=1e-5)
            assert int(bn_mod.num_batches_tracked.item()) == int(sd[f"{BNS[i]}.num_batches_tracked"]) + 1
            act[i] = T("act", i)
            relu_out, _ = lw.bn_relu(yref, bnp[i])
            _close_bf16(act[i], relu_out, f"act[{CONVS[i]}] (BN + ReLU in the GEMM epilogue)")
            y[i] = yref
            continue
        y[i] = T("y", i)
        if i == 0:
            ref = lw.conv_pre_bn(xin, W["conv1"], weights_bf16=False)                      # pcs.py:106
        elif i == 6:
            cb = T("cb")
            ref = lw.conv_pre_bn(act[1], W["seg_conv1"][:, :64], cloud_bias=cb, pts_per_cloud=N)   # pcs.py:117-123
        else:
            ref = lw.conv_pre_bn(act[i - 1], W[CONVS[i]])                                   # pcs.py:107-127
        _close_bf16(y[i], ref, f"y[{CONVS[i]}]")
        st = lw.bn_batch_stats(y[i])
        _close_red(T("stats_f", i), st, f"stats_f[{BNS[i]}]")
        _close_red(bnp[i], lw.bn_params(st, P, sd[f"{BNS[i]}.weight"], sd[f"{BNS[i]}.bias"]), f"bnp[{BNS[i]}]", rel=1e-4)
        rm, rv = lw.running_stats(st, P, sd[f"{CONVS[i]}.bias"].astype(np.float64), sd[f"{BNS[i]}.running_mean"].astype(np.float64),
                                  sd[f"{BNS[i]}.running_var"].astype(np.float64))
        bn_mod = getattr(m, BNS[i])
        _close_red(_np(bn_mod.running_mean), rm, f"{BNS[i]}.running_mean", rel=1e-5)
        _close_red(_np(bn_mod.running_var), rv, f"{BNS[i]}.running_var", rel=1e-5)
        assert int(bn_mod.num_batches_tracked.item()) == int(sd[f"{BNS[i]}.num_batches_tracked"]) + 1
        if i in (5, 8):
            continue                     # global_feat feeds the max-pool, seg_conv3 feeds the head kernel
        act[i] = T("act", i)
        relu_out, _ = lw.bn_relu(y[i], bnp[i])
        if i in (6, 7) and p_drop > 0:
            keep = (act[i] != 0).astype(np.float64)
            live = relu_out > 1e-3
            frac = 1.0 - keep[live].mean()
            assert abs(frac - p_drop) < 0.01, (BNS[i], frac)
            keeps[i] = keep
            _close_bf16(act[i], relu_out * keep * keep_scale, f"act[{CONVS[i]}] (dropout)")
        else:
            _close_bf16(act[i], relu_out, f"act[{CONVS[i]}]")

    g_ref, ystar_ref, arg_ref = lw.maxpool(y[5], bnp[5], B, N)                              # pcs.py:114
    assert np.array_equal(T("argidx"), arg_ref), "max-pool arg indices must be bit-exact"
    assert np.array_equal(T("ystar"), ystar_ref), "max-pool extremum values must be bit-exact"
    _close_red(T("g"), g_ref, "pooled feature", rel=1e-5)
    g = T("g")
    _close_red(T("cb"), g @ W["seg_conv1"][:, 64:].T, "per-cloud seg_conv1 term", rel=1e-4)

    lg = logits.detach().cpu().numpy().astype(np.float64).reshape(P, C)
    lg_ref, a_s3 = lw.head_logits(y[8], bnp[8], W["seg_conv4"], sd["seg_conv4.bias"])       # pcs.py:127-131
    _close_red(lg, lg_ref, "logits", rel=1e-4)

    dl, loss_num, wsum = lw.ce_grad(lg, labels, cw, float(wsum_t.item()))                    # pcs.py:216,251
    cef = ce.view(torch.float64).cpu().numpy()
    cei = ce.view(torch.int64).cpu().numpy()
    assert abs(cef[1] - wsum) <= 1e-9 * wsum
    assert abs(cef[0] - loss_num) <= 1e-5 * abs(loss_num)
    valid = labels.reshape(-1) >= 0
    assert int(cei[3]) == int(valid.sum())                                                   # integer outputs: exact
    assert int(cei[2]) == int(((lg.argmax(1) == labels.reshape(-1)) & valid).sum())          # pcs.py:261-264

    # ---------------------------------------------------------------- backward
    dz, dy = {}, {}
    _close_red(grads["seg_conv4.weight"][:, :, 0], dl.T @ a_s3, "dW seg_conv4", rel=3e-4)
    _close_red(grads["seg_conv4.bias"], dl.sum(0), "db seg_conv4", rel=3e-4)
    dz[8] = T("dz", 8)
    t8 = np.float32(bnp[8][:, 0]) * y[8].astype(np.float32) + np.float32(bnp[8][:, 1])
    _close_bf16(dz[8], (dl @ W["seg_conv4"]) * (t8 > 0), "dz[seg_conv3]")

    def bn_backward(i, sparse=None):
        sb = T("stats_b", i)
        dz_i = sparse if sparse is not None else dz[i]
        _close_red(sb, lw.bn_bwd_stats(dz_i, y[i], bnp[i]), f"stats_b[{BNS[i]}]", rel=3e-4)
        _close_red(grads[f"{BNS[i]}.weight"], sb[1], f"dgamma {BNS[i]}", rel=1e-5)
        _close_red(grads[f"{BNS[i]}.bias"], sb[0], f"dbeta {BNS[i]}", rel=1e-5)
        coef = T("coef", i)[:, :3]
        _close_red(coef, lw.bn_bwd_coef(sb, P, bnp[i]), f"coef[{BNS[i]}]", rel=1e-4)
        dy[i] = T("dy", i)
        _close_bf16(dy[i], lw.bn_bwd_apply(dz_i, y[i], coef), f"dy[{CONVS[i]}]")
        # conv bias gradient = column sum of dy: mathematically ~0 behind train-mode BN; compare as a reduction
        ref_db = dy[i].sum(0)
        assert np.abs(grads[f"{CONVS[i]}.bias"] - ref_db).max() <= 1e-3 * np.abs(dy[i]).sum(0).max() + 1e-7

    # seg_conv3 / seg_conv2
    bn_backward(8)
    _close_red(grads["seg_conv3.weight"][:, :, 0], lw.wgrad(dy[8], act[7]), "dW seg_conv3", rel=3e-4)
    dz[7] = T("dz", 7)
    _close_bf16(dz[7], lw.dgrad_masked(dy[8], W["seg_conv3"], y[7], bnp[7], keeps.get(7), keep_scale), "dz[seg_conv2]")
    bn_backward(7)
    _close_red(grads["seg_conv2.weight"][:, :, 0], lw.wgrad(dy[7], act[6]), "dW seg_conv2", rel=3e-4)
    dW1 = grads["seg_conv1.weight"][:, :, 0]
    if 6 in FOLDED:
        # ---- folded seg_conv1: dz6 is written straight into the skip-join operand, masked by the stored activation
        dz[6] = T("dy", 6)
        da6 = dy[7] @ lw.bf16_round(W["seg_conv2"])
        _close_bf16(dz[6], lw.bf16_round(da6 * keep_scale) * (act[6] > 0), "dz[seg_conv1] (mask from the stored activation)")
        _close_red(T("stats_b", 6)[0], dz[6].sum(0), "sum dz[seg_conv1]", rel=3e-4)
        S1 = T("cloudsum6")
        _close_red(S1, dz[6].reshape(B, N, -1).sum(1), "per-cloud sum dz[seg_conv1]", rel=3e-4)
        Q6 = T("qraw", 6)
        _close_red(Q6, dz[6].T @ act[1], "Q[seg_conv1]", rel=3e-4)
        yhat6 = y[6] * bnp[6][:, 2] + bnp[6][:, 3]
        sb6 = np.stack([dz[6].sum(0), (dz[6] * yhat6).sum(0)])
        _close_red(grads["bn_seg1.weight"], sb6[1], "dgamma bn_seg1", rel=1e-3)
        _close_red(grads["bn_seg1.bias"], sb6[0], "dbeta bn_seg1", rel=3e-4)
        assert np.abs(grads["seg_conv1.bias"]).max() == 0.0
        coef6_ref = lw.bn_bwd_coef(sb6, P, bnp[6])
        _close_red(T("coef", 6)[:, :2], coef6_ref[:, :2], "coef[bn_seg1] A, Bc", rel=1e-3)
        dy[6] = lw.bn_bwd_apply(dz[6], y[6], coef6_ref)                      # never materialised by the CUDA path
        _close_red(dW1[:, :64], lw.wgrad(dy[6], act[1]), "dW seg_conv1[:, :64] (definition)", rel=2e-3)
        dcb_ref = dy[6].reshape(B, N, -1).sum(1)
        _close_red(T("dcb"), dcb_ref, "dcb (definition)", rel=2e-3)
        coef6 = T("coef", 6)
        Wpf = lw.bf16_round(W["seg_conv1"][:, :64])
        wcat6 = T("wcat6")
        _close_bf16(wcat6[:, :64], lw.bf16_round(W["conv3"]).T, "wcat6 W3^T")
        _close_bf16(wcat6[:, 64:576], (coef6[:, 0:1] * Wpf).T, "wcat6 scaled Wpf^T")
        _close_bf16(wcat6[:, 576:], (Wpf.T @ (coef6[:, 1:2] * Wpf)).T, "wcat6 S6")
        mean6 = -bnp[6][:, 3] / bnp[6][:, 2]
        _close_red(T("cst6"), (coef6[:, 1] * (T("cb") - mean6) + coef6[:, 3]) @ Wpf, "per-cloud constant rows", rel=1e-3)
    else:
        dz[6] = T("dz", 6)
        _close_bf16(dz[6], lw.dgrad_masked(dy[7], W["seg_conv2"], y[6], bnp[6], keeps.get(6), keep_scale), "dz[seg_conv1]")
        # seg_conv1: point-feature columns + per-cloud global columns (cat / repeat / max backward, pcs.py:114-120)
        bn_backward(6)
        _close_red(dW1[:, :64], lw.wgrad(dy[6], act[1]), "dW seg_conv1[:, :64]", rel=3e-4)
        dcb_ref = dy[6].reshape(B, N, -1).sum(1)
        _close_red(T("dcb"), dcb_ref, "dcb", rel=1e-3)
    dcb = T("dcb")
    _close_red(dW1[:, 64:], dcb.T @ g, "dW seg_conv1[:, 64:]", rel=1e-4)
    dzv_ref = (dcb @ W["seg_conv1"][:, 64:]) * (g > 0)
    _close_red(T("dzv"), dzv_ref, "max-pool routed gradient", rel=1e-4)
    dzv = T("dzv")
    arg = T("argidx")
    dz6 = np.zeros((B, N, 1024))
    np.put_along_axis(dz6, arg[:, None, :].astype(np.int64), dzv[:, None, :], axis=1)
    dz[4] = T("dz", 4)
    if 5 in FOLDED:
        # ---- folded BatchNorm backward of global_feat: no y5 / dy5 (oracle/folded_bn_ref.py), max-pool gradient rows through
        # the side buffer, S5 / W5 Gc4 on the tensor cores
        dz5 = dz6.reshape(P, 1024)
        sb5 = T("stats_b", 5)
        _close_red(sb5, lw.bn_bwd_stats(dz5, y[5], bnp[5]), "stats_b[bn_global]", rel=3e-4)
        A5 = bnp[5][:, 0]
        Bc5 = -A5 * bnp[5][:, 2] * sb5[1] / P
        D5 = -A5 * sb5[0] / P
        mean5 = -bnp[5][:, 3] / bnp[5][:, 2]
        coef5 = T("coef", 5)
        _close_red(coef5, np.stack([A5, Bc5, D5 - Bc5 * mean5, D5], axis=1), "coef[bn_global]", rel=1e-4)
        _close_red(grads["bn_global.weight"], sb5[1], "dgamma bn_global", rel=1e-5)
        _close_red(grads["bn_global.bias"], sb5[0], "dbeta bn_global", rel=1e-5)
        assert np.abs(grads["global_feat.bias"]).max() == 0.0
        Wb5 = lw.bf16_round(W["global_feat"])
        S5b = T("s5b")
        _close_bf16(S5b, lw.bf16_round(coef5[:, 1:2] * Wb5).T @ Wb5, "S5 = W5^T diag(Bc) W5")
        cst5 = T("cstf", 5)[0]
        _close_red(cst5, coef5[:, 2] @ Wb5, "const[global_feat]", rel=1e-4)
        # side buffer: one slot per arg-max row (the lowest routing (cloud, channel) index), bit-exact
        rowflat = (arg.astype(np.int64) + (np.arange(B) * N)[:, None]).reshape(-1)
        active = dzv.reshape(-1) != 0
        slot_ref = np.full(P, 0x7F7F7F7F, np.int64)
        np.minimum.at(slot_ref, rowflat[active], np.arange(B * 1024)[active])
        assert np.array_equal(debug_tensor(eng, B, N, "rowslot").cpu().numpy()[0].astype(np.int64), slot_ref), "side-buffer slots"
        ch = np.tile(np.arange(1024), B)
        E_ref = np.zeros((B * 1024, 1024))
        np.add.at(E_ref, slot_ref[rowflat[active]], (coef5[ch[active], 0] * dzv.reshape(-1)[active])[:, None] * Wb5[ch[active]])
        side = T("side")
        owned = np.unique(slot_ref[rowflat[active]])              # slots nobody owns are never initialised nor read
        _close_red(side[owned], E_ref[owned], "max-pool gradient rows", rel=1e-4)
        side = np.where(np.isin(np.arange(B * 1024), owned)[:, None], side, 0.0)
        Q5 = T("qraw", 5)
        _close_red(Q5, dz5.T @ act[4], "Q[global_feat]", rel=3e-4)
        da4 = act[4] @ S5b.T + cst5
        has = slot_ref < B * 1024
        da4[has] += side[slot_ref[has]]
        _close_bf16(dz[4], da4 * (act[4] > 0), "dz[conv5] (folded data gradient)")
        coef5_ref = lw.bn_bwd_coef(sb5, P, bnp[5])
        dy5_ref = lw.bn_bwd_apply(dz5, y[5], coef5_ref)
        _close_rms(dz[4], (dy5_ref @ Wb5) * (act[4] > 0), "dz[conv5] (definition)", rel=2e-2)
        s4 = T("stats_b", 4)[1]
        _close_red(s4, act[4].sum(0), "sum act[conv5]", rel=3e-4)
        G4 = np.triu(T("gram", 5))
        _close_red(G4, np.triu(act[4].T @ act[4]), "gram[global_feat] (upper triangle)", rel=1e-4)      # fp32 split-K atomics
        G4 = G4 + np.triu(G4, 1).T
        Gc5b = T("gc5b")
        _close_bf16(Gc5b, G4 - np.outer(s4, s4) / P, "centred Gram matrix")
        dW5 = grads["global_feat.weight"][:, :, 0]
        _close_red(dW5, coef5[:, 0:1] * Q5 + coef5[:, 1:2] * (Wb5 @ Gc5b) + np.outer(coef5[:, 3], s4), "dW global_feat", rel=3e-4)
        _close_red(dW5, lw.wgrad(dy5_ref, act[4]), "dW global_feat (definition)", rel=3e-3)
    else:
        bn_backward(5, sparse=dz6.reshape(P, 1024))
        _close_red(grads["global_feat.weight"][:, :, 0], lw.wgrad(dy[5], act[4]), "dW global_feat", rel=3e-4)
        if 4 in FOLDED:
            _close_bf16(dz[4], (dy[5] @ lw.bf16_round(W["global_feat"])) * (act[4] > 0), "dz[conv5] (mask from the stored activation)")
        else:
            _close_bf16(dz[4], lw.dgrad_masked(dy[5], W["global_feat"], y[4], bnp[4]), "dz[conv5]")

    def folded_backward(i, prev):
        """conv i without y / dy (oracle/folded_bn_ref.py identities (2), (3)), every quantity from the CUDA tensors"""
        sb = T("stats_b", i)
        _close_red(sb[0], dz[i].sum(0), f"sum dz[{CONVS[i]}]", rel=3e-4)
        _close_red(sb[1], act[i].sum(0), f"sum act[{CONVS[i]}]", rel=3e-4)
        Q = T("qraw", i)
        _close_red(Q, dz[i].T @ act[prev], f"Q[{CONVS[i]}]", rel=3e-4)
        Wb = lw.bf16_round(W[CONVS[i]])
        s = T("colsum", i)[0]
        G = T("gram", i) + np.outer(s, s) / P            # the kernels keep the CENTRED Gram matrix
        gamma = sd[f"{BNS[i]}.weight"].astype(np.float64)
        mean, invstd = -bnp[i][:, 3] / bnp[i][:, 2], bnp[i][:, 2]
        fr = fb.folded_layer_backward(Wb, np.zeros(Wb.shape[0]), s, G, P, gamma, mean, invstd, Q, sb[0])
        _close_red(grads[f"{BNS[i]}.weight"], fr["dgamma"], f"dgamma {BNS[i]}", rel=1e-4)
        _close_red(grads[f"{BNS[i]}.bias"], fr["dbeta"], f"dbeta {BNS[i]}", rel=1e-5)
        # dgamma must also be what the un-folded definition gives: sum dz * yhat with yhat from the fp64 y
        _close_red(grads[f"{BNS[i]}.weight"], (dz[i] * (y[i] * bnp[i][:, 2] + bnp[i][:, 3])).sum(0), f"dgamma {BNS[i]} (definition)", rel=2e-3)
        _close_red(grads[f"{CONVS[i]}.weight"][:, :, 0], fr["dW"], f"dW {CONVS[i]}", rel=3e-4)
        assert np.abs(grads[f"{CONVS[i]}.bias"]).max() == 0.0
        # and the un-folded definition: dW = dy^T a_prev with dy = A dz + Bc y + Cc
        coef_ref = lw.bn_bwd_coef(np.stack([sb[0], grads[f"{BNS[i]}.weight"]]), P, bnp[i])
        dy_ref = lw.bn_bwd_apply(dz[i], y[i], coef_ref)
        _close_red(grads[f"{CONVS[i]}.weight"][:, :, 0], lw.wgrad(dy_ref, act[prev]), f"dW {CONVS[i]} (definition)", rel=2e-3)
        Bw = T("bwf", i)
        Co = Wb.shape[0]
        _close_bf16(Bw[:, :Co], fr["W_dz"].T, f"Bw[{CONVS[i]}] scaled weights")
        _close_bf16(Bw[:, Co:], fr["S"].T, f"Bw[{CONVS[i]}] S")
        cst = T("cstf", i)[0]
        _close_red(cst, fr["const"], f"const[{CONVS[i]}]", rel=1e-4)
        dz[prev] = T("dz", prev)
        da = np.concatenate([dz[i], act[prev]], axis=1) @ Bw.T + cst
        t = np.float32(bnp[prev][:, 0]) * y[prev].astype(np.float32) + np.float32(bnp[prev][:, 1])
        _close_bf16(dz[prev], da * (t > 0), f"dz[{CONVS[prev]}] (folded data gradient)")
        # the un-folded definition rounds dy to bf16 where the folded GEMM rounds S: same value, different roundings
        _close_rms(dz[prev], (dy_ref @ Wb) * (t > 0), f"dz[{CONVS[prev]}] (definition)", rel=2e-2)

    for i, prev in ((4, 3), (3, 2)):
        if i in FOLDED:
            folded_backward(i, prev)
            continue
        bn_backward(i)
        _close_red(grads[f"{CONVS[i]}.weight"][:, :, 0], lw.wgrad(dy[i], act[prev]), f"dW {CONVS[i]}", rel=3e-4)
        dz[prev] = T("dz", prev)
        _close_bf16(dz[prev], lw.dgrad_masked(dy[i], W[CONVS[i]], y[prev], bnp[prev]), f"dz[{CONVS[prev]}]")
    # conv3, then the skip join into point_feat (pcs.py:107,120): both contributions in one accumulator
    bn_backward(2)
    _close_red(grads["conv3.weight"][:, :, 0], lw.wgrad(dy[2], act[1]), "dW conv3", rel=3e-4)
    dz[1] = T("dz", 1)
    da_join = dy[2] @ lw.bf16_round(W["conv3"]) + dy[6] @ lw.bf16_round(W["seg_conv1"][:, :64])
    t1 = np.float32(bnp[1][:, 0]) * y[1].astype(np.float32) + np.float32(bnp[1][:, 1])
    if 6 in FOLDED:
        da_f = np.concatenate([dy[2], dz[6], act[1]], axis=1) @ T("wcat6").T + np.repeat(T("cst6"), N, axis=0)
        _close_bf16(dz[1], da_f * (t1 > 0), "dz[conv2] (skip join, folded seg_conv1)")
        _close_rms(dz[1], da_join * (t1 > 0), "dz[conv2] (skip join, definition)", rel=2e-2)
    else:
        _close_bf16(dz[1], da_join * (t1 > 0), "dz[conv2] (skip join)")
    bn_backward(1)
    _close_red(grads["conv2.weight"][:, :, 0], lw.wgrad(dy[1], act[0]), "dW conv2", rel=3e-4)
    dz[0] = T("dz", 0)
    _close_bf16(dz[0], lw.dgrad_masked(dy[1], W["conv2"], y[0], bnp[0]), "dz[conv1]")
    bn_backward(0)
    _close_red(grads["conv1.weight"][:, :, 0], lw.wgrad(dy[0], xin), "dW conv1", rel=3e-4)
