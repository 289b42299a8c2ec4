"""Training forward / loss / backward / running statistics of the CUDA path vs the oracle (dropout p=0)."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import pointnet_oracle as orc

pytestmark = pytest.mark.gpu

CASES = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "case_*.npz")))
# Stated bf16 tolerances for the END-TO-END training comparison against the fp64 oracle.
# Train-mode BatchNorm + the global arg-max make the end-to-end map ill-conditioned on random weights:
# a CPU emulation that only rounds the same tensors to bf16 shows logits max-error ~0.1-0.17 x max|logit|
# (rms ~0.02) and trunk-gradient cosine ~0.8-0.87; even TF32 rounding (the reference's own CUDA default)
# only reaches trunk cosine ~0.97 (DESIGN.md "Numerics").  The tight, kernel-by-kernel proof is
# tests/test_layerwise_gpu.py; the bounds below catch gross end-to-end errors.
LOGIT_MAX_TOL = 0.30   # max |dlogit| / max |logit|
LOGIT_RMS_TOL = 0.05   # rms |dlogit| / max |logit|
LOSS_TOL = 1e-2        # relative
COS_HEAD = 0.94        # seg_conv4, bn_seg3
COS_MID = 0.80         # seg_conv1..3, bn_seg1..2, bn_global
COS_TRUNK = 0.55       # conv1..5, global_feat, bn1..5 (gradient flows only through arg-max routes)


def _model(C, sd):
    import pcseg_b200
    m = pcseg_b200.PointNetSegmentation(C)
    m.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}, strict=True)
    m = m.cuda().train()
    m.dropout.p = 0.0
    return m


def _min_cos(name):
    mod = name.split(".")[0]
    if mod in ("seg_conv4", "bn_seg3"):
        return COS_HEAD
    if mod in ("seg_conv1", "seg_conv2", "seg_conv3", "bn_seg1", "bn_seg2", "bn_global"):
        return COS_MID
    return COS_TRUNK


def _logit_err(got, ref):
    d = np.abs(got - ref)
    s = np.abs(ref).max()
    return d.max() / s, np.sqrt((d * d).mean()) / s


# parameters that receive gradient ONLY through the pooled feature (pcs.py:114-120)
GLOBAL_ONLY = ("conv3", "conv4", "conv5", "global_feat", "bn3", "bn4", "bn5", "bn_global")


def _check_grads(named_grads, ref_grads, check_cos=True, clouds=None):
    """clouds <= 2: bn_seg1's backward removes the batch mean of dy, so with two clouds the gradients of the two pooled
    features are EXACTLY opposite (dg[0] = -dg[1]) and everything upstream of the max-pool is a difference of near-equal
    terms decided by a handful of ReLU / arg-max ties: only finiteness is checked for those tensors."""
    bad = []
    for name, g in named_grads:
        if clouds is not None and clouds <= 2 and name.split(".")[0] in GLOBAL_ONLY:
            if not np.isfinite(g).all():
                bad.append((name, "not finite"))
            continue
        r = np.asarray(ref_grads[name], np.float64).reshape(g.shape)
        g = g.astype(np.float64)
        if name.endswith(".bias") and name.split(".")[0] in orc.CONV_NAMES[:-1]:
            # mathematically zero (bias before train-mode BN); only require it to be tiny
            other = np.abs(np.asarray(ref_grads[name.replace(".bias", ".weight")])).max()
            if np.abs(g).max() > 5e-2 * other + 1e-6:
                bad.append((name, "nonzero bias grad", np.abs(g).max()))
            continue
        scale = np.abs(r).max()
        if scale < 1e-7:
            if np.abs(g).max() > 2e-3:
                bad.append((name, "expected ~0", np.abs(g).max()))
            continue
        cos = float((g * r).sum() / (np.linalg.norm(g) * np.linalg.norm(r) + 1e-30))
        ratio = np.linalg.norm(g) / np.linalg.norm(r)
        if check_cos and (cos < _min_cos(name) or not (0.5 < ratio < 2.0)):
            bad.append((name, cos, ratio))
    assert not bad, bad


def _run_case(C, sd, x, labels, cw):
    m = _model(C, sd)
    xt = torch.from_numpy(x).cuda()
    lt = torch.from_numpy(labels).cuda()
    crit = torch.nn.CrossEntropyLoss(ignore_index=-1, weight=torch.from_numpy(cw).cuda())
    logits = m(xt)
    loss = crit(logits.contiguous().view(-1, C), lt.view(-1))      # exactly the reference call sequence, pcs.py:244-254
    loss.backward()
    return m, logits.detach().cpu().numpy(), float(loss.item())


@pytest.mark.parametrize("path", CASES, ids=[os.path.basename(p)[:-4] for p in CASES])
def test_train_step_matches_oracle_on_golden_inputs(path):
    gold = np.load(path)
    C, seed = int(gold["C"]), int(gold["seed"])
    sd = orc.synth_state(C, seed)
    x, labels, cw = gold["x"], gold["labels"], gold["class_w"]
    m, logits, loss = _run_case(C, sd, x, labels, cw)

    ref_logits, cache, newbuf = orc.forward_train(sd, x)
    np.testing.assert_allclose(ref_logits, gold["train_logits"], atol=5e-5)     # oracle itself is pinned
    emax, erms = _logit_err(logits, ref_logits)
    assert emax < LOGIT_MAX_TOL and erms < LOGIT_RMS_TOL, (emax, erms)
    ref_loss, dlog = orc.weighted_ce(ref_logits, labels, cw)
    assert abs(loss - ref_loss) < LOSS_TOL * abs(ref_loss)
    assert abs(loss - float(gold["loss"])) < LOSS_TOL * abs(ref_loss)

    grads = orc.backward(cache, dlog)
    # B == 1 makes every gradient through the global branch mathematically zero -> no cosine there
    named = [(n, p.grad.detach().cpu().numpy()) for n, p in m.named_parameters()]
    if x.shape[0] == 1:
        named = [(n, g) for n, g in named if n.split(".")[0] in ("seg_conv4", "bn_seg3", "seg_conv3", "seg_conv2", "bn_seg2")]
    _check_grads(named, grads)

    for name, buf in m.named_buffers():
        if name.endswith("num_batches_tracked"):
            assert int(buf.item()) == int(gold["buf/" + name])
        else:
            np.testing.assert_allclose(buf.cpu().numpy(), newbuf[name], rtol=3e-2, atol=5e-3, err_msg=name)


@pytest.mark.parametrize("B,N,C", [(4, 512, 5), (2, 1000, 3), (8, 2048, 5), (3, 37, 2), (4, 512, 12), (8, 2048, 24)])
def test_train_step_matches_oracle(B, N, C):
    sd = orc.synth_state(C, 7 * B + N)
    rng = np.random.default_rng(B * N)
    x = rng.random((B, N, 4), dtype=np.float32)
    labels = rng.integers(0, C, (B, N)).astype(np.int64)
    if N == 1000:
        x[0, 600:] = 0.0
        labels[0, 600:] = -1
    cw = (0.5 + rng.random(C)).astype(np.float32)
    m, logits, loss = _run_case(C, sd, x, labels, cw)
    ref_logits, cache, newbuf = orc.forward_train(sd, x)
    emax, erms = _logit_err(logits, ref_logits)
    assert emax < LOGIT_MAX_TOL and erms < LOGIT_RMS_TOL, (emax, erms)
    ref_loss, dlog = orc.weighted_ce(ref_logits, labels, cw)
    assert abs(loss - ref_loss) < LOSS_TOL * abs(ref_loss)
    grads = orc.backward(cache, dlog)
    _check_grads([(n, p.grad.detach().cpu().numpy()) for n, p in m.named_parameters()], grads, clouds=B)


def test_fused_trainer_matches_autograd_path():
    """FusedTrainer (fused CE + backward + Adam) vs the autograd path + torch.optim.Adam on the same step."""
    import pcseg_b200
    C, B, N = 5, 4, 768
    sd = orc.synth_state(C, 42)
    rng = np.random.default_rng(5)
    x = torch.from_numpy(rng.random((B, N, 4), dtype=np.float32)).cuda()
    labels = torch.from_numpy(rng.integers(-1, C, (B, N)).astype(np.int64)).cuda()
    cw = torch.tensor([0.5, 1.0, 2.0, 0.75, 0.75], device="cuda")

    m1 = _model(C, sd)
    opt = torch.optim.Adam(m1.parameters(), lr=1e-3, weight_decay=1e-4)
    crit = torch.nn.CrossEntropyLoss(ignore_index=-1, weight=cw)
    opt.zero_grad()
    out = m1(x)
    loss1 = crit(out.contiguous().view(-1, C), labels.view(-1))
    loss1.backward()
    g1 = torch.cat([p.grad.reshape(-1) for p in m1._param_list()])
    opt.step()

    m2 = _model(C, sd)
    tr = pcseg_b200.FusedTrainer(m2, class_weights=cw, lr=1e-3, weight_decay=1e-4)
    res = tr.step(x, labels)
    assert abs(float(res["loss"].item()) - float(loss1.item())) < 1e-4 * abs(float(loss1.item())) + 1e-5
    g2 = tr.flat["grads"]
    # same kernels, but dlogits come from torch's fp32 CE in one path and from the fused CE in the other: ~1e-7 input
    # differences flip individual bf16 roundings downstream, so compare with a bf16-sized tolerance
    cos = torch.nn.functional.cosine_similarity(g1, g2, dim=0).item()
    assert cos > 0.999, cos
    # (norm-wise: a single flipped arg-max route moves individual trunk elements by a few per cent of the largest one)
    assert (g1 - g2).norm().item() < 5e-2 * g1.norm().item()
    p1 = torch.cat([p.detach().reshape(-1) for p in m1._param_list()])
    p2 = tr.flat["params"]
    # Adam normalises the step to ~lr, so compare parameters with an lr-sized tolerance
    assert (p1 - p2).abs().max().item() < 2.5e-3
    valid = int((labels >= 0).sum().item())
    assert int(res["valid"].item()) == valid
    pred = out.argmax(-1)
    assert int(res["correct"].item()) == int(((pred == labels) & (labels >= 0)).sum().item())


def test_dropout_statistics_and_determinism():
    """p=0.3: kept fraction ~0.7 on the seg-head activations is not observable from outside, so check the
    observable contract instead: training output differs from p=0, is finite, and backward runs."""
    C, B, N = 5, 2, 1024
    sd = orc.synth_state(C, 9)
    m = _model(C, sd)
    x = torch.rand(B, N, 4, device="cuda")
    with torch.no_grad():
        base = m(x).clone()
    m.dropout.p = 0.3
    torch.manual_seed(0)
    out = m(x)
    assert torch.isfinite(out).all()
    assert (out.detach() - base).abs().max().item() > 1e-4
    out.sum().backward()
    assert all(torch.isfinite(p.grad).all() for p in m.parameters())


def test_backward_twice_is_rejected():
    C = 3
    m = _model(C, orc.synth_state(C, 2))
    x = torch.rand(2, 128, 4, device="cuda")
    a = m(x)
    b = m(x)
    b.sum().backward()
    with pytest.raises(RuntimeError):
        a.sum().backward()
