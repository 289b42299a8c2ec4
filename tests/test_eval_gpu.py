"""Eval-mode forward of the CUDA path vs the numpy oracle and the reference-generated golden vectors."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import pointnet_oracle as orc

pytestmark = pytest.mark.gpu

# bf16 operands, fp32 accumulation: stated tolerance (north_star allows a looser bf16 bound)
LOGIT_TOL_REL_TO_MAX = 2e-2
ARGMAX_AGREE = 0.999

CASES = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "case_*.npz")))


def _model(C, sd):
    import pcseg_b200
    m = pcseg_b200.PointNetSegmentation(C)
    m.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}, strict=True)
    return m.cuda().eval()


def _argmax_agreement(got, ref):
    """argmax agreement, not counting points whose top-2 reference logits are closer than the bf16 error bound."""
    if ref.shape[-1] == 1:
        return 1.0, 1.0
    top2 = np.sort(ref, axis=-1)[..., -2:]
    margin = top2[..., 1] - top2[..., 0]
    decided = margin > 2 * LOGIT_TOL_REL_TO_MAX * np.abs(ref).max()
    agree_all = (got.argmax(-1) == ref.argmax(-1)).mean()
    agree_decided = (got.argmax(-1) == ref.argmax(-1))[decided].mean() if decided.any() else 1.0
    return agree_all, agree_decided


@pytest.mark.parametrize("path", CASES, ids=[os.path.basename(p)[:-4] for p in CASES])
def test_eval_matches_golden(path):
    gold = np.load(path)
    C, seed = int(gold["C"]), int(gold["seed"])
    sd = orc.synth_state(C, seed)
    m = _model(C, sd)
    with torch.no_grad():
        got = m(torch.from_numpy(gold["x"]).cuda()).cpu().numpy()
    ref = gold["eval_logits"]
    assert got.shape == ref.shape
    err = np.abs(got - ref).max() / np.abs(ref).max()
    assert err < LOGIT_TOL_REL_TO_MAX, err
    _, agree = _argmax_agreement(got, ref)
    assert agree >= ARGMAX_AGREE


@pytest.mark.parametrize("B,N,C", [(1, 16384, 5), (3, 1000, 3), (2, 129, 5), (5, 64, 8), (2, 4096, 5), (1, 1, 5), (1, 7, 3), (4, 31, 1),
                                   (2, 1000, 9), (2, 1500, 16), (1, 2048, 32)])
def test_eval_matches_oracle(B, N, C):
    sd = orc.synth_state(C, 100 + B + N)
    m = _model(C, sd)
    rng = np.random.default_rng(N)
    x = rng.random((B, N, 4), dtype=np.float32)
    if N == 1000:
        x[1, 700:] = 0.0     # zero-padded tail, as collate_fn would produce
    with torch.no_grad():
        logits, labels = m.predict(torch.from_numpy(x).cuda())
    got = logits.cpu().numpy()
    ref = orc.forward_eval(sd, x, dtype=np.float32)
    err = np.abs(got - ref).max() / np.abs(ref).max()
    assert err < LOGIT_TOL_REL_TO_MAX, err
    agree_all, agree = _argmax_agreement(got, ref)
    assert agree >= ARGMAX_AGREE, (agree_all, agree)
    if B * N >= 1000:
        assert agree_all >= 0.999, agree_all                       # north_star: identical argmax on >= 99.9 % of ALL points
    assert (labels.cpu().numpy() == got.argmax(-1)).all()          # integer output: bit-exact vs own logits


def test_eval_reprepares_after_weight_change():
    C = 5
    sd = orc.synth_state(C, 1)
    m = _model(C, sd)
    x = torch.rand(2, 256, 4, device="cuda")
    with torch.no_grad():
        a = m(x).clone()
        m.seg_conv4.bias.add_(1.0)
        b = m(x)
    assert torch.allclose(b - a, torch.ones_like(a), atol=1e-5)


def test_forward_rejects_bad_input():
    import pcseg_b200
    m = pcseg_b200.PointNetSegmentation(5).cuda().eval()
    with pytest.raises(ValueError):
        m(torch.rand(10, 4, device="cuda"))            # same tuple-unpack error as pcs.py:100
    with pytest.raises(RuntimeError):
        m(torch.rand(1, 10, 4))                         # CPU tensor: no fallback
    replica = torch.nn.parallel.replicate(m, [0])[0]    # what nn.DataParallel (pcs.py:211) would run
    with pytest.raises(RuntimeError, match="DataParallel"):
        replica(torch.rand(1, 10, 4, device="cuda"))


@pytest.mark.parametrize("C", [5, 13])
def test_evaluate_metrics_match_torch_and_sklearn(C):
    """model.evaluate: weighted-CE loss, accuracy counters and confusion matrix (integer, exact) of one validation batch
    (pcs.py:289-304, 319-343) against torch / sklearn applied to the SAME logits."""
    import pcseg_b200
    from sklearn.metrics import confusion_matrix, f1_score
    m = _model(C, orc.synth_state(C, 12))
    rng = np.random.default_rng(4)
    x = torch.from_numpy(rng.random((3, 700, 4), dtype=np.float32)).cuda()
    labels = torch.from_numpy(rng.integers(-1, C, (3, 700)).astype(np.int64)).cuda()
    cw = torch.tensor(([0.5, 1.0, 2.0, 0.75, 0.75] * 3)[:C], device="cuda")
    out = m.evaluate(x, labels, cw)
    logits = out["logits"]
    ref_loss = torch.nn.functional.cross_entropy(logits.view(-1, C), labels.view(-1), weight=cw, ignore_index=-1)
    assert abs(out["loss"].item() - ref_loss.item()) < 1e-5 * abs(ref_loss.item())
    valid = labels >= 0
    pred = logits.argmax(-1)
    assert int(out["valid"].item()) == int(valid.sum().item())
    assert int(out["correct"].item()) == int(((pred == labels) & valid).sum().item())
    yt, yp = labels[valid].cpu().numpy(), pred[valid].cpu().numpy()
    assert np.array_equal(out["confusion"].cpu().numpy(), confusion_matrix(yt, yp, labels=list(range(C))))
    f1, macro, weighted = pcseg_b200.f1_scores(out["confusion"])
    np.testing.assert_allclose(f1.cpu().numpy(), f1_score(yt, yp, average=None, labels=list(range(C))), atol=1e-12)
    assert abs(macro.item() - f1_score(yt, yp, average="macro")) < 1e-12
    assert abs(weighted.item() - f1_score(yt, yp, average="weighted")) < 1e-12
