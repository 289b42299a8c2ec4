"""Ragged (un-padded) execution, SURVEY §8(f) rank 1: a zero-padded batch of the reference's collate_fn (pcs.py:44-63) run on
its real points only must give what the padded batch gives (padding contract, SURVEY §8 row P: pad rows are real inputs of
the BatchNorm batch statistics and of the max-pool)."""
import os

import numpy as np
import pytest
import torch

from oracle import pointnet_oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _legacy_dense_step(monkeypatch):
    """The packed (ragged) step runs the kernels that materialise every y / dy; dense batches default to the folded
    BatchNorm path, a different (equally valid) set of bf16 rounding points.  These tests isolate the
    PACKING scheme, so their padded reference steps run the same kernels as the packed ones."""
    monkeypatch.setenv("PCSEG_FOLDED", "0")
    monkeypatch.setenv("PCSEG_RAGGED_MIN_PAD", "0")          # always take the packed path here, however little padding there is
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _model(C, seed, train=False, p_drop=None):
    import pcseg_b200
    m = pcseg_b200.PointNetSegmentation(C)
    m.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in orc.synth_state(C, seed).items()})
    m = m.cuda()
    if p_drop is not None:
        m.dropout.p = p_drop
    return m.train() if train else m.eval()


def _padded_batch(B, N, lengths, C, seed):
    """what collate_fn builds: zero rows / label -1 / mask False after the first lengths[b] rows"""
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = torch.rand(B, N, 4, device="cuda", generator=g)
    y = torch.randint(0, C, (B, N), device="cuda", generator=g)
    for b, L in enumerate(lengths):
        x[b, L:] = 0
        y[b, L:] = -1
    return x, y


CASES = [
    (3, 1000, [1000, 517, 1], 3),        # full cloud with a ragged tail tile, mid cloud, single point
    (2, 256, [256, 0], 5),               # N multiple of 128, one empty cloud
    (4, 4096, [4095, 4096, 129, 128], 5),
    (1, 100, [100], 3),                  # nothing padded, N not a multiple of 128 (filler rows are duplicates)
    (5, 777, [1, 2, 3, 776, 777], 8),
    (3, 900, [900, 450, 2], 12),         # more than 8 classes: the wide head kernels
]


@pytest.mark.parametrize("B,N,lengths,C", CASES)
def test_ragged_eval_is_bit_identical_to_the_padded_batch(B, N, lengths, C):
    m = _model(C, 11)
    x, _ = _padded_batch(B, N, lengths, C, 5)
    with torch.no_grad():
        dense, dense_lab = m.predict(x)
        rag, rag_lab = m.predict(x, lengths=lengths)
        rag_fwd = m(x, lengths=torch.tensor(lengths))          # tensors are accepted as well
    assert rag.shape == (B, N, C)
    assert torch.equal(dense, rag)                              # real rows AND pad rows
    assert torch.equal(dense_lab, rag_lab)
    assert torch.equal(rag, rag_fwd)


def test_ragged_eval_ignores_the_content_of_pad_rows():
    """rows at and after lengths[b] are never read: the result is the one of the ZERO-padded batch"""
    m = _model(5, 12)
    lengths = [700, 31, 1024]
    x, _ = _padded_batch(3, 1024, lengths, 5, 6)
    junk = x.clone()
    for b, L in enumerate(lengths):
        junk[b, L:] = 123.0
    with torch.no_grad():
        assert torch.equal(m(x), m(junk, lengths=lengths))


def test_ragged_eval_matches_the_reference_golden_vectors():
    d = np.load(os.path.join(GOLDEN, "case_ragged_c3.npz"))
    C, seed = int(d["C"]), int(d["seed"])
    labels = d["labels"]
    lengths = (labels != -1).sum(1).tolist()
    assert min(lengths) < labels.shape[1]                       # the fixture really is ragged
    m = _model(C, seed)
    with torch.no_grad():
        got = m(torch.from_numpy(d["x"]).cuda(), lengths=lengths).cpu().numpy()
    ref = d["eval_logits"]
    assert np.abs(got - ref).max() / np.abs(ref).max() < 2e-2   # bf16 operands, fp32 accumulation (DESIGN.md numerics)


def test_ragged_eval_metrics_and_shared_workspace():
    """evaluate() on a ragged batch; dense and ragged calls of different shapes interleave on one module"""
    m = _model(5, 13)
    for (B, N, lengths) in [(2, 640, [640, 17]), (3, 300, [5, 300, 299]), (2, 640, [1, 639])]:
        x, y = _padded_batch(B, N, lengths, 5, 7)
        with torch.no_grad():
            a = m.evaluate(x, y, class_weights=[1, 2, 3, 4, 5])
            b = m.evaluate(x, y, class_weights=[1, 2, 3, 4, 5], lengths=lengths)
        assert torch.equal(a["logits"], b["logits"])
        assert torch.equal(a["confusion"], b["confusion"])
        assert a["valid"].item() == sum(lengths) == b["valid"].item()
        assert a["loss"].item() == b["loss"].item()


def test_ragged_argument_checks():
    m = _model(3, 14)
    x = torch.rand(2, 64, 4, device="cuda")
    with pytest.raises(ValueError):
        m(x, lengths=[64])
    with pytest.raises(ValueError):
        m(x, lengths=[64, 65])


def test_ragged_autograd_path_equals_the_padded_one():
    """the reference loop itself (criterion on the module's output, loss.backward(), pcs.py:244-254) with lengths=: the
    caller's dlogits of the padded logits are packed (pad rows summed into the representative row)"""
    C, B, N, lengths = 3, 3, 700, [700, 250, 9]
    x, y = _padded_batch(B, N, lengths, C, 17)
    crit = torch.nn.CrossEntropyLoss(weight=torch.tensor([1.0, 2.0, 0.5], device="cuda"), ignore_index=-1)
    res = []
    for ls in (None, lengths):
        m = _model(C, 23, train=True, p_drop=0.0)
        out = m(x, lengths=ls)
        loss = crit(out.contiguous().view(-1, C), y.view(-1))
        # a loss term that also touches the pad rows' logits (their gradients must be summed into the representative row)
        loss = loss + 1e-3 * out.pow(2).mean()
        loss.backward()
        res.append((loss.item(), torch.cat([p.grad.reshape(-1) for p in m.parameters()]).double()))
    assert abs(res[0][0] - res[1][0]) < 2e-4 * max(1.0, abs(res[0][0]))
    cos = torch.nn.functional.cosine_similarity(res[0][1], res[1][1], dim=0).item()
    assert cos > 0.999, cos
    assert abs(res[0][1].norm().item() - res[1][1].norm().item()) < 2e-2 * res[0][1].norm().item()


def _one_step(C, B, N, lengths, use_lengths, p_drop=0.0, steps=1, seed=21):
    import pcseg_b200
    m = _model(C, seed, train=True, p_drop=p_drop)
    tr = pcseg_b200.FusedTrainer(m, class_weights=[1.0 + 0.5 * k for k in range(C)], lr=1e-3, use_cuda_graph=False)
    losses = []
    x, y = _padded_batch(B, N, lengths, C, 9)
    for _ in range(steps):
        out = tr.step(x, y, lengths=lengths if use_lengths else None)
        losses.append(out["loss"].item())
    grads = m._flat["grads"].clone()
    return m, tr, losses, grads, out


@pytest.mark.parametrize("B,N,lengths,C", [(3, 1000, [1000, 517, 64], 3), (4, 2048, [2048, 2047, 700, 1], 5), (2, 300, [300, 300], 5),
                                              (3, 800, [800, 300, 31], 12)])
def test_ragged_training_step_equals_the_padded_step(B, N, lengths, C):
    """dropout off: loss, logits, BatchNorm running statistics and every parameter gradient of the packed step equal the
    padded step's.  Not bit-exact: the batch sums are accumulated in a different order (and the filler rows are removed
    again in fp64), so bf16 roundings may flip; the tolerances are far below the bf16-vs-fp64 distance of the path
    (tests/test_train_gpu.py)."""
    md, trd, ld, gd, outd = _one_step(C, B, N, lengths, False)
    mr, trr, lr_, gr, outr = _one_step(C, B, N, lengths, True)
    assert abs(ld[0] - lr_[0]) < 2e-4 * max(1.0, abs(ld[0]))
    assert outd["valid"].item() == outr["valid"].item() == sum(lengths)
    assert abs(outd["correct"].item() - outr["correct"].item()) <= max(2, sum(lengths) // 500)
    a, b = trd.last_logits, trr.last_logits
    assert a.shape == b.shape == (B, N, C)
    assert (a - b).abs().max().item() < 2e-2 * a.abs().max().item()
    assert ((a - b).pow(2).mean().sqrt() / a.pow(2).mean().sqrt()).item() < 2e-3
    # running statistics (they include the pad rows, weighted by their count)
    assert torch.allclose(md._flat["bn"], mr._flat["bn"], rtol=2e-3, atol=1e-5)
    # gradients, tensor by tensor
    for t, (o, n) in enumerate(md._flat["offs"]):
        ga, gb = gd[o:o + n].double(), gr[o:o + n].double()
        na, nb = ga.norm().item(), gb.norm().item()
        if t < 18 and t % 2 == 1:
            # biases of the 9 convolutions ahead of a train-mode BN: mathematically zero, rounding noise in both runs
            assert na < 1e-3 * gd.double().norm().item() and nb < 1e-3 * gr.double().norm().item()
            continue
        cos = (ga @ gb).item() / (na * nb)
        assert cos > 0.995, (o, n, cos)
        assert abs(na - nb) < 2e-2 * na, (o, n, na, nb)


def test_ragged_training_matches_the_fp64_oracle_like_the_padded_path():
    """the packed step against the fp64 oracle run on the PADDED batch: it must sit as close to the oracle as the padded
    CUDA step does (train-mode bf16 error bounds: tests/test_train_gpu.py; this small, heavily padded batch is harder)"""
    C, B, N, lengths = 3, 3, 512, [512, 200, 33]
    x, y = _padded_batch(B, N, lengths, C, 9)
    st = orc.synth_state(C, 21)
    cw = np.array([1.0 + 0.5 * k for k in range(C)])
    logits, cache, _ = orc.forward_train(st, x.cpu().numpy().astype(np.float64))
    loss, dlogits = orc.weighted_ce(logits, y.cpu().numpy(), cw)
    errs = {}
    for use_lengths in (False, True):
        m, tr, losses, grads, out = _one_step(C, B, N, lengths, use_lengths)
        assert abs(losses[0] - loss) < 0.05 * abs(loss)
        got = tr.last_logits.cpu().numpy()
        assert np.abs(got - logits).max() < 0.30 * np.abs(logits).max()
        errs[use_lengths] = np.sqrt(((got - logits) ** 2).mean()) / np.sqrt((logits ** 2).mean())
    assert errs[True] < 0.12 and errs[True] < 1.25 * errs[False] + 0.01, errs


def test_ragged_training_trajectory_and_dropout():
    """several optimizer steps: the packed run tracks the padded run without dropout, and trains with dropout"""
    C, B, N, lengths = 3, 4, 640, [640, 320, 100, 639]
    _, _, ld, _, _ = _one_step(C, B, N, lengths, False, steps=12)
    _, _, lr_, _, _ = _one_step(C, B, N, lengths, True, steps=12)
    assert ld[-1] < ld[0] and lr_[-1] < lr_[0]
    assert max(abs(a - b) for a, b in zip(ld, lr_)) < 0.03 * ld[0]
    _, _, lp, _, _ = _one_step(C, B, N, lengths, True, p_drop=0.3, steps=12)
    assert all(np.isfinite(lp)) and lp[-1] < lp[0]


def test_ragged_and_dense_steps_interleave():
    """a dense (CUDA-graph) trainer keeps working when ragged steps are mixed in"""
    import pcseg_b200
    C = 3
    m = _model(C, 22, train=True, p_drop=0.0)
    tr = pcseg_b200.FusedTrainer(m, lr=1e-3)
    x, y = _padded_batch(2, 512, [512, 100], C, 3)
    losses = []
    for i in range(8):
        out = tr.step(x, y, lengths=[512, 100] if i % 2 else None)
        losses.append(out["loss"].item())
    assert all(np.isfinite(losses)) and losses[-1] < losses[0]


def test_ragged_batches_of_changing_padded_length_share_one_binding():
    """every DataLoader batch is padded to ITS longest cloud (pcs.py:50): ragged calls bind a capacity bucket, not the
    exact padded length, and still reproduce the padded batch (whose BN / max-pool semantics depend on that length)"""
    m = _model(3, 15)
    for N, lengths in [(1000, [1000, 400]), (1500, [7, 1500]), (3000, [2999, 3000]), (1001, [1001, 1])]:
        x, _ = _padded_batch(2, N, lengths, 3, N)
        with torch.no_grad():
            assert torch.equal(m(x), m(x, lengths=lengths))
    eval_keys = [k for k in m._engine.bindings if not k[2]]
    assert sum(1 for k in eval_keys if k[1] == 4096) == 1           # one capacity-bucketed binding served all four
    # training: two padded lengths through one binding, each equal to its padded step
    for N, lengths in [(900, [900, 333, 10]), (1200, [5, 1200, 600])]:
        _, trd, ld, gd, _ = _one_step(3, 3, N, lengths, False)
        _, trr, lr_, gr, _ = _one_step(3, 3, N, lengths, True)
        assert abs(ld[0] - lr_[0]) < 2e-4 * max(1.0, abs(ld[0]))
        cos = torch.nn.functional.cosine_similarity(gd.double(), gr.double(), dim=0).item()
        assert cos > 0.999, cos
