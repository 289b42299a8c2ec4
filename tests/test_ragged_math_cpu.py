"""The ragged-execution scheme of the CUDA path, restated in fp64 (oracle/ragged_ref.py), reproduces the
reference's zero-padded batch EXACTLY: logits, loss and every parameter gradient.  CPU only — this pins the mathematics
(multiplicity-weighted BatchNorm sums, pre-multiplied row gradients, max-pool routing); tests/test_ragged_gpu.py checks
that the kernels implement it."""
import numpy as np
import pytest

from oracle import pointnet_oracle as orc
from oracle import ragged_ref as rr


def _padded(B, N, lengths, C, seed):
    rng = np.random.default_rng(seed)
    x = rng.random((B, N, 4))
    y = rng.integers(0, C, (B, N))
    for b, L in enumerate(lengths):
        x[b, L:] = 0.0
        y[b, L:] = -1
    return x, y


@pytest.mark.parametrize("B,N,lengths,C", [(3, 40, [40, 17, 1], 3), (2, 33, [33, 0], 5), (4, 64, [64, 63, 20, 5], 4), (2, 25, [25, 25], 3)])
def test_packed_scheme_equals_the_padded_batch(B, N, lengths, C):
    sd = orc.synth_state(C, 3 * B + N)
    x, y = _padded(B, N, lengths, C, N)
    cw = 0.5 + np.arange(C) / C

    logits, cache, _ = orc.forward_train(sd, x)                      # the reference's padded batch
    loss, dlogits = orc.weighted_ce(logits, y, cw)
    ref = orc.backward(cache, dlogits)

    xp, labp, mult, cloud = rr.pack(x, y, lengths)
    assert mult.sum() == B * N                                        # multiplicities account for every padded row
    lp, pc = rr.forward_train_packed(sd, xp, mult, cloud, B, N)
    loss_p, dlp = orc.weighted_ce(lp[None], labp[None], cw)
    got = rr.backward_packed(pc, dlp[0])

    # logits: real rows and (through the representative row) every pad row
    r = 0
    for b, L in enumerate(lengths):
        np.testing.assert_allclose(lp[r:r + L], logits[b, :L], rtol=0, atol=1e-10)
        r += L
        if L < N:
            np.testing.assert_allclose(np.broadcast_to(lp[r], (N - L, C)), logits[b, L:], rtol=0, atol=1e-10)
            r += 1
    assert abs(loss - loss_p) < 1e-12
    for name, g in ref.items():
        scale = max(np.abs(g).max(), 1e-12)
        if name.endswith(".bias") and name.split(".")[0] in orc.CONV_NAMES[:-1]:
            assert np.abs(got[name]).max() < 1e-9                      # zero in exact arithmetic (bias ahead of a BN)
            continue
        np.testing.assert_allclose(got[name], g, rtol=0, atol=1e-9 * scale + 1e-13, err_msg=name)


def test_caller_gradients_on_pad_rows_are_summed_into_the_representative_row():
    """the autograd path: an arbitrary dlogits of the PADDED logits (non-zero on pad rows too) -> packed rows"""
    B, N, lengths, C = 2, 30, [30, 11], 3
    sd = orc.synth_state(C, 9)
    x, y = _padded(B, N, lengths, C, 4)
    logits, cache, _ = orc.forward_train(sd, x)
    d = np.random.default_rng(1).normal(size=logits.shape) * 1e-2      # pad rows included
    ref = orc.backward(cache, d)
    xp, labp, mult, cloud = rr.pack(x, y, lengths)
    lp, pc = rr.forward_train_packed(sd, xp, mult, cloud, B, N)
    dp, r = [], 0
    for b, L in enumerate(lengths):
        dp.append(d[b, :L])
        if L < N:
            dp.append(d[b, L:].sum(axis=0, keepdims=True))              # what k_pack_dlogits does
    dp = np.concatenate(dp)
    # backward_packed multiplies by the multiplicity itself: hand it the per-row MEAN for the representative row
    got = rr.backward_packed(pc, dp / mult[:, None])
    for name, g in ref.items():
        if name.endswith(".bias") and name.split(".")[0] in orc.CONV_NAMES[:-1]:
            continue
        np.testing.assert_allclose(got[name], g, rtol=0, atol=1e-9 * max(np.abs(g).max(), 1e-12) + 1e-13, err_msg=name)


def test_packed_scheme_against_the_reference_fixture():
    """the ragged fixture was produced by the UNMODIFIED reference (collate_fn + model.train() + loss.backward(),
    tests/golden/make_golden.py): the packed scheme reproduces its logits, loss and gradients without computing a pad row
    more than once"""
    import os
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "case_ragged_c3.npz"))
    C, seed = int(gold["C"]), int(gold["seed"])
    sd = orc.synth_state(C, seed)
    x, y, cw = gold["x"].astype(np.float64), gold["labels"], gold["class_w"].astype(np.float64)
    B, N, _ = x.shape
    lengths = (y != -1).sum(1).tolist()
    assert min(lengths) < N
    xp, labp, mult, cloud = rr.pack(x, y, lengths)
    assert len(xp) < B * N                                             # fewer rows than the padded batch
    lp, pc = rr.forward_train_packed(sd, xp, mult, cloud, B, N)
    loss_p, dlp = orc.weighted_ce(lp[None], labp[None], cw)
    assert abs(loss_p - float(gold["loss"])) < 1e-5
    r = 0
    for b, L in enumerate(lengths):
        np.testing.assert_allclose(lp[r:r + L], gold["train_logits"][b, :L], rtol=0, atol=5e-5)
        r += L + (1 if L < N else 0)
    got = rr.backward_packed(pc, dlp[0])
    for key in gold.files:
        if key.startswith("g/"):
            name = key[2:]
            if name.endswith(".bias") and name.split(".")[0] in orc.CONV_NAMES[:-1]:
                continue
            ref = gold[key]
            np.testing.assert_allclose(got[name], ref, rtol=0, atol=1e-3 * np.abs(ref).max() + 1e-6, err_msg=name)
