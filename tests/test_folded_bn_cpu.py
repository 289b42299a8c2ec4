"""The Gram-predicted BatchNorm statistics and the folded BatchNorm backward that the CUDA training step uses for conv5 and
global_feat (oracle/folded_bn_ref.py) reproduce the reference's autograd result EXACTLY (fp64): parameter
gradients of both layers, the gradient flowing into conv4's output, and the batch statistics themselves.  CPU only — this
pins the algebra; tests/test_layerwise_gpu.py / test_train_gpu.py check that the kernels implement it."""
import numpy as np
import pytest

from oracle import folded_bn_ref as fb
from oracle import pointnet_oracle as orc


@pytest.mark.parametrize("B,N,C,seed", [(2, 96, 3, 1), (3, 50, 5, 2), (1, 200, 4, 3)])
def test_folded_backward_equals_autograd(B, N, C, seed):
    sd = orc.synth_state(C, seed)
    rng = np.random.default_rng(seed)
    x = rng.random((B, N, 4))
    y = rng.integers(0, C, (B, N))
    y[0, -3:] = -1
    cw = 0.5 + np.arange(C) / C
    logits, cache, _ = orc.forward_train(sd, x)
    _, dlogits = orc.weighted_ce(logits, y, cw)
    taps = {}
    ref = orc.backward(cache, dlogits, taps)
    got = fb.trunk_tail_backward(cache, taps["dg"])
    # (1) predicted statistics == batch statistics
    np.testing.assert_allclose(got["_pred"]["invstd5"], got["_pred"]["invstd5_ref"], rtol=1e-9)
    np.testing.assert_allclose(got["_pred"]["invstd4"], got["_pred"]["invstd4_ref"], rtol=1e-9)
    # (2) parameter gradients
    for name in ("global_feat.weight", "bn_global.weight", "bn_global.bias", "conv5.weight", "bn5.weight", "bn5.bias"):
        scale = max(np.abs(ref[name]).max(), 1e-12)
        np.testing.assert_allclose(got[name], ref[name], rtol=0, atol=1e-9 * scale, err_msg=name)
    for name in ("global_feat.bias", "conv5.bias"):                 # zero in exact arithmetic (bias ahead of a BN)
        assert np.abs(got[name]).max() < 1e-9 and np.abs(ref[name]).max() < 1e-9
    # (3) data gradient into conv4's output
    scale = np.abs(taps["dx:conv5"]).max()
    np.testing.assert_allclose(got["_da3"], taps["dx:conv5"], rtol=0, atol=1e-9 * scale)


def test_predicted_stats_with_bias_and_negative_gamma():
    rng = np.random.default_rng(0)
    a = np.maximum(rng.normal(size=(500, 24)), 0)
    W = rng.normal(size=(40, 24))
    b = rng.normal(size=40)
    y = a @ W.T + b
    s, G = fb.gram(a)
    s1, s2 = fb.predicted_stats(W, b, s, G, 500)
    np.testing.assert_allclose(s1, y.sum(0), rtol=1e-12)
    np.testing.assert_allclose(s2, (y * y).sum(0), rtol=1e-12)
