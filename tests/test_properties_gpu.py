"""Size-independent properties of the CUDA path at BASELINE.json's full sizes (where the fp64 oracle is too slow)
plus behavioural checks of the training loop."""
import copy

import numpy as np
import pytest
import torch

from oracle import pointnet_oracle as orc

pytestmark = pytest.mark.gpu


def _model(C, seed, train=False):
    import pcseg_b200
    m = pcseg_b200.PointNetSegmentation(C)
    m.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in orc.synth_state(C, seed).items()})
    m = m.cuda()
    return m.train() if train else m.eval()


def test_eval_point_permutation_equivariance_full_size():
    """cfg2 shape (8 x 16 384): permuting the points of every cloud permutes the logits, bit for bit (a point's logits
    depend on the other points only through the order-independent max-pool, pcs.py:114)."""
    m = _model(5, 1)
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.rand(8, 16384, 4, device="cuda", generator=g)
    perm = torch.stack([torch.randperm(16384, device="cuda", generator=g) for _ in range(8)])
    with torch.no_grad():
        a = m(x)
        b = m(torch.gather(x, 1, perm[:, :, None].expand(-1, -1, 4)))
    assert torch.equal(torch.gather(a, 1, perm[:, :, None].expand(-1, -1, 5)), b)


def test_eval_clouds_are_independent_full_size():
    """In eval mode a cloud's logits do not depend on the other clouds of the batch (pcs.py:98-133 has no cross-cloud op
    once BatchNorm uses running statistics)."""
    m = _model(5, 2)
    x = torch.rand(8, 16384, 4, device="cuda")
    with torch.no_grad():
        full = m(x).clone()
        for b in (0, 5):
            assert torch.equal(m(x[b:b + 1].contiguous())[0], full[b])


def test_eval_large_scene_and_argmax_consistency():
    """cfg5 shape: one 1M-point cloud; integer output (labels) must equal argmax of the returned logits exactly."""
    m = _model(5, 3)
    x = torch.rand(1, 1 << 20, 4, device="cuda")
    with torch.no_grad():
        logits, labels = m.predict(x)
    assert torch.isfinite(logits).all()
    assert torch.equal(labels, logits.argmax(-1))
    # a strided sample against the fp64 oracle needs the global feature of the whole cloud -> check a small prefix cloud
    xs = x[:, :4096].contiguous()
    with torch.no_grad():
        got = m(xs).cpu().numpy()
    ref = orc.forward_eval(orc.synth_state(5, 3), xs.cpu().numpy(), dtype=np.float32)
    assert np.abs(got - ref).max() / np.abs(ref).max() < 2e-2


def test_eval_zero_padded_rows_enter_the_max_pool():
    """Padding contract (pcs.py:53-61): padded rows are real zero inputs; appending them can only change
    a cloud's logits through the max-pool, and replacing them by copies of an existing point must not change anything."""
    m = _model(3, 4)
    x = torch.rand(2, 1000, 4, device="cuda")
    xp = torch.cat([x, x[:, :24]], dim=1).contiguous()             # duplicate points: max-pool unchanged
    with torch.no_grad():
        a = m(x)
        b = m(xp)
    assert torch.equal(a, b[:, :1000])


def test_training_reduces_loss_like_the_cpu_port():
    """30 fused steps (forward + weighted CE + backward + Adam, dropout off) on a learnable synthetic task: the loss must
    fall, and track the torch-CPU port of the reference step started from the same weights."""
    import pcseg_b200
    from oracle.torch_port import TorchCpuPort
    C, B, N = 3, 4, 512
    rng = np.random.default_rng(0)
    x = rng.random((B, N, 4), dtype=np.float32)
    labels = (x[..., 0] * 3).astype(np.int64).clip(0, C - 1)      # class = slab along the first coordinate
    labels[0, 400:] = -1
    x[0, 400:] = 0
    sd = orc.synth_state(C, 77)
    # start from reference-style BN state (gamma 1, beta 0) so that both runs are well conditioned
    for k in list(sd):
        if k.startswith("bn") and k.endswith(".weight"):
            sd[k] = np.ones_like(sd[k])
        if k.startswith("bn") and k.endswith(".bias"):
            sd[k] = np.zeros_like(sd[k])
    cw = np.array([1.0, 0.7, 1.3], np.float32)
    m = pcseg_b200.PointNetSegmentation(C)
    m.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()})
    m = m.cuda().train()
    m.dropout.p = 0.0
    tr = pcseg_b200.FusedTrainer(m, class_weights=cw, lr=1e-3, weight_decay=1e-4)
    xt, lt = torch.from_numpy(x).cuda(), torch.from_numpy(labels).cuda()
    ours = [float(tr.step(xt, lt)["loss"].item()) for _ in range(30)]

    torch.set_num_threads(4)
    port = TorchCpuPort(C, state=sd)
    ref = [port.train_step(torch.from_numpy(x), torch.from_numpy(labels), torch.from_numpy(cw), dropout_p=0.0) for _ in range(30)]
    assert ours[-1] < 0.7 * ours[0], ours
    assert abs(ours[0] - ref[0]) < 2e-2 * ref[0]
    # trajectories stay close (bf16 forward noise makes them drift apart slowly)
    assert max(abs(a - b) for a, b in zip(ours, ref)) < 0.15 * ref[0], (ours[::5], ref[::5])
    # the running statistics and step counters moved like the reference's
    assert int(m.bn1.num_batches_tracked.item()) == int(sd["bn1.num_batches_tracked"]) + 30


def test_deepcopy_and_state_dict_roundtrip_after_training():
    import pcseg_b200
    m = _model(5, 6, train=True)
    x = torch.rand(2, 256, 4, device="cuda")
    m(x).sum().backward()
    clone = copy.deepcopy(m)
    m.eval(); clone.eval()
    with torch.no_grad():
        assert torch.equal(m(x), clone(x))
    fresh = pcseg_b200.PointNetSegmentation(5).cuda().eval()
    fresh.load_state_dict(m.state_dict(), strict=True)
    with torch.no_grad():
        assert torch.equal(m(x), fresh(x))


def test_eval_is_deterministic_and_input_dtype_agnostic():
    m = _model(5, 8)
    x = torch.rand(3, 777, 4, device="cuda")
    with torch.no_grad():
        a = m(x)
        b = m(x.double())                       # converted to fp32 like any torch module input would be
        c = m(x.transpose(0, 1).contiguous().transpose(0, 1))     # non-contiguous view
    assert torch.equal(a, b) and torch.equal(a, c)


def test_cuda_graph_step_equals_eager_step():
    """FusedTrainer replays a captured CUDA graph from the third step on; it must train exactly like eager launches
    (same kernels, device-resident seed / Adam state), up to atomics-order noise."""
    import pcseg_b200
    C, B, N = 5, 4, 640
    rng = np.random.default_rng(3)
    x = torch.from_numpy(rng.random((B, N, 4), dtype=np.float32)).cuda()
    labels = torch.from_numpy(rng.integers(0, C, (B, N)).astype(np.int64)).cuda()
    losses = {}
    params = {}
    for use_graph in (False, True):
        m = _model(C, 21, train=True)
        m.dropout.p = 0.0
        tr = pcseg_b200.FusedTrainer(m, class_weights=torch.ones(C), use_cuda_graph=use_graph)
        losses[use_graph] = [float(tr.step(x, labels)["loss"].item()) for _ in range(8)]
        params[use_graph] = tr.flat["params"].clone()
        assert (tr._graph is not None) == use_graph
        assert int(m.bn1.num_batches_tracked.item()) == 7 + 8
        assert int(tr.state.view(torch.int64)[1].item()) == 8           # device-side Adam step counter
    # two EAGER runs already differ by ~0.5 % after a few steps (fp64 atomics order -> last-bit differences in the batch
    # statistics -> individual bf16 roundings flip -> amplified by the ill-conditioned train-mode network)
    assert abs(losses[False][0] - losses[True][0]) < 1e-5 * abs(losses[False][0])
    for i, (a, b) in enumerate(zip(losses[False], losses[True])):
        assert abs(a - b) < (2e-2 if i < 3 else 6e-2) * abs(a), (losses[False], losses[True])
    assert (params[False] - params[True]).abs().max().item() < 1.7e-2        # 8 Adam steps of lr = 1e-3 each way
    # dropout masks must differ from step to step under graph replay (seed advances on the device)
    m = _model(C, 22, train=True)
    tr = pcseg_b200.FusedTrainer(m, class_weights=torch.ones(C))
    outs = []
    for _ in range(5):
        tr.step(x, labels)
        outs.append(tr.last_logits.clone())
    assert tr._graph is not None
    assert (outs[-1] - outs[-2]).abs().max().item() > 1e-4


def test_prefetched_host_batches_train_like_device_batches():
    """FusedTrainer.prefetch / step_prefetched (pinned host -> device on a copy stream, double buffered) feeds the same
    data as passing device tensors."""
    import pcseg_b200
    C, B, N = 3, 4, 384          # (4 clouds: with 2 the pooled-feature gradients cancel exactly and the trajectory is chaotic)
    rng = np.random.default_rng(9)
    batches = [(torch.from_numpy(rng.random((B, N, 4), dtype=np.float32)).pin_memory(),
                torch.from_numpy(rng.integers(-1, C, (B, N)).astype(np.int64)).pin_memory()) for _ in range(5)]
    losses = []
    for mode in ("device", "prefetch"):
        m = _model(C, 31, train=True)
        m.dropout.p = 0.0
        tr = pcseg_b200.FusedTrainer(m, class_weights=torch.ones(C), use_cuda_graph=False)
        out = []
        if mode == "device":
            for xb, lb in batches:
                out.append(float(tr.step(xb.cuda(), lb.cuda())["loss"].item()))
        else:
            ticket = tr.prefetch(*batches[0])
            for i in range(len(batches)):
                nxt = tr.prefetch(*batches[i + 1]) if i + 1 < len(batches) else None
                out.append(float(tr.step_prefetched(ticket)["loss"].item()))
                ticket = nxt
        losses.append(out)
    # identical data -> identical first loss; afterwards the two runs drift like any two runs of the same data do (fp64
    # atomics order -> last-bit differences in the batch statistics -> flipped bf16 roundings, amplified by the train-mode
    # network: a few per cent after five steps, occasionally more than 2 %): wrong or stale batches would differ by O(1)
    assert abs(losses[0][0] - losses[1][0]) < 1e-6 * abs(losses[0][0])
    assert abs(losses[0][1] - losses[1][1]) < 2e-2 * abs(losses[0][1]), losses
    for a, b in zip(*losses):
        assert abs(a - b) < 8e-2 * abs(a), losses


def test_training_with_dropout_converges_and_eval_improves():
    """150 fused steps with dropout 0.3 on a learnable task (class = slab of the x coordinate): finite loss throughout,
    clear decrease, eval-mode accuracy (running statistics, fused inference head) well above chance afterwards."""
    import pcseg_b200
    C, B, N = 4, 8, 1024
    rng = np.random.default_rng(11)

    def batch():
        x = rng.random((B, N, 4), dtype=np.float32)
        y = np.minimum((x[..., 0] * C).astype(np.int64), C - 1)
        return torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda()

    torch.manual_seed(0)
    m = pcseg_b200.PointNetSegmentation(C).cuda().train()        # reference-style default initialisation
    tr = pcseg_b200.FusedTrainer(m, class_weights=torch.ones(C), lr=1e-3, weight_decay=1e-4)
    losses = []
    for _ in range(150):
        x, y = batch()
        losses.append(tr.step(x, y)["loss"])
    losses = torch.stack(losses).cpu().numpy()
    assert np.isfinite(losses).all()
    assert losses[-10:].mean() < 0.5 * losses[:5].mean(), (losses[:5], losses[-10:])
    m.eval()
    x, y = batch()
    out = m.evaluate(x, y)
    acc = out["correct"].item() / out["valid"].item()
    assert acc > 0.8, acc
    f1, macro, _ = pcseg_b200.f1_scores(out["confusion"])
    assert macro.item() > 0.75
    for bn in (m.bn1, m.bn_global, m.bn_seg3):
        assert torch.isfinite(bn.running_mean).all() and (bn.running_var > 0).all()
        assert int(bn.num_batches_tracked.item()) == 150


def test_variable_length_batches_share_one_workspace():
    """Every DataLoader batch of the reference has its own max_points (pcs.py:50).  Bindings of different shapes share one
    grow-only workspace; results must equal those of a fresh model, for interleaved eval and training shapes."""
    import pcseg_b200
    C = 5
    sd = orc.synth_state(C, 41)
    shapes = [(2, 300), (3, 1000), (2, 300), (1, 77), (3, 1000), (4, 512)]
    rng = np.random.default_rng(5)
    xs = [torch.from_numpy(rng.random((b, n, 4), dtype=np.float32)).cuda() for b, n in shapes]

    def fresh():
        m = pcseg_b200.PointNetSegmentation(C)
        m.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()})
        return m.cuda()

    shared = fresh().eval()
    with torch.no_grad():
        outs = [shared(x).clone() for x in xs]
        eng = shared._get_engine(xs[0].device)
        assert len({b.ws_ptr for b in eng.bindings.values()}) == 1          # one workspace, many bindings
        for x, o in zip(xs, outs):
            assert torch.equal(fresh().eval()(x), o)

    # training: forward/backward pairs of different shapes, gradients equal a fresh model's
    shared = fresh().train()
    shared.dropout.p = 0.0
    for x in xs[:4]:
        for m in (shared, fresh().train()):
            m.dropout.p = 0.0
            m.zero_grad()
            m(x).square().mean().backward()
            g = torch.cat([p.grad.reshape(-1) for p in m.parameters()])
            if m is shared:
                g_shared = g.clone()
                # undo the running-stat update so that the next shape starts from the same state as a fresh model
                shared.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()})
        cos = torch.nn.functional.cosine_similarity(g_shared, g, dim=0).item()
        assert cos > 0.9995, cos


def test_point_sharded_inference_is_bit_identical():
    """SURVEY §8(e), one cloud over several GPUs: slices of the points of the same clouds, the pooled feature reduced with
    MAX between them (here: the collective is emulated on one GPU), give bit-identical logits and labels."""
    m = _model(5, 31)
    x = torch.rand(2, 3000, 4, device="cuda")
    with torch.no_grad():
        full, full_lab = m.predict(x)
    shards = [x[:, :1000].contiguous(), x[:, 1000:1001].contiguous(), x[:, 1001:].contiguous()]
    pooled = []
    for xs in shards:                                   # pass 1: every "rank" computes its local pooled feature
        m.predict_point_sharded(xs, reduce_max=lambda p: pooled.append(p.clone()))
    gmax = torch.stack(pooled).max(dim=0)[0]
    outs = [m.predict_point_sharded(xs, reduce_max=lambda p: p.copy_(gmax)) for xs in shards]
    assert torch.equal(torch.cat([o[0] for o in outs], dim=1), full)
    assert torch.equal(torch.cat([o[1] for o in outs], dim=1), full_lab)
    # without the exchange a slice only sees its own points
    alone = m.predict_point_sharded(shards[0], reduce_max=lambda p: None)[0]
    assert torch.equal(alone, m(shards[0]))


def test_predict_stream_matches_predict():
    """PredictStream (copies on side streams, two batches in flight) returns what predict() returns, batch after batch"""
    import pcseg_b200
    m = _model(5, 32)
    ps = pcseg_b200.PredictStream(m)
    g = torch.Generator().manual_seed(0)
    batches = [torch.rand(2, 1500, 4, generator=g).pin_memory() for _ in range(5)]
    expect = []
    with torch.no_grad():
        for xb in batches:
            expect.append(m.predict(xb.cuda())[1].cpu())
    prev, got = None, []
    for xb in batches:
        t = ps.submit(xb)
        if prev is not None:
            got.append(ps.result(prev).clone())
        prev = t
    labels, logits = ps.result(prev, want_logits=True)
    got.append(labels.clone())
    assert logits.shape == (2, 1500, 5)
    for a, b in zip(expect, got):
        assert torch.equal(a, b)
    # ragged batch through the same stream
    x = batches[0].clone()
    x[1, 400:] = 0
    t = ps.submit(x, lengths=[1500, 400])
    with torch.no_grad():
        assert torch.equal(ps.result(t), m.predict(x.cuda())[1].cpu())


def test_optimizer_state_has_the_torch_adam_layout_and_resumes():
    """FusedTrainer.checkpoint_dict writes what the reference saves at pcs.py:373-382: `optimizer_state_dict` must load into
    torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=1e-4) (pcs.py:217), and load_optimizer_state must resume a
    run exactly where it stopped (moments, step count, hyper-parameters)."""
    import pcseg_b200
    C, B, N = 3, 4, 256
    rng = np.random.default_rng(11)
    x = torch.from_numpy(rng.random((B, N, 4), dtype=np.float32)).cuda()
    labels = torch.from_numpy(rng.integers(0, C, (B, N)).astype(np.int64)).cuda()
    m = _model(C, 13, train=True)
    m.dropout.p = 0.0
    tr = pcseg_b200.FusedTrainer(m, class_weights=torch.ones(C), lr=2e-3, weight_decay=1e-4, use_cuda_graph=False)
    for _ in range(3):
        tr.step(x, labels)
    ckpt = tr.checkpoint_dict(epoch=7, train_loss=0.5, val_loss=0.6, f1_class2=0.1, f1_per_class=[0.1, 0.2, 0.3])
    assert set(ckpt) >= {"epoch", "model_state_dict", "optimizer_state_dict", "train_loss", "val_loss", "f1_class2", "f1_per_class", "num_classes"}
    osd = ckpt["optimizer_state_dict"]
    # (1) the reference's optimizer accepts it
    ref_model = pcseg_b200.PointNetSegmentation(C).cuda()
    ref_model.load_state_dict(ckpt["model_state_dict"], strict=True)
    opt = torch.optim.Adam(ref_model.parameters(), lr=0.001, weight_decay=1e-4)
    opt.load_state_dict(osd)
    st = opt.state_dict()
    assert len(st["state"]) == 38 and st["param_groups"][0]["lr"] == 2e-3
    assert float(st["state"][0]["step"]) == 3.0
    p0 = next(iter(ref_model.parameters()))
    assert torch.equal(opt.state[p0]["exp_avg"].cpu(), osd["state"][0]["exp_avg"].cpu())
    # (2) a fresh trainer resumed from the checkpoint takes the same next step as the original one
    m2 = _model(C, 13, train=True)
    m2.dropout.p = 0.0
    m2.load_state_dict(ckpt["model_state_dict"], strict=True)
    tr2 = pcseg_b200.FusedTrainer(m2, class_weights=torch.ones(C), use_cuda_graph=False)
    tr2.load_optimizer_state(osd)
    assert tr2.step_count == 3 and tr2.lr == 2e-3
    a = tr.step(x, labels)["loss"].item()
    b = tr2.step(x, labels)["loss"].item()
    assert abs(a - b) < 1e-4 * abs(a)
    # (Adam normalises every element's step to ~lr: compare the parameters with an lr-sized tolerance)
    assert (tr.flat["params"] - tr2.flat["params"]).abs().max().item() < 2.5 * 2e-3
    assert int(tr2.state.view(torch.int64)[1].item()) == 4
