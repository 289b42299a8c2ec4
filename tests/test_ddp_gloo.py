"""Host-side data-parallel protocol on CPU (gloo, world_size 2): the collectives FusedTrainer issues
(asynchronous all-reduce of the loss normaliser, un-normalised backward, bucketed SUM all-reduce of the flat gradient
arena, division by the global normaliser in the optimizer; and the blocking normalise-before-backward form) reproduce the
reference's single-process nn.DataParallel step (pcs.py:209-211, 244-254): one weighted-mean loss over the whole
global batch, per-replica BatchNorm statistics, summed replica gradients.  Compute is done by the CPU port in
oracle/ (tests may use the oracle; the product path needs a GPU)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _flat_grads(port):
    return torch.cat([port.p[k].grad.reshape(-1) for k in port.params])


def _worker(rank, world, port_file, q, deferred=True):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = port_file
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import importlib.util
    spec = importlib.util.spec_from_file_location("trainer_mod", os.path.join(ROOT, "point-cloud-cnn-segmentation_b200", "trainer_protocol.py"))
    proto = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(proto)
    from oracle.torch_port import TorchCpuPort
    import torch.nn.functional as F

    torch.set_num_threads(1)
    C, B, N = 3, 4, 96
    rng = np.random.default_rng(0)
    x = torch.from_numpy(rng.random((B, N, 4), dtype=np.float32))
    labels = torch.from_numpy(rng.integers(-1, C, (B, N)).astype(np.int64))
    cw = torch.tensor([1.0, 0.5, 2.0])
    shard = slice(rank * B // world, (rank + 1) * B // world)       # DataParallel's dim-0 chunking
    port = TorchCpuPort(C, seed=7)
    logits = port.forward(x[shard], True, dropout_p=0.0)
    lab = labels[shard].reshape(-1)
    valid = lab >= 0
    wsum = cw[lab[valid]].sum().double().reshape(1)
    flat = torch.zeros(sum(port.p[k].numel() for k in port.params))
    sync = proto.GradSync(flat)
    loss_sum = F.cross_entropy(logits.view(-1, C), lab, weight=cw, ignore_index=-1, reduction="sum")
    if deferred:
        sync.launch_tensor(wsum)                                      # in flight while backward runs; joined by wait()
        loss_sum.backward()                                           # un-normalised loss (pcseg_backward with wsum = 1)
    else:
        sync.reduce_normaliser(wsum)                                  # global normaliser BEFORE backward
        (loss_sum / wsum.float()).backward()
    flat.copy_(_flat_grads(port))
    n = flat.numel()
    # two slices per bucket, like grad_buckets(): exercises the coalesced launch (one collective per bucket)
    early, late = [(n // 3, n // 2), (3 * n // 4, n)], [(0, n // 3), (n // 2, 3 * n // 4)]
    sync.launch(early)
    sync.launch(late)
    sync.wait()
    if deferred:
        flat /= wsum.float()                                          # k_adam's grad_div
    if rank == 0:
        q.put(flat.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
@pytest.mark.parametrize("deferred", [True, False])
def test_two_rank_protocol_equals_dataparallel_semantics(tmp_path, deferred):
    sys.path.insert(0, ROOT)
    from oracle.torch_port import TorchCpuPort
    import torch.nn.functional as F

    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port_no = str(29500 + (os.getpid() % 2000) + (2000 if deferred else 0))
    procs = [ctx.Process(target=_worker, args=(r, world, port_no, q, deferred)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0

    # emulated single-process DataParallel: each replica normalises with ITS OWN batch statistics, the loss is one
    # weighted mean over all gathered logits, replica gradients are added
    C, B, N = 3, 4, 96
    rng = np.random.default_rng(0)
    x = torch.from_numpy(rng.random((B, N, 4), dtype=np.float32))
    labels = torch.from_numpy(rng.integers(-1, C, (B, N)).astype(np.int64))
    cw = torch.tensor([1.0, 0.5, 2.0])
    torch.set_num_threads(1)
    ports = [TorchCpuPort(C, seed=7) for _ in range(world)]
    outs = [ports[r].forward(x[r * B // world:(r + 1) * B // world], True, dropout_p=0.0) for r in range(world)]
    loss = F.cross_entropy(torch.cat(outs).view(-1, C), labels.view(-1), weight=cw, ignore_index=-1)
    loss.backward()
    ref = sum(_flat_grads(p) for p in ports).numpy()
    np.testing.assert_allclose(got, ref, rtol=1e-4, atol=1e-6)


def test_grad_buckets_partition_the_arena():
    import importlib.util
    spec = importlib.util.spec_from_file_location("trainer_mod", os.path.join(ROOT, "point-cloud-cnn-segmentation_b200", "trainer_protocol.py"))
    proto = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(proto)
    import ctypes
    lib = ctypes.CDLL(os.path.join(ROOT, "point-cloud-cnn-segmentation_b200", "lib", "libpcseg_b200.so"))
    lib.pcseg_param_offset.restype = ctypes.c_longlong
    lib.pcseg_param_numel.restype = ctypes.c_longlong
    lib.pcseg_param_count.restype = ctypes.c_longlong
    for C in (3, 5, 8):
        offs = [(lib.pcseg_param_offset(C, t), lib.pcseg_param_numel(C, t)) for t in range(38)]
        total = lib.pcseg_param_count(C)
        early, late = proto.grad_buckets(offs, total)
        cover = np.zeros(total, np.int32)
        for a, b in early + late:
            cover[a:b] += 1
        assert (cover == 1).all()
        # early bucket = tensors finished by backward phase 1: global_feat..seg_conv4 (10..19) and bn_global..bn_seg3 (30..37)
        mask = np.zeros(total, bool)
        for a, b in early:
            mask[a:b] = True
        for t, (o, n) in enumerate(offs):
            assert mask[o:o + n].all() == (10 <= t <= 19 or t >= 30), t


def _shard_worker(rank, world, port_no, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = port_no
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import pointnet_oracle as orc
    C, B, N = 3, 2, 200
    sd = orc.synth_state(C, 5)
    x = np.random.default_rng(1).random((B, N, 4))
    cut = [0, 77, N]                                                 # unequal slices of the points of the same clouds
    xs = x[:, cut[rank]:cut[rank + 1]]
    _, pooled = orc.forward_eval(sd, xs, return_pooled=True)         # part 1 on the local points
    t = torch.from_numpy(np.ascontiguousarray(pooled))
    dist.all_reduce(t, op=dist.ReduceOp.MAX)                         # the one exchange (model.predict_point_sharded)
    logits = orc.forward_eval(sd, xs, pooled=t.numpy())              # part 2 on the local points
    q.put((rank, logits))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_point_sharded_inference_protocol():
    """SURVEY §8(e) within-one-cloud sharding: slices of the points + MAX all-reduce of the pooled feature == the whole cloud"""
    sys.path.insert(0, ROOT)
    from oracle import pointnet_oracle as orc
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port_no = str(31500 + (os.getpid() % 2000))
    procs = [ctx.Process(target=_shard_worker, args=(r, world, port_no, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=240) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    C, B, N = 3, 2, 200
    x = np.random.default_rng(1).random((B, N, 4))
    ref = orc.forward_eval(orc.synth_state(C, 5), x)
    np.testing.assert_allclose(np.concatenate([got[0], got[1]], axis=1), ref, rtol=0, atol=1e-10)
