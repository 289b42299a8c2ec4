"""torchrun --nproc-per-node G tools/ddp_equiv_check.py: G data-parallel ranks (FusedTrainer, NCCL) against ONE process that
emulates the reference's nn.DataParallel (pcs.py:209-211, 244-254) on rank 0: per-replica BatchNorm statistics, one global
weighted-mean loss, summed replica gradients.  Also checks that replicas that were constructed from DIFFERENT seeds start
from rank 0's parameters (broadcast at construction) and that every rank draws its own dropout stream.
Prints one JSON line on rank 0 and exits non-zero when a check fails (tests/test_multigpu_gpu.py runs it)."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import pcseg_b200  # noqa: E402

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
C, B, N = 5, 4 * world, 2048
g = torch.Generator().manual_seed(1)
x = torch.rand(B, N, 4, generator=g).to(dev)
y = torch.randint(-1, C, (B, N), generator=g).to(dev)
cw = torch.tensor([0.5, 1.0, 2.0, 0.75, 1.5], device=dev)
per = B // world


def model(seed):
    torch.manual_seed(seed)
    m = pcseg_b200.PointNetSegmentation(C).to(dev).train()
    m.dropout.p = 0.0
    return m


m = model(3 + 17 * rank)                         # deliberately different initialisation on every rank
tr = pcseg_b200.FusedTrainer(m, class_weights=cw, use_cuda_graph=False)
p0 = tr.flat["params"].clone()
dist.broadcast(p0, src=0)
same_init = bool(torch.equal(p0, tr.flat["params"]))
seeds = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
dist.all_gather(seeds, tr.state.view(torch.int64)[0:1].clone())
distinct_seeds = len({int(s.item()) for s in seeds}) == world
params_before = tr.flat["params"].clone()
out = tr.step(x[rank * per:(rank + 1) * per].contiguous(), y[rank * per:(rank + 1) * per].contiguous())
wsum_dp = tr.global_wsum.item()
g_dp = tr.flat["grads"].clone() / (wsum_dp if tr.deferred else 1.0)      # deferred: the arena holds the un-normalised sum
loss_dp = out["loss"].item()
p_after = tr.flat["params"].clone()
pa = p_after.clone()
dist.broadcast(pa, src=0)
same_after = bool(torch.equal(pa, p_after))
ok, res = True, {}
if rank == 0:
    # emulation: each replica's step WITHOUT all-reduce but with the global normaliser, gradients added
    wsum = cw[y.clamp(min=0)].mul(y >= 0).sum().double().reshape(1)
    total, num = None, 0.0
    for r in range(world):
        mr = model(3)
        eng = mr._get_engine(dev)
        f = mr._ensure_flat(dev)
        ce = torch.zeros(32, dtype=torch.uint8, device=dev)
        xs, ys = x[r * per:(r + 1) * per].contiguous(), y[r * per:(r + 1) * per].contiguous()
        logits = eng.forward_train(xs, f["params"], f["bn"], 0, 0.0, ys, cw, ce)
        eng.backward(xs, f["params"], f["grads"], logits=logits, labels=ys, class_w=cw, wsum=wsum)
        total = f["grads"].clone() if total is None else total + f["grads"]
        num += ce.view(torch.float64)[0].item()
    cos = torch.nn.functional.cosine_similarity(g_dp.double(), total.double(), dim=0).item()
    rel = ((g_dp - total).norm() / total.norm()).item()
    # Adam on the emulated gradient from the same start: the optimizer step of the data-parallel ranks
    ref = model(3)
    fr = ref._ensure_flat(dev)
    eng = ref._get_engine(dev)
    ma, va = torch.zeros_like(total), torch.zeros_like(total)
    eng.adam(fr["params"], total, ma, va, 1, 1e-3, (0.9, 0.999), 1e-8, 1e-4)
    dp_step = (p_after - params_before)
    ref_step = (fr["params"] - params_before)
    step_cos = torch.nn.functional.cosine_similarity(dp_step.double(), ref_step.double(), dim=0).item()
    res = dict(world=world, grad_cosine=cos, grad_rel_diff=rel, loss_dp=loss_dp, loss_emulated=num / wsum.item(), wsum_dp=wsum_dp,
               wsum=wsum.item(), same_init=same_init, distinct_dropout_seeds=distinct_seeds, params_identical_after_step=same_after,
               adam_step_cosine=step_cos, deferred=bool(tr.deferred), comm=tr.comm)
    ok = (cos > 0.9995 and rel < 3e-2 and abs(loss_dp - num / wsum.item()) < 1e-5 * abs(loss_dp) and abs(wsum_dp - wsum.item()) < 1e-6 * wsum.item()
          and same_init and distinct_seeds and same_after and step_cos > 0.95)      # (first Adam step ~ lr * sign(g): near-zero gradient elements flip)
    res["ok"] = bool(ok)
    print(json.dumps(res))
flag = torch.tensor([1 if (ok and same_init and same_after) else 0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if flag.item() == 1 else 1)
