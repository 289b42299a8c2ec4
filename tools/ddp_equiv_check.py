"""torchrun --nproc-per-node 2 tools/ddp_equiv_check.py: two data-parallel ranks (4 clouds each) against ONE process that
emulates nn.DataParallel on rank 0 (per-replica BN statistics, one global weighted-mean loss, summed gradients)."""
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
C, B, N = 5, 8, 2048
g = torch.Generator().manual_seed(1)
x = torch.rand(B, N, 4, generator=g).to(dev)
y = torch.randint(-1, C, (B, N), generator=g).to(dev)
cw = torch.tensor([0.5, 1.0, 2.0, 0.75, 1.5], device=dev)
per = B // world


def model():
    torch.manual_seed(3)
    m = pcseg_b200.PointNetSegmentation(C).to(dev).train()
    m.dropout.p = 0.0
    return m


m = model()
tr = pcseg_b200.FusedTrainer(m, class_weights=cw, use_cuda_graph=False)
out = tr.step(x[rank * per:(rank + 1) * per].contiguous(), y[rank * per:(rank + 1) * per].contiguous())
g_dp = tr.flat["grads"].clone()
loss_dp = out["loss"].item()
if rank == 0:
    # emulation: each replica's step WITHOUT all-reduce but with the global normaliser, gradients added
    wsum = cw[y.clamp(min=0)].mul(y >= 0).sum().double().reshape(1)
    total, num = None, 0.0
    for r in range(world):
        mr = model()
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
    print(f"DDP vs emulated DataParallel: grad cosine {cos:.6f}, rel diff {rel:.2e}, loss {loss_dp:.6f} vs {num / wsum.item():.6f}")
dist.barrier()
dist.destroy_process_group()
