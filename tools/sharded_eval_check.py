"""torchrun --nproc-per-node G tools/sharded_eval_check.py: point-sharded inference of one scene over G GPUs (NCCL MAX
all-reduce of the pooled feature) against the un-sharded forward on rank 0."""
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
torch.manual_seed(7)
m = pcseg_b200.PointNetSegmentation(5).to(dev).eval()
N = 1 << 20
g = torch.Generator(device="cpu").manual_seed(3)
x = torch.rand(1, N, 4, generator=g)
per = N // world
xs = x[:, rank * per:(rank + 1) * per].contiguous().to(dev)
with torch.no_grad():
    logits, labels = m.predict_point_sharded(xs)
    ok = True
    if rank == 0:
        full, full_lab = m.predict(x.to(dev))
        ok = bool(torch.equal(full[:, :per], logits) and torch.equal(full_lab[:, :per], labels))
    for _ in range(5):
        m.predict_point_sharded(xs)
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        m.predict_point_sharded(xs)
    e1.record()
    torch.cuda.synchronize()
t = torch.tensor([e0.elapsed_time(e1) / 20], device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print(f"sharded over {world} GPUs: bit-identical to the un-sharded forward on rank 0's slice: {ok}; {t.item():.3f} ms per scene "
          f"= {N / t.item() / 1e3:.1f} M points/s")
flag = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
dist.destroy_process_group()
sys.exit(0 if flag.item() == 1 else 1)
