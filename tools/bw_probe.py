import torch, time
dev = "cuda"
n = 1 << 30  # 1 GiB of bytes
x = torch.empty(n, dtype=torch.uint8, device=dev)
y = torch.empty(n, dtype=torch.uint8, device=dev)
def timeit(f, reps=10):
    f(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
t = timeit(lambda: x.zero_()); print(f"write-only (memset 1 GiB): {n/t/1e6:.0f} GB/s")
xf = x.view(torch.float32)
t = timeit(lambda: xf.sum()); print(f"read-only (sum 1 GiB fp32): {n/t/1e6:.0f} GB/s")
t = timeit(lambda: y.copy_(x)); print(f"copy 1 GiB: {2*n/t/1e6:.0f} GB/s (read+write)")
xb = x.view(torch.bfloat16)[: n // 4]; yb = y.view(torch.bfloat16)[: n // 4]
t = timeit(lambda: torch.relu(xb, out=yb) if False else yb.copy_(xb)); print(f"copy 256 MiB: {2*(n//2)/t/1e6:.0f} GB/s")
