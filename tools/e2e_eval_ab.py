"""A/B of the host-fed inference loop: synchronous round trip vs PredictStream (copies on side streams)"""
import os, sys, time
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import pcseg_b200
B, N, C = 8, 16384, 5
dev = torch.device("cuda", 0)
m = pcseg_b200.PointNetSegmentation(C).to(dev).eval()
x_host = torch.rand(B, N, 4).pin_memory()
out_host = torch.empty((B, N), dtype=torch.int64).pin_memory()
def sync_step():
    xd = x_host.to(dev, non_blocking=True)
    with torch.no_grad():
        _, labels = m.predict(xd)
    out_host.copy_(labels, non_blocking=True)
    torch.cuda.synchronize()
ps = pcseg_b200.PredictStream(m)
pend = [None]
def pipe_step():
    t = ps.submit(x_host)
    prev, pend[0] = pend[0], t
    if prev is not None:
        ps.result(prev)
def resident_step():
    with torch.no_grad():
        m.predict(xd0)
xd0 = x_host.to(dev)
for name, fn in (("resident (no copies)", resident_step), ("sync round trip", sync_step), ("PredictStream", pipe_step), ("sync round trip", sync_step), ("PredictStream", pipe_step)):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(100):
        fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 100
    print(f"{name:22s}: {dt*1e3:.3f} ms/step = {B*N/dt/1e6:.1f} M points/s")
# host-only cost of a submit
t0 = time.perf_counter()
for _ in range(100):
    with torch.no_grad():
        m.predict(xd0)
t_host = (time.perf_counter() - t0) / 100
torch.cuda.synchronize()
print(f"host time of predict() enqueue: {t_host*1e3:.3f} ms")
