#!/usr/bin/env python
"""Benchmark of the point-cloud segmentation hot path (BASELINE.json metric: segmentation points/sec,
fwd+bwd training step).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--no-cpu-baseline]
                    [--workload cfg2|cfg3_train|cfg4|cfg1_eval|cfg2_eval|cfg3_eval|cfg5_eval|cfg5_eval_sharded]

One "step" = one full training step (forward with batch statistics and dropout p=0.3, weighted
cross-entropy, backward of all 38 parameter tensors, NCCL gradient all-reduce when N > 1, Adam) on one
batch of synthetic clouds.  N = 1 workload: BASELINE.json configs[1] = batch 8 x 16 384 points, C = 5.
N > 1: every rank runs that batch on its own clouds (weak scaling, data parallel).
Prints ONE JSON line (rank 0): value = device-timed throughput with resident inputs, e2e = the same through the public API
with pinned-host inputs copied every step and the result read back, roofline = the dominant tcgen05 GEMM against the
measured bf16 peak, cpu_baseline = the torch-CPU port of the reference step on the host cores, torch_eager_same_gpu = the
reference network with stock torch layers on this GPU (TF32 and IEEE fp32).  The *_eval workloads time inference,
cfg5_eval_sharded splits one 1M-point scene over the ranks (strong scaling).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

WORKLOADS = {
    # name: (clouds per rank, points per cloud, mode)
    "cfg2": (8, 16384, "train"),
    "cfg3_train": (16, 131072, "train"),
    "cfg1_eval": (1, 16384, "eval"),      # configs[0]: one 16k-point cloud, inference (the reference's CPU-runnable case)
    "cfg2_eval": (8, 16384, "eval"),
    "cfg3_eval": (16, 131072, "eval"),
    "cfg4": (8, 65536, "train"),          # configs[3] at 8 GPUs: global 64 x 64k -> 8 clouds x 65 536 per GPU
    # configs[3] as quoted: GLOBAL batch 64 x 65 536 split over the ranks (32 / 16 / 8 clouds per GPU at 2 / 4 / 8 GPUs; all
    # 64 on one GPU): strong scaling, gradients all-reduced over NCCL
    "cfg4_strong": (64, 65536, "train_strong"),
    "cfg5_eval": (1, 1048576, "eval"),    # configs[4]: one 1M-point scene, inference
    # configs[4] with the scene's POINTS split over the ranks: strong scaling, one MAX all-reduce of the
    # 1024-float pooled feature per step; at --gpus 1 it is cfg5_eval through the two-part entry point
    "cfg5_eval_sharded": (1, 1048576, "eval_sharded"),
}
NUM_CLASSES = 5
DROPOUT_P = 0.3
# algorithmic FLOPs per point: fwd 2*(1 392 896 + 128 C); fwd+bwd = 3x fwd - 512
FWD_FLOP_PER_PT = 2 * (1392896 + 128 * NUM_CLASSES)
TRAIN_FLOP_PER_PT = 3 * FWD_FLOP_PER_PT - 512
GFEAT_FLOP_PER_PT = 2 * 1024 * 1024          # one global_feat GEMM (forward, dgrad or wgrad), per point
# dram__bytes_read.sum + dram__bytes_write.sum per launch (MB) of the global_feat GEMMs at 8 x 16 384 points, from the
# committed `ncu --set full` capture named in NCU_TRAFFIC_SOURCE (refreshed whenever those kernels change).  Tags: 5 forward
# (statistics + max-pool epilogue, nothing stored), 21 data gradient (a5 S + side rows, masked by a5), 53 Gram matrix a5^T a5
# (upper-triangle tiles), 69 inference forward (max-pool epilogue).  Algorithmic minimum: 268 MB per read or written
# 1024-channel bf16 tensor.
NCU_TRAFFIC_SOURCE = "profiles/r02_ncu_full_gemm_cfg2_train.txt"
NCU_TRAFFIC_MB_CFG2 = {}


def load_ncu_traffic():
    """{tag: MB} parsed from the committed capture summary (lines '# traffic tag=<t> MB=<x>'); empty when absent."""
    path = os.path.join(ROOT, NCU_TRAFFIC_SOURCE)
    out = {}
    if os.path.exists(path):
        for ln in open(path):
            if ln.startswith("# traffic tag="):
                try:
                    t, mb = ln.split()[2:4]
                    out[int(t.split("=")[1])] = float(mb.split("=")[1])
                except (ValueError, IndexError):
                    pass
    return out


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return dict(hbm=d["hbm_gbs"], bf16_burst=d["bf16_tflops"], bf16_sustained=d["bf16_tflops_sustained"], source="measured")
    return dict(hbm=6650.0, bf16_burst=1590.0, bf16_sustained=1400.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def mark(self):
        """index of the next sample (used to cut the samples that fall inside the timed region)"""
        return len(self.lines)

    def stop(self, lo=0, hi=None):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        hi = len(self.lines) if hi is None else hi
        lines = self.lines[lo:hi] if hi > lo else self.lines[max(0, lo - 2):hi + 2]
        for ln in lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for n, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def synth_batch(B, N, C, seed):
    rng = np.random.default_rng(seed)
    x = rng.random((B, N, 4), dtype=np.float32)               # xyz,e ~ U[0,1)
    labels = rng.integers(0, C, (B, N)).astype(np.int64)
    return x, labels


# ------------------------------------------------------------------------------------------------
# CPU leg: the reference's own implementation on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_step_fn(mode, C):
    """The UNMODIFIED reference module (baseline/_ref/point_cloud_segmentation.py, staged by build(); kind "reference")
    running its own step (pcs.py:241-258 / 450-452) on all host threads; the torch-CPU port of oracle/ (kind "port") only
    if that copy is missing."""
    import torch
    from oracle import reference_module as refmod
    if refmod.available():
        runner, kind = refmod.ReferenceCpuStep(C, seed=1234, threads=os.cpu_count()), "reference"
    else:
        from oracle.torch_port import TorchCpuPort
        runner, kind = TorchCpuPort(C, seed=1234, threads=os.cpu_count()), "port"
    cw = torch.ones(C)

    def train_step(x, labels):
        return runner.train_step(torch.from_numpy(x), torch.from_numpy(labels), cw, DROPOUT_P)

    def eval_step(x, labels):
        return runner.eval_step(torch.from_numpy(x))

    return (train_step if mode == "train" else eval_step), kind


def time_cpu(mode, B, N, C, steps, warmup, budget_s):
    """Times the reference step on the stated B x N batch when (steps + warmup) of it fit the budget, otherwise on a bounded
    sample (fewer whole clouds, or one shorter cloud).  Returns a dict: value (points/s), ms, b, n, same_config, kind, sample."""
    step, kind = cpu_step_fn(mode, C)
    xs, ls = synth_batch(1, 2048, C, 99)
    for _ in range(2):                               # warm the probe: the first calls pay thread-pool / allocator start-up
        step(xs, ls)
    t0 = time.perf_counter()
    step(xs, ls)
    step(xs, ls)
    per_pt = (time.perf_counter() - t0) / 2 / 2048
    total_steps = steps + warmup
    pts_budget = max(2048, int(budget_s / max(per_pt, 1e-9) / max(total_steps, 1)))
    if pts_budget >= B * N:
        b, n = B, N
    elif pts_budget >= N:
        b, n = max(1, pts_budget // N), N
    else:
        b, n = 1, max(2048, (pts_budget // 1024) * 1024)
    x, lab = synth_batch(b, n, C, 5)
    for _ in range(warmup):
        step(x, lab)
    t0 = time.perf_counter()
    for _ in range(steps):
        step(x, lab)
    dt = (time.perf_counter() - t0) / max(steps, 1)
    what = ("the unmodified reference module (baseline/_ref), CPU fp32" if kind == "reference"
            else "torch-CPU fp32 port of the reference step (oracle/torch_port.py)")
    sample = f"{b} cloud(s) x {n} points per step ({steps} timed steps, {what}, all host threads, {mode})"
    return dict(value=b * n / dt, ms=dt * 1e3, b=b, n=n, same_config=(b == B and n == N), kind=kind, sample=sample)


def run_reference(args, B, N, mode):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    steps = max(1, args.steps)
    warmup = max(0, min(args.warmup, 3))
    # the stated batch when the whole run fits about two minutes of CPU work, otherwise a bounded sample of it, SAID in
    # config.workload (PCSEG_REF_BUDGET_S overrides the budget: the contract test uses a few seconds)
    r = time_cpu(mode, B, N, NUM_CLASSES, steps, warmup, budget_s=float(os.environ.get("PCSEG_REF_BUDGET_S", "150")))
    workload = workload_desc(args.workload, B, N, mode)
    if not r["same_config"]:
        workload += f" -- CPU arm timed on a bounded sample: {r['b']} x {r['n']} points"
    line = {
        "impl": "reference", "metric": metric_name(mode), "value": r["value"], "unit": "points/s", "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": r["ms"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": {"workload": workload, "num_classes": NUM_CLASSES, "same_config": r["same_config"],
                                        "sample_batch": [r["b"], r["n"]]},
        "cpu_baseline": {"value": r["value"], "unit": "points/s", "cores": cores, "kind": r["kind"], "sample": r["sample"]},
        "e2e": {"value": r["value"], "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def metric_name(mode):
    return "segmentation points/sec (fwd+bwd train step)" if mode == "train" else "segmentation points/sec (fwd inference)"


def workload_desc(name, B, N, mode):
    return f"{name}: {mode} step, batch {B} x {N} points per GPU, C={NUM_CLASSES}, dropout {DROPOUT_P if mode == 'train' else 0}"


# ------------------------------------------------------------------------------------------------
# GPU leg
# ------------------------------------------------------------------------------------------------
def measure_inference(model, eng, x_dev, x_host, B, N, steps, barrier, dev, world):
    """Inference (pcs.py:448-452) on the workload's batch: device-timed with resident inputs, end to end through
    PredictStream (pinned host points in, argmax labels out, every step), and the dominant kernel's roofline."""
    import torch
    import torch.distributed as dist
    import pcseg_b200
    from pcseg_b200.engine import profile_enable, profile_read
    model.eval()
    with torch.no_grad():
        for _ in range(3):
            model(x_dev)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0 = pcseg_b200.launch_count()
        e0.record()
        for _ in range(steps):
            model(x_dev)
        e1.record()
        barrier()
        launches = pcseg_b200.launch_count() - c0
        ms = e0.elapsed_time(e1) / steps
        profile_enable(eng, B, N, True, train=False)
        for _ in range(steps):
            model(x_dev)
        torch.cuda.synchronize()
        prof = profile_read(eng, B, N, train=False)
        profile_enable(eng, B, N, False, train=False)
    ps = pcseg_b200.PredictStream(model)
    prev = None
    for _ in range(3):
        t = ps.submit(x_host)
        if prev is not None:
            ps.result(prev)
        prev = t
    barrier()
    e2e_steps = max(steps, 50)
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        t = ps.submit(x_host)
        ps.result(prev)
        prev = t
    ps.result(prev)
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps
    tt = torch.tensor([ms, e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms, e2e_ms = float(tt[0].item()), float(tt[1].item())
    peaks = measured_peaks()
    pts = B * N * world
    out = {"metric": metric_name("eval"), "value": pts / (ms * 1e-3), "unit": "points/s", "ms_per_step": ms, "steps": steps,
           "gpu_launches": launches,
           "e2e": {"value": pts / (e2e_ms * 1e-3), "unit": "points/s", "h2d_bytes_per_step": x_host.numel() * 4,
                   "d2h_bytes_per_step": B * N * 8, "ms_per_step": e2e_ms, "timed_steps": e2e_steps},
           "step_tflops_per_gpu": FWD_FLOP_PER_PT * B * N / (ms * 1e-3) / 1e12}
    out["step_frac_of_bf16_burst"] = out["step_tflops_per_gpu"] / peaks["bf16_burst"]
    if 69 in prof:
        kms = prof[69][0] / prof[69][1]
        achieved = GFEAT_FLOP_PER_PT * B * N / (kms * 1e-3) / 1e12
        traffic = load_ncu_traffic()
        out["roofline"] = {"bound": "tensor", "kernel": "global_feat inference GEMM (1024x1024, bias + ReLU + max-pool epilogue, nothing stored)",
                           "achieved": achieved, "peak": peaks["bf16_burst"], "unit": "TFLOP/s", "frac": achieved / peaks["bf16_burst"],
                           "frac_of_sustained": achieved / peaks["bf16_sustained"], "ms_per_launch": kms,
                           "traffic": (traffic[69] * 1e6 if (B, N) == (8, 16384) and 69 in traffic else None),
                           "traffic_source": NCU_TRAFFIC_SOURCE, "share_of_step": kms / ms}
    return out


def run_ours(args, B, N, mode):
    import torch
    import torch.distributed as dist

    import pcseg_b200
    from pcseg_b200.engine import profile_enable, profile_read

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"       # keep stdout to the one JSON line (NCCL prints its version banner there)
        dist.init_process_group("nccl", device_id=dev)
    C = NUM_CLASSES
    torch.manual_seed(1234)                        # same random-init weights on every rank
    model = pcseg_b200.PointNetSegmentation(C).to(dev)
    strong = mode == "train_strong"
    if strong:
        if B % world:
            raise SystemExit(f"cfg4_strong: {B} clouds do not split over {world} ranks")
        B //= world                                # DataParallel's dim-0 chunking of the global batch (pcs.py:211)
        mode = "train"
    sharded = mode == "eval_sharded"
    if sharded:
        N_total = N
        N = N // world                             # this rank's slice of every cloud
        mode = "eval"
    x_np, lab_np = synth_batch(B, N, C, 100 + rank)
    x_host = torch.from_numpy(x_np).pin_memory()
    lab_host = torch.from_numpy(lab_np).pin_memory()
    x_dev = x_host.to(dev)
    lab_dev = lab_host.to(dev)
    cw = torch.ones(C, device=dev)

    if mode == "train":
        model.train()
        trainer = pcseg_b200.FusedTrainer(model, class_weights=cw, lr=1e-3, weight_decay=1e-4, device=dev,
                                          overlap=os.environ.get("PCSEG_DDP_OVERLAP", "1") != "0")

        def step_resident():
            return trainer.step(x_dev, lab_dev)["loss"]

        e2e_ticket = [None]

        def step_e2e():
            # every step: pinned host -> device copy of ITS inputs (double buffered on a copy stream, so the copy of step
            # i+1 overlaps step i), the step, and a device -> host read of the loss (pcs.py:237-238, 258)
            if e2e_ticket[0] is None:
                e2e_ticket[0] = trainer.prefetch(x_host, lab_host)
            cur = e2e_ticket[0]
            e2e_ticket[0] = trainer.prefetch(x_host, lab_host)
            return float(trainer.step_prefetched(cur)["loss"].item())
        h2d = x_host.numel() * 4 + lab_host.numel() * 8
        d2h = 8
        flop_per_pt = TRAIN_FLOP_PER_PT
    else:
        model.eval()
        model.set_precision(args.precision)

        def step_resident():
            with torch.no_grad():
                return model.predict_point_sharded(x_dev)[0] if sharded else model(x_dev)

        out_host = torch.empty((B, N), dtype=torch.int64).pin_memory()
        pstream = None if sharded else pcseg_b200.PredictStream(model)
        pending = [None]

        def step_e2e():
            # every step: this batch's points from pinned host memory, per-point predicted labels back on the host
            # (pcs.py:446-454).  The copies are double buffered on side streams (PredictStream): a step submits batch i
            # and consumes the labels of batch i-1.
            if sharded:
                xd = x_host.to(dev, non_blocking=True)
                with torch.no_grad():
                    _, labels = model.predict_point_sharded(xd)
                out_host.copy_(labels, non_blocking=True)
                torch.cuda.synchronize()
                return out_host
            t = pstream.submit(x_host)
            prev, pending[0] = pending[0], t
            return pstream.result(prev) if prev is not None else None
        h2d = x_host.numel() * 4
        d2h = B * N * 8
        flop_per_pt = FWD_FLOP_PER_PT

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    warmup = max(3, args.warmup)
    steps = max(1, args.steps)
    sampler = ClockSampler(local_rank)
    sampler.start()                                 # started early: nvidia-smi needs ~100s of ms before its first sample
    for _ in range(warmup):
        step_resident()
    barrier()
    # A fresh box needs more than W short steps before a ~2 ms step is steady (measured: the first 20-step region after 5
    # warm-up steps on an idle GPU ran 5 % slow, 1.984 vs 1.879 ms, with SM clocks already reported at max): keep stepping,
    # untimed, until the GPU has been busy for SETTLE_MS in total.  Reported in config.extra_warmup_steps.
    SETTLE_MS = 250.0
    t_settle = time.perf_counter()
    for _ in range(5):
        step_resident()
    torch.cuda.synchronize()
    per_step_ms = torch.tensor([(time.perf_counter() - t_settle) * 1e3 / 5], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(per_step_ms, op=dist.ReduceOp.MAX)      # every rank must run the SAME number of steps (collectives inside)
    extra_warmup = 5 + int(min(200, max(0.0, SETTLE_MS / max(float(per_step_ms.item()), 1e-3) - 5)))
    for _ in range(extra_warmup - 5):
        step_resident()
    barrier()

    eng = model._get_engine(dev)
    launches0 = pcseg_b200.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    mark0 = sampler.mark()
    e0.record()
    for _ in range(steps):
        step_resident()
    e1.record()
    barrier()
    mark1 = sampler.mark()
    launches = pcseg_b200.launch_count() - launches0
    ms_total = e0.elapsed_time(e1)
    graph_replay = mode == "train" and trainer._graph is not None
    if graph_replay:
        # a replayed CUDA graph re-launches the kernels recorded at capture time without passing through the
        # library's host-side counter: count the kernels of one eager step instead
        trainer.profiling = True
        c0 = pcseg_b200.launch_count()
        step_resident()
        torch.cuda.synchronize()
        launches = (pcseg_b200.launch_count() - c0) * steps
        trainer.profiling = False

    # second pass of the same K steps with CUDA events around every tcgen05 GEMM launch (eager launches) -> roofline
    prof = {}
    ms_prof_total = None
    if mode == "train":
        trainer.profiling = True
        profile_enable(eng, B, N, True)
        barrier()
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0.record()
        for _ in range(steps):
            step_resident()
        p1.record()
        barrier()
        ms_prof_total = p0.elapsed_time(p1)
        prof = profile_read(eng, B, N)
        profile_enable(eng, B, N, False)
        trainer.profiling = False
        if getattr(trainer, "peer_events", None):
            torch.cuda.synchronize()
            ms = [a.elapsed_time(b) for a, b in trainer.peer_events]
            prof[94] = (sum(ms), len(ms))
            trainer.peer_events = []

    # the other half of the metric ("points/sec (fwd, fwd+bwd)"): inference on the same batch, same model, same process
    fwd = None
    if mode == "train" and not args.no_fwd:
        fwd = measure_inference(model, eng, x_dev, x_host, B, N, steps, barrier, dev, world)
        model.train()

    # end-to-end: pinned host inputs copied every step, result read back every step
    for _ in range(3):
        step_e2e()
    barrier()
    # (host-clocked: short steps are repeated at least 50 times so that the region is not a few milliseconds long)
    e2e_steps = steps if ms_total / steps >= 5.0 else max(steps, 50)
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_e2e()
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) * steps / e2e_steps          # normalised to K steps
    clocks = sampler.stop(mark0, mark1)             # samples taken during the device-timed region (neighbours if it was < 100 ms)

    t = torch.tensor([ms_total, e2e_s * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, e2e_ms_total = float(t[0].item()), float(t[1].item())
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    pts_per_step = B * N * world                    # (sharded: N is the per-rank slice, so this is the whole scene)
    ms_per_step = ms_total / steps
    value = pts_per_step / (ms_per_step * 1e-3)
    e2e_value = pts_per_step / (e2e_ms_total / steps * 1e-3)
    peaks = measured_peaks()

    roof = None
    kernels = {}
    if prof:
        # the 1024 x 1024 layer: forward, data gradient, and (folded BatchNorm backward) the Gram matrix of its input that
        # replaces the weight-gradient GEMM; legacy step (PCSEG_FOLDED=0): tag 37 is the weight gradient
        names = {5: "global_feat fwd GEMM (1024x1024, BN statistics + max-pool epilogue)",
                 21: "global_feat data-gradient GEMM (mask + column-sum epilogue)",
                 37: "global_feat weight-gradient GEMM (MN-major, split-K)",
                 53: "Gram matrix a5^T a5 (MN-major, split-K, upper-triangle tiles = 62.5 % of 2*1024*1024 FLOP/point)"}
        flop = {5: GFEAT_FLOP_PER_PT, 21: GFEAT_FLOP_PER_PT, 37: GFEAT_FLOP_PER_PT, 53: GFEAT_FLOP_PER_PT * 20 // 32}
        from pcseg_b200.engine import KERNEL_TAGS
        for tag, (ms, n) in prof.items():
            kernels[str(tag)] = {"ms_per_launch": ms / n, "launches": n}
            if tag in KERNEL_TAGS:
                kernels[str(tag)]["kernel"] = KERNEL_TAGS[tag]
        tag = max((tg for tg in (5, 21, 37, 53) if tg in prof), key=lambda tg: prof[tg][0] / prof[tg][1])
        ms, n = prof[tag]
        achieved = flop[tag] * B * N / (ms / n * 1e-3) / 1e12
        share = sum(prof.get(tg, (0, 1))[0] for tg in (5, 21, 37, 53)) / ms_prof_total
        traffic = load_ncu_traffic()
        roof = {"bound": "tensor", "kernel": names[tag], "achieved": achieved, "peak": peaks["bf16_burst"], "unit": "TFLOP/s",
                "frac": achieved / peaks["bf16_burst"], "frac_of_sustained": achieved / peaks["bf16_sustained"],
                "peak_sustained": peaks["bf16_sustained"],
                "traffic": (traffic[tag] * 1e6 if (B, N) == (8, 16384) and tag in traffic else None),
                "traffic_unit": "bytes per launch (ncu dram read+write)", "traffic_source": NCU_TRAFFIC_SOURCE,
                "peak_source": peaks["source"] + " bf16 burst (kernel event-timed on its own inside a short region at boost clock)",
                "ms_per_launch": ms / n, "global_feat_gemms_share_of_step": share,
                "measured": "CUDA events around each GEMM launch during a second pass of the same K steps (eager launches); "
                            "the headline value is from the first pass (CUDA-graph replay, no per-kernel events)"}
    step_tflops = flop_per_pt * B * N / (ms_per_step * 1e-3) / 1e12        # per GPU

    cores = os.cpu_count() or 1
    if args.gpus == 1 and not args.no_cpu_baseline:
        r = time_cpu(mode, B, N, C, steps=2, warmup=1, budget_s=12.0)
        cpu = {"value": r["value"], "unit": "points/s", "cores": cores, "kind": r["kind"], "sample": r["sample"]}
        # the reference network through stock torch eager (cuDNN / cuBLAS fp32) on this same B200,
        # on a bounded sample of the workload (whole clouds; the reference keeps ~37 KB of saved tensors per point)
        from oracle.torch_port import time_stock_torch_on_gpu
        tb = max(1, min(B, 262144 // N)) if N <= 262144 else 1
        tn = min(N, 262144)
        te = time_stock_torch_on_gpu(mode, tb, tn, C, steps=10, warmup=3, device=dev)
        torch_eager = {"tf32": te["tf32"], "ieee_fp32": te["ieee"], "unit": "points/s", "kind": "port",
                       "sample": f"{tb} cloud(s) x {tn} points per step, 10 timed steps, stock nn.Conv1d/BatchNorm1d eager in the "
                                 f"reference's channel-major layout, {mode}"}
    else:
        cpu = None
        torch_eager = None

    line = {
        "metric": metric_name(mode), "value": value, "unit": "points/s", "n_gpus": world, "steps": steps, "warmup": warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if (sharded or strong) else "weak", "vs_baseline": None,
        "dtype": args.precision if mode == "eval" else "bf16",
        "data": "synthetic",
        "config": {"workload": workload_desc(args.workload, B, N, mode), "num_classes": C, "l2_policy": "working set (GBs of activations) >> 126 MB L2, no flush needed",
                   "optimizer": "Adam lr 1e-3 wd 1e-4 (inside the timed step)" if mode == "train" else None,
                   "cuda_graph": bool(graph_replay) if mode == "train" else False,
                   "extra_warmup_steps": extra_warmup,
                   "global_batch": [B * world, N] if not sharded else [B, N * world],
                   "parallelism": (f"points of each cloud sharded over {world} ranks, MAX all-reduce of the pooled feature" if sharded
                                   else (f"dp{world}" if world > 1 else "single"))},
        "e2e": {"value": e2e_value, "unit": "points/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_ms_total / steps, "timed_steps": e2e_steps},
        "gpu_launches": launches, "clocks": clocks, "roofline": roof, "cpu_baseline": cpu, "torch_eager_same_gpu": torch_eager,
        "step_tflops_per_gpu": step_tflops, "step_frac_of_bf16_burst": step_tflops / peaks["bf16_burst"],
        "step_frac_of_bf16_sustained": step_tflops / peaks["bf16_sustained"],
        "gemm_kernels": kernels,
    }
    if fwd is not None:
        line["fwd"] = fwd
        line["metric"] = "segmentation points/sec (fwd+bwd train step; fwd inference under key 'fwd')"
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-fwd", action="store_true", help="train workloads: skip the inference half of the metric")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "bf16x3"],
                    help="inference arithmetic of the *_eval workloads: bf16x3 = split-bf16, fp32-grade logits (DESIGN.md §3.6)")
    args = ap.parse_args()
    B, N, mode = WORKLOADS[args.workload]
    if args.impl == "reference":
        mode = "train" if mode == "train_strong" else ("eval" if mode == "eval_sharded" else mode)
        run_reference(args, B, N, mode)
    else:
        run_ours(args, B, N, mode)


if __name__ == "__main__":
    main()
