"""Multi-threaded CPU port of the reference training / inference step, used ONLY as the timed CPU
baseline of bench.py (`cpu_baseline`, `--impl reference`) and checked against the golden vectors in
tests/test_oracle_golden.py.  TEST / BENCH INFRASTRUCTURE ONLY — the product never imports it.

It restates point_cloud_segmentation.py:98-133 (forward), :216/:251 (weighted CE), :254 (autograd
backward) and :217/:255 (Adam) with stock torch CPU operators on point-major (P, C) matrices, so it runs
on all host cores through the same BLAS/oneDNN kernels the reference itself would use on this box.
"""
import math

import torch
import torch.nn.functional as F

from . import pointnet_oracle as orc


class TorchCpuPort:
    def __init__(self, num_classes, state=None, seed=1234, threads=None):
        if threads:
            torch.set_num_threads(threads)
        self.C = num_classes
        sd = state if state is not None else orc.synth_state(num_classes, seed)
        self.p = {k: torch.tensor(v).clone() for k, v in sd.items()}
        self.params = [k for k in self.p if self.p[k].dtype == torch.float32 and "running" not in k]
        for k in self.params:
            self.p[k].requires_grad_(True)
        self.opt = torch.optim.Adam([self.p[k] for k in self.params], lr=1e-3, weight_decay=1e-4)   # pcs.py:217

    def _block(self, a, conv, bn, training, w_cols=None):
        W = self.p[f"{conv}.weight"][:, :, 0]
        if w_cols is not None:
            W = W[:, w_cols]
        y = F.linear(a, W, self.p[f"{conv}.bias"])                                                 # Conv1d(k=1), pcs.py:106-127
        y = F.batch_norm(y, self.p[f"{bn}.running_mean"], self.p[f"{bn}.running_var"], self.p[f"{bn}.weight"],
                         self.p[f"{bn}.bias"], training, 0.1, 1e-5)
        if training:
            self.p[f"{bn}.num_batches_tracked"] += 1
        return F.relu(y)

    def forward(self, x, training, dropout_p=0.3):
        B, N, _ = x.shape
        a = x.reshape(B * N, -1)
        for conv, bn, _, _ in orc.TRUNK:
            a = self._block(a, conv, bn, training)
            if conv == "conv2":
                pf = a                                                                              # pcs.py:107
        g = a.view(B, N, -1).max(dim=1)[0]                                                          # pcs.py:114
        cat = torch.cat([pf, g.repeat_interleave(N, dim=0)], dim=1)                                 # pcs.py:117-120
        a = self._block(cat, "seg_conv1", "bn_seg1", training)
        a = F.dropout(a, dropout_p, training)                                                       # pcs.py:124
        a = self._block(a, "seg_conv2", "bn_seg2", training)
        a = F.dropout(a, dropout_p, training)                                                       # pcs.py:126
        a = self._block(a, "seg_conv3", "bn_seg3", training)
        z = F.linear(a, self.p["seg_conv4.weight"][:, :, 0], self.p["seg_conv4.bias"])              # pcs.py:128
        return z.view(B, N, self.C)

    def train_step(self, x, labels, class_w=None, dropout_p=0.3):
        """optimizer.zero_grad / forward / weighted CE / backward / optimizer.step, pcs.py:241-255."""
        self.opt.zero_grad()
        logits = self.forward(x, True, dropout_p)
        loss = F.cross_entropy(logits.view(-1, self.C), labels.view(-1), weight=class_w, ignore_index=-1)
        loss.backward()
        self.opt.step()
        return loss.item()

    @torch.no_grad()
    def eval_step(self, x):
        return self.forward(x, False).argmax(dim=2)                                                 # pcs.py:450-452


class StockTorchModel(torch.nn.Module):
    """The reference network written with the stock torch layers the reference itself uses, in ITS layout: channel-major
    (B, C, N) activations through nn.Conv1d(k=1) / nn.BatchNorm1d / F.relu / torch.max / repeat / cat / nn.Dropout
    (pcs.py:70-96 layers, :98-133 forward).  bench.py times it with torch eager on the GPU box's B200 ("the unmodified
    reference model on the same B200 through stock torch eager") next to the CUDA path; pinned to
    the reference's golden vectors by tests/test_oracle_golden.py.  TEST / BENCH INFRASTRUCTURE ONLY."""

    def __init__(self, num_classes, state=None, seed=1234):
        super().__init__()
        nn = torch.nn
        dims = {c: (ci, co) for c, _, ci, co in orc.TRUNK + orc.HEAD}
        dims["seg_conv4"] = (128, num_classes)
        for c in orc.CONV_NAMES:
            setattr(self, c, nn.Conv1d(dims[c][0], dims[c][1], 1))
        for (_, b, _, co) in orc.TRUNK + orc.HEAD:
            setattr(self, b, nn.BatchNorm1d(co))
        self.dropout = nn.Dropout(0.3)
        self.C = num_classes
        sd = state if state is not None else orc.synth_state(num_classes, seed)
        self.load_state_dict({k: torch.tensor(v) for k, v in sd.items()}, strict=True)

    def forward(self, x):
        B, N, _ = x.shape
        a = x.transpose(2, 1)                                                        # pcs.py:103
        for conv, bn, _, _ in orc.TRUNK:
            a = F.relu(getattr(self, bn)(getattr(self, conv)(a)))                    # pcs.py:106-113
            if conv == "conv2":
                pf = a
        g = torch.max(a, 2, keepdim=True)[0]                                         # pcs.py:114
        a = torch.cat([pf, g.repeat(1, 1, N)], 1)                                    # pcs.py:117-120
        a = self.dropout(F.relu(self.bn_seg1(self.seg_conv1(a))))                    # pcs.py:123-124
        a = self.dropout(F.relu(self.bn_seg2(self.seg_conv2(a))))                    # pcs.py:125-126
        a = F.relu(self.bn_seg3(self.seg_conv3(a)))                                  # pcs.py:127
        return self.seg_conv4(a).transpose(2, 1)                                     # pcs.py:128-131


def time_stock_torch_on_gpu(mode, B, N, C, steps=10, warmup=3, device="cuda"):
    """points/s of StockTorchModel with torch eager on `device`, once with TF32 convolutions/matmuls allowed (torch's
    default for cuDNN convolutions) and once in IEEE fp32.  Train step = pcs.py:241-255 (zero_grad, forward, weighted CE,
    backward, Adam step, loss.item()); eval = pcs.py:450-452 under no_grad."""
    out = {}
    dev = torch.device(device)
    g = torch.Generator(device="cpu").manual_seed(5)
    x = torch.rand(B, N, 4, generator=g).to(dev)
    labels = torch.randint(0, C, (B, N), generator=g).to(dev)
    cw = torch.ones(C, device=dev)
    saved = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    try:
        for name, tf32 in (("tf32", True), ("ieee", False)):
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cuda.matmul.allow_tf32 = tf32
            model = StockTorchModel(C).to(dev)
            if mode == "train":
                model.train()
                opt = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=1e-4)
                crit = torch.nn.CrossEntropyLoss(weight=cw, ignore_index=-1)

                def step():
                    opt.zero_grad()
                    loss = crit(model(x).contiguous().view(-1, C), labels.view(-1))
                    loss.backward()
                    opt.step()
                    return loss.item()
            else:
                model.eval()

                def step():
                    with torch.no_grad():
                        return torch.argmax(model(x), dim=2)
            for _ in range(warmup):
                step()
            torch.cuda.synchronize(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                step()
            e1.record()
            torch.cuda.synchronize(dev)
            out[name] = B * N / (e0.elapsed_time(e1) / steps * 1e-3)
            del model
            torch.cuda.empty_cache()
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = saved
    return out
