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
