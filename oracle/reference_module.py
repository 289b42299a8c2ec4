"""Runs the UNMODIFIED reference module for timing.  BENCH / TEST INFRASTRUCTURE ONLY (same rules as the rest of oracle/).

`__graft_entry__.build()` copies the reference's single source file, point_cloud_segmentation.py ("pcs.py"), into the
git-ignored directory baseline/_ref/ while /root/reference is visible (build container); the copy travels to the GPU box
with the tree.  Here it is imported as is, with `h5py` stubbed: the reference imports it at module scope (pcs.py:6) but
touches it only in PointCloudDataset.__init__ (pcs.py:22-23), which the hot path never calls.  Nothing of
the repo's own kernels, models or engine is on this path.
"""
import contextlib
import importlib.util
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_FILE = os.path.join(ROOT, "baseline", "_ref", "point_cloud_segmentation.py")


def available():
    return os.path.exists(REF_FILE)


def load():
    """import baseline/_ref/point_cloud_segmentation.py (its import-time prints, pcs.py:16-18, go to stderr)"""
    sys.modules.setdefault("h5py", types.ModuleType("h5py"))
    spec = importlib.util.spec_from_file_location("pcs_reference", REF_FILE)
    mod = importlib.util.module_from_spec(spec)
    with contextlib.redirect_stdout(sys.stderr):
        spec.loader.exec_module(mod)
    return mod


class ReferenceCpuStep:
    """The reference's own module, loss, optimizer and call sequence on the host cores:
    train: pcs.py:216-217 (criterion, Adam lr 1e-3 wd 1e-4) and the loop body pcs.py:241-258;
    eval : pcs.py:432, 450-452 (model.eval(), no_grad forward, argmax)."""

    def __init__(self, num_classes, seed=1234, threads=None):
        import torch
        self.torch = torch
        if threads:
            torch.set_num_threads(threads)
        self.pcs = load()
        torch.manual_seed(seed)
        self.C = num_classes
        self.model = self.pcs.PointNetSegmentation(num_classes=num_classes)          # pcs.py:206
        self.optimizer = torch.optim.Adam(self.model.parameters(), lr=0.001, weight_decay=1e-4)   # pcs.py:217
        self.criterion = None

    def train_step(self, x, labels, class_w=None, dropout_p=0.3):
        torch = self.torch
        if self.criterion is None:
            self.criterion = torch.nn.CrossEntropyLoss(ignore_index=-1, weight=class_w)           # pcs.py:216
        self.model.train()                                                                        # pcs.py:230
        self.optimizer.zero_grad()                                                                # pcs.py:241
        outputs = self.model(x)                                                                   # pcs.py:244
        outputs = outputs.contiguous().view(-1, self.C)                                           # pcs.py:247
        loss = self.criterion(outputs, labels.view(-1))                                           # pcs.py:248-251
        loss.backward()                                                                           # pcs.py:254
        self.optimizer.step()                                                                     # pcs.py:255
        return loss.item()                                                                        # pcs.py:258

    def eval_step(self, x):
        torch = self.torch
        self.model.eval()                                                                         # pcs.py:432
        with torch.no_grad():                                                                     # pcs.py:450
            outputs = self.model(x)                                                               # pcs.py:451
            return torch.argmax(outputs, dim=2)                                                   # pcs.py:452
