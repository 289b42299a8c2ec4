"""Drop-in replacement of the reference `PointNetSegmentation` (point_cloud_segmentation.py:65-133).

Same constructor, same `forward((B, max_points, 4)) -> (B, max_points, num_classes)`, same 65-entry
state_dict (the sub-modules below are only parameter containers: the math runs in libpcseg_b200.so).
"""
import torch
import torch.nn as nn

from .engine import Engine, param_layout, bn_layout

_CONVS = ["conv1", "conv2", "conv3", "conv4", "conv5", "global_feat", "seg_conv1", "seg_conv2", "seg_conv3", "seg_conv4"]
_BNS = ["bn1", "bn2", "bn3", "bn4", "bn5", "bn_global", "bn_seg1", "bn_seg2", "bn_seg3"]


class _SegTrainFn(torch.autograd.Function):
    """Autograd bridge for the drop-in path: forward = pcseg_forward_train, backward = pcseg_backward
    with the caller's dlogits (whatever loss the user put on top, e.g. pcs.py:216,251)."""

    @staticmethod
    def forward(ctx, module, x, lengths, *params):
        logits = module._run_train_forward(x, lengths=lengths)
        ctx.module = module
        ctx.save_for_backward(x)
        ctx.token = module._fwd_token
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        module = ctx.module
        (x,) = ctx.saved_tensors
        if ctx.token != module._fwd_token:
            raise RuntimeError("pcseg_b200: activations of this forward were overwritten by a later training forward; "
                               "call backward() before the next forward (as the reference loop does, pcs.py:244-254)")
        grads = module._run_backward(x, dlogits.contiguous().float())
        return (None, None, None) + tuple(grads)


class PointNetSegmentation(nn.Module):
    def __init__(self, num_classes, input_dim=4):
        super(PointNetSegmentation, self).__init__()
        if input_dim != 4:
            raise NotImplementedError("pcseg_b200 kernels are specialised for input_dim=4 (x, y, z, e)")
        # registration order == reference (pcs.py:70-96) so that state_dict order and default init RNG use match
        self.conv1 = nn.Conv1d(input_dim, 64, 1)
        self.conv2 = nn.Conv1d(64, 64, 1)
        self.conv3 = nn.Conv1d(64, 64, 1)
        self.conv4 = nn.Conv1d(64, 128, 1)
        self.conv5 = nn.Conv1d(128, 1024, 1)
        self.global_feat = nn.Conv1d(1024, 1024, 1)
        self.seg_conv1 = nn.Conv1d(1088, 512, 1)
        self.seg_conv2 = nn.Conv1d(512, 256, 1)
        self.seg_conv3 = nn.Conv1d(256, 128, 1)
        self.seg_conv4 = nn.Conv1d(128, num_classes, 1)
        self.bn1 = nn.BatchNorm1d(64)
        self.bn2 = nn.BatchNorm1d(64)
        self.bn3 = nn.BatchNorm1d(64)
        self.bn4 = nn.BatchNorm1d(128)
        self.bn5 = nn.BatchNorm1d(1024)
        self.bn_global = nn.BatchNorm1d(1024)
        self.bn_seg1 = nn.BatchNorm1d(512)
        self.bn_seg2 = nn.BatchNorm1d(256)
        self.bn_seg3 = nn.BatchNorm1d(128)
        self.dropout = nn.Dropout(0.3)

        self.num_classes = num_classes
        # arithmetic of the INFERENCE path: "bf16" (bf16 tensor-core operands, fp32 accumulation: logits within ~3e-3 of
        # max|logit|) or "bf16x3" (split-bf16: every operand is a bf16 pair hi + lo and every GEMM accumulates the three
        # products hi*hi + hi*lo + lo*hi in fp32: fp32-grade logits, <= 1e-3 relative, at about 3x the tensor-core work)
        self.precision = "bf16"
        self._engine = None
        self._flat = None          # dict(params, grads, bn) of flat arenas
        self._fwd_token = 0
        self._manual_version = 0   # bumped when the arena is modified behind autograd's back (fused optimizer)

    def __getstate__(self):
        """copy.deepcopy / pickling: the engine (ctypes handles, workspaces) and the arena bookkeeping are per-instance
        run-time state and are rebuilt lazily on the next forward."""
        state = super().__getstate__() if hasattr(super(), "__getstate__") else self.__dict__.copy()
        state = dict(state)
        state["_engine"] = None
        state["_flat"] = None
        return state

    # ------------------------------------------------------------------ flat arenas
    def _param_list(self):
        ps = []
        for n in _CONVS:
            m = getattr(self, n)
            ps += [m.weight, m.bias]
        for n in _BNS:
            m = getattr(self, n)
            ps += [m.weight, m.bias]
        return ps

    def _flat_quick_ok(self, device):
        """Cheap per-step check used by the fused trainer: first / last parameter and the last running_var still alias the
        arenas (a full check of all 56 tensors costs ~50 us of host time per step)."""
        f = self._flat
        if f is None or f["params"].device != device:
            return False
        offs = f["offs"]
        base = f["params"].data_ptr()
        last_bn = getattr(self, _BNS[-1])
        return (self.conv1.weight.data_ptr() == base + 4 * offs[0][0]
                and last_bn.bias.data_ptr() == base + 4 * offs[-1][0]
                and last_bn.running_var.data_ptr() == f["bn"].data_ptr() + 4 * (f["bn"].numel() - last_bn.num_features))

    def _ensure_flat(self, device):
        """Parameters / running stats live in flat fp32 arenas (one contiguous gradient arena for the
        all-reduce); re-flatten if .to()/load_state_dict replaced the tensors."""
        offs, total = param_layout(self.num_classes)
        ps = self._param_list()
        f = self._flat
        ok = f is not None and f["params"].device == device
        if ok:
            base = f["params"].data_ptr()
            for p, (o, n) in zip(ps, offs):
                if p.data_ptr() != base + 4 * o or p.dtype != torch.float32:
                    ok = False
                    break
        if ok:
            bbase = f["bn"].data_ptr()
            boffs, _ = bn_layout()
            for name, (om, ov) in zip(_BNS, boffs):
                m = getattr(self, name)
                if m.running_mean.data_ptr() != bbase + 4 * om or m.running_var.data_ptr() != bbase + 4 * ov:
                    ok = False
                    break
        if ok:
            return f
        flat_p = torch.empty(total, dtype=torch.float32, device=device)
        for p, (o, n) in zip(ps, offs):
            view = flat_p[o:o + n].view(p.shape)
            view.copy_(p.data.to(device=device, dtype=torch.float32))
            p.data = view
        boffs, btotal = bn_layout()
        flat_bn = torch.empty(btotal, dtype=torch.float32, device=device)
        for name, (om, ov) in zip(_BNS, boffs):
            m = getattr(self, name)
            c = m.num_features
            flat_bn[om:om + c].copy_(m.running_mean.to(device=device, dtype=torch.float32))
            flat_bn[ov:ov + c].copy_(m.running_var.to(device=device, dtype=torch.float32))
            m.running_mean = flat_bn[om:om + c]
            m.running_var = flat_bn[ov:ov + c]
            if m.num_batches_tracked.device != device:
                m.num_batches_tracked = m.num_batches_tracked.to(device)
        # (storage padded to a multiple of 4 floats: the NVLink peer all-reduce moves float4s, peer_allreduce.cuh)
        flat_g = torch.zeros((total + 3) // 4 * 4, dtype=torch.float32, device=device)[:total]
        self._flat = dict(params=flat_p, grads=flat_g, bn=flat_bn, offs=offs)
        self._manual_version += 1
        return self._flat

    def _get_engine(self, device):
        if self._engine is None or self._engine.device != device:
            self._engine = Engine(self.num_classes, device)
        return self._engine

    def _weights_key(self):
        ks = [self._manual_version]
        for p in self._param_list():
            ks.append(p._version)
        for name in _BNS:
            m = getattr(self, name)
            ks.append(m.running_mean._version)
            ks.append(m.running_var._version)
        return tuple(ks)

    def grad_views(self):
        f = self._flat
        return [f["grads"][o:o + n].view(p.shape) for p, (o, n) in zip(self._param_list(), f["offs"])]

    # ------------------------------------------------------------------ kernels
    def _check_input(self, x):
        if getattr(self, "_is_replica", False):
            raise RuntimeError("pcseg_b200.PointNetSegmentation cannot run as an nn.DataParallel replica (pcs.py:209-211): use one "
                               "process per GPU with pcseg_b200.FusedTrainer (INTEGRATION.md, multi-GPU), which reproduces the "
                               "DataParallel result")
        batch_size, max_points, _ = x.shape      # same unpack (and ValueError) as pcs.py:100
        if x.shape[2] != 4:
            raise RuntimeError(f"expected input with 4 channels (x, y, z, e), got {x.shape[2]}")
        if not x.is_cuda:
            raise RuntimeError("pcseg_b200: input must be a CUDA tensor (sm_100a kernels only, no CPU fallback)")
        return x.contiguous().float()

    def _run_train_forward(self, x, labels=None, class_w=None, ce=None, state=None, lengths=None):
        f = self._ensure_flat(x.device)
        eng = self._get_engine(x.device)
        p = float(self.dropout.p) if self.dropout.training else 0.0
        # with a device-resident step state the dropout seed lives on the GPU (CUDA-graph friendly)
        seed = int(torch.empty((), dtype=torch.int64).random_().item()) if (p > 0 and state is None) else 0
        logits = eng.forward_train(x, f["params"], f["bn"], seed, p, labels, class_w, ce, state, lengths=lengths)
        torch._foreach_add_([getattr(self, n).num_batches_tracked for n in _BNS], 1)
        self._fwd_token += 1
        self._manual_version += 1      # running statistics changed
        return logits

    def _run_backward(self, x, dlogits):
        f = self._flat
        eng = self._get_engine(x.device)
        eng.backward(x, f["params"], f["grads"], dlogits=dlogits)
        return [g.clone() for g in self.grad_views()]

    def set_precision(self, precision):
        """Inference arithmetic: "bf16" (default) or "bf16x3" (fp32-grade, see __init__).  Training always runs bf16."""
        if precision not in ("bf16", "bf16x3"):
            raise ValueError("precision must be 'bf16' or 'bf16x3'")
        self.precision = precision
        return self

    def forward(self, x, lengths=None):
        """`lengths` (optional, B ints: real points per cloud of a zero-padded batch, pcs.py:44-63) selects ragged
        execution: pad rows cost nothing and the result is the one of the padded batch (see include/pcseg_b200.h)."""
        x = self._check_input(x)
        if self.training:
            if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
                self._ensure_flat(x.device)
                return _SegTrainFn.apply(self, x, lengths, *self._param_list())
            return self._run_train_forward(x, lengths=lengths)
        f = self._ensure_flat(x.device)
        eng = self._get_engine(x.device)
        return eng.forward_eval(x, f["params"], f["bn"], self._weights_key(), lengths=lengths, x3=self.precision == "bf16x3")

    @torch.no_grad()
    def predict(self, x, lengths=None):
        """Fused inference + argmax (pcs.py:450-452): returns (logits, labels int64 (B, N))."""
        x = self._check_input(x)
        if self.training:
            raise RuntimeError("predict() is an eval-mode call; use model.eval() first")
        f = self._ensure_flat(x.device)
        eng = self._get_engine(x.device)
        return eng.forward_eval(x, f["params"], f["bn"], self._weights_key(), want_labels=True, lengths=lengths,
                                x3=self.precision == "bf16x3")


    @torch.no_grad()
    def predict_point_sharded(self, x_local, group=None, reduce_max=None):
        """Inference of clouds whose POINTS are split over the ranks of `group` (one 1M-point scene on several
        GPUs): x_local (B, N_local, 4) is this rank's slice of the same B clouds (slices may differ in length).  The only
        exchange is a MAX all-reduce of the (B, 1024) pooled feature between global_feat and seg_conv1 (pcs.py:114-117);
        returns (logits, labels) of the local points, bit-identical to the un-sharded forward.  `reduce_max` overrides
        the collective (tests)."""
        import torch.distributed as dist
        x = self._check_input(x_local)
        if self.training:
            raise RuntimeError("predict_point_sharded() is an eval-mode call; use model.eval() first")
        f = self._ensure_flat(x.device)
        eng = self._get_engine(x.device)
        if reduce_max is None:
            def reduce_max(pooled):
                if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
                    dist.all_reduce(pooled, op=dist.ReduceOp.MAX, group=group)
        return eng.forward_eval_sharded(x, f["params"], f["bn"], self._weights_key(), reduce_max, want_labels=True,
                                        x3=self.precision == "bf16x3")

    @torch.no_grad()
    def evaluate(self, x, labels, class_weights=None, lengths=None):
        """One validation batch without host synchronisation: replaces the per-batch loss / accuracy code of
        pcs.py:289-304 and the second F1 sweep of pcs.py:319-343.  Returns device tensors: logits, loss (weighted-mean CE),
        correct, valid and the C x C confusion matrix (rows = true class, columns = prediction)."""
        x = self._check_input(x)
        if self.training:
            raise RuntimeError("evaluate() is an eval-mode call; use model.eval() first")
        f = self._ensure_flat(x.device)
        eng = self._get_engine(x.device)
        logits = eng.forward_eval(x, f["params"], f["bn"], self._weights_key(), lengths=lengths, x3=self.precision == "bf16x3")
        cw = None if class_weights is None else torch.as_tensor(class_weights, dtype=torch.float32, device=x.device).contiguous()
        ce, conf, _ = eng.eval_metrics(logits, labels.contiguous(), cw)
        f64, i64 = ce.view(torch.float64), ce.view(torch.int64)
        return dict(logits=logits, loss=f64[0] / f64[1], correct=i64[2], valid=i64[3], confusion=conf)


class PredictStream:
    """Host-fed inference with the copies off the critical path (replaces the synchronous `.to(device)` ... `.cpu()` round
    trip of pcs.py:446-454): the host->device copy of batch i+1 and the device->host copy of batch i-1's labels run on
    side streams while batch i computes.  Double buffered: at most two batches are in flight, and `result(ticket)` of a
    batch must be consumed before the second `submit` after it.

        ps = PredictStream(model)
        t = ps.submit(points_pinned)                 # (B, N, 4) fp32 host tensor, ideally pinned
        labels = ps.result(t)                        # (B, N) int64 pinned host tensor (argmax labels, pcs.py:452)
    """

    def __init__(self, model):
        if model.training:
            raise RuntimeError("PredictStream is an eval-mode helper; use model.eval() first")
        self.model = model
        self.device = next(model.parameters()).device
        self.h2d = torch.cuda.Stream(device=self.device)
        self.d2h = torch.cuda.Stream(device=self.device)
        self.slots = [None, None]
        self.next = 0

    @torch.no_grad()
    def submit(self, points_host, lengths=None):
        slot, self.next = self.next, self.next ^ 1
        s = self.slots[slot]
        if s is None or s["x"].shape != points_host.shape:
            B, N, _ = points_host.shape
            s = dict(x=torch.empty(points_host.shape, dtype=torch.float32, device=self.device),
                     labels_host=torch.empty((B, N), dtype=torch.int64).pin_memory(),
                     h2d_done=torch.cuda.Event(), compute_done=torch.cuda.Event(), d2h_done=torch.cuda.Event(), keep=None)
            self.slots[slot] = s
        cur = torch.cuda.current_stream(self.device)
        with torch.cuda.stream(self.h2d):
            self.h2d.wait_event(s["compute_done"])           # the batch that last used this input buffer has been computed
            s["x"].copy_(points_host, non_blocking=True)
            s["h2d_done"].record(self.h2d)
        cur.wait_event(s["h2d_done"])
        logits, labels = self.model.predict(s["x"], lengths=lengths)
        s["compute_done"].record(cur)
        with torch.cuda.stream(self.d2h):
            self.d2h.wait_event(s["compute_done"])
            s["labels_host"].copy_(labels, non_blocking=True)
            s["d2h_done"].record(self.d2h)
        s["keep"] = (logits, labels)                         # alive until the copy has been consumed
        return slot

    def result(self, ticket, want_logits=False):
        s = self.slots[ticket]
        s["d2h_done"].synchronize()
        return (s["labels_host"], s["keep"][0]) if want_logits else s["labels_host"]


def lengths_from_masks(masks):
    """Real points per cloud from the `masks` tensor of the reference's collate_fn (pcs.py:58-63: True on the first
    len(points) rows of every cloud)."""
    return masks.sum(dim=1).to(torch.int64).cpu().tolist()


def f1_scores(confusion):
    """Per-class F1, macro F1 and weighted F1 from a confusion matrix (what sklearn.f1_score returns at pcs.py:341-343)."""
    conf = confusion.to(torch.float64)
    tp = conf.diag()
    support = conf.sum(dim=1)
    predicted = conf.sum(dim=0)
    denom = support + predicted
    f1 = torch.where(denom > 0, 2 * tp / denom.clamp(min=1), torch.zeros_like(tp))
    present = (support + predicted) > 0
    macro = f1[present].mean() if present.any() else f1.sum() * 0
    weighted = (f1 * support).sum() / support.sum().clamp(min=1)
    return f1, macro, weighted


def load_checkpoint(path, map_location=None):
    """Load a `best_model.pth` written by the reference (pcs.py:373-382): reads `num_classes` and
    `model_state_dict`, strips a DataParallel `module.` prefix if present (pcs.py:410-428)."""
    ckpt = torch.load(path, map_location=map_location, weights_only=False)
    model = PointNetSegmentation(num_classes=ckpt["num_classes"])
    sd = ckpt["model_state_dict"]
    if any(k.startswith("module.") for k in sd):
        sd = {k.replace("module.", "", 1): v for k, v in sd.items()}
    model.load_state_dict(sd, strict=True)
    return model, ckpt
