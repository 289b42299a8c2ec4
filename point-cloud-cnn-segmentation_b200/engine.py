"""Host-side driver of the C ABI: owns contexts, workspaces and flat arenas.
PyTorch is used for device memory and streams only."""
import ctypes as C
from collections import OrderedDict

import torch

from ._lib import lib, check, ptr

MAX_CLASSES = 32
NUM_PARAM_TENSORS = 38
NUM_BN = 9


def param_layout(num_classes):
    """[(offset, numel)] of the 38 parameter tensors in state_dict order + total."""
    offs = [(int(lib.pcseg_param_offset(num_classes, t)), int(lib.pcseg_param_numel(num_classes, t)))
            for t in range(NUM_PARAM_TENSORS)]
    return offs, int(lib.pcseg_param_count(num_classes))


def bn_layout():
    offs = [(int(lib.pcseg_bn_buffer_offset(j, 0)), int(lib.pcseg_bn_buffer_offset(j, 1))) for j in range(NUM_BN)]
    return offs, int(lib.pcseg_bn_buffer_count())


class _Binding:
    """One C context bound to one batch shape.  The device workspace is owned by the Engine and SHARED by all bindings of
    the same mode (activations only have to live from a forward to its backward)."""

    def __init__(self, num_classes, B, N, mode):
        """mode: 0 inference (bf16), 1 training, 2 inference with split-bf16 operands (fp32-grade logits)"""
        self.handle = C.c_void_p()
        check(lib.pcseg_create(C.byref(self.handle), num_classes), "pcseg_create")
        self.shape = (B, N, int(mode))
        self.nbytes = int(lib.pcseg_workspace_bytes(B, N, num_classes, int(mode)))
        if self.nbytes <= 0:
            raise ValueError(f"unsupported shape B={B} N={N} C={num_classes}")
        self.ws_ptr = None
        self.eval_key = None

    def bind(self, ws_ptr, ws_bytes):
        B, N, mode = self.shape
        check(lib.pcseg_bind(self.handle, B, N, C.c_void_p(ws_ptr), ws_bytes, mode), "pcseg_bind")
        self.ws_ptr = ws_ptr
        self.eval_key = None

    def __del__(self):
        try:
            if self.handle:
                lib.pcseg_destroy(self.handle)
        except Exception:
            pass


def train_ragged_min_pad():
    """Smallest pad fraction of a TRAINING batch for which the packed (ragged) step is used.  The dense step runs the
    folded-BatchNorm kernels inside CUDA graphs, the packed step the materialising kernels launched eagerly: measured on
    8 x 16 384 points the packed step wins below ~77 % valid points.  The padded batch IS the reference computation, so
    routing a mostly-full batch to the dense path changes nothing but speed.  PCSEG_RAGGED_MIN_PAD overrides."""
    import os
    return float(os.environ.get("PCSEG_RAGGED_MIN_PAD", "0.25"))


def host_lengths(lengths, B, N, min_pad_fraction=0.0):
    """Per-cloud point counts of a zero-padded batch as a ctypes int array (None stays None).  Accepts a list, a numpy
    array or a tensor (a CUDA tensor costs one device->host read); the reference's collate_fn (pcs.py:44-63) knows
    them on the host: `masks.sum(1)`.  Returns None (dense execution of the padded batch) when less than
    `min_pad_fraction` of the rows are padding."""
    if lengths is None:
        return None
    vals = lengths.detach().cpu().tolist() if torch.is_tensor(lengths) else [int(v) for v in lengths]
    if len(vals) != B:
        raise ValueError(f"lengths has {len(vals)} entries for a batch of {B} clouds")
    vals = [int(v) for v in vals]
    if any(v < 0 or v > N for v in vals):
        raise ValueError(f"lengths must be within 0..{N}")
    if all(v == N for v in vals):
        return None                      # nothing is padded: the dense path is the same computation without the packing
    if 1.0 - sum(vals) / float(B * N) < min_pad_fraction:
        return None                      # mostly full: the padded batch on the dense path is faster
    return (C.c_int * B)(*vals)


def ragged_capacity(N, step=4096):
    """Ragged calls bind a padded length rounded up to a multiple of `step`: the padded length of a DataLoader batch is
    its longest cloud (pcs.py:50) and changes every batch, the binding (TMA descriptors) is reused as long as the
    capacity bucket does not change.  Costs nothing at run time: only real rows are computed."""
    return (N + step - 1) // step * step


class Engine:
    """One per (module, device).  Keeps one grow-only workspace per mode (train / eval) and a cache of per-shape
    bindings into it, so variable-length batches (every DataLoader batch of the reference has its own max_points,
    pcs.py:50) neither re-allocate device memory nor thrash: a new shape only costs the host-side TMA descriptor setup."""

    def __init__(self, num_classes, device, max_bindings=32):
        if not (1 <= num_classes <= MAX_CLASSES):
            raise ValueError(f"num_classes must be in 1..{MAX_CLASSES}")
        self.C = num_classes
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("pcseg_b200 runs on CUDA (sm_100a) devices only; there is no CPU fallback")
        self.bindings = OrderedDict()
        self.max_bindings = max_bindings
        self._ws = {True: None, False: None}          # mode -> (storage tensor, aligned ptr, usable bytes)

    def _workspace(self, train, need):
        ws = self._ws[train]
        if ws is None or ws[2] < need:
            # grow-only; all bindings of this mode point into the old block and are dropped
            for key in [k for k in self.bindings if (k[2] == 1) == train]:
                del self.bindings[key]
            self._ws[train] = None
            del ws
            # torch caching-allocator blocks are 512-byte aligned; over-allocate and align to 1024
            storage = torch.empty(need + 1024, dtype=torch.uint8, device=self.device)
            ptr_aligned = (storage.data_ptr() + 1023) & ~1023
            self._ws[train] = (storage, ptr_aligned, need)
        return self._ws[train]

    def binding(self, B, N, train, x3=False):
        train = bool(train)
        mode = 1 if train else (2 if x3 else 0)
        key = (B, N, mode)
        b = self.bindings.get(key)
        if b is None:
            b = _Binding(self.C, B, N, mode)
        _, ws_ptr, ws_bytes = self._workspace(train, b.nbytes)
        if key not in self.bindings:                 # (the workspace may just have dropped every cached binding)
            self.bindings[key] = b
            while len(self.bindings) > self.max_bindings:
                self.bindings.popitem(last=False)
        else:
            self.bindings.move_to_end(key)
        if b.ws_ptr != ws_ptr:
            with torch.cuda.device(self.device):
                b.bind(ws_ptr, ws_bytes)
        return b

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    # ---- eval
    def forward_eval(self, x, flat_params, flat_bn, weights_key, want_labels=False, lengths=None, x3=False):
        B, N, _ = x.shape
        lengths = host_lengths(lengths, B, N)
        if x3 and lengths is not None:
            raise ValueError("precision 'bf16x3' runs dense batches only (no ragged execution)")
        b = self.binding(B, N if lengths is None else ragged_capacity(N), False, x3=x3)
        with torch.cuda.device(self.device):
            if b.eval_key != weights_key:
                check(lib.pcseg_prepare_eval(b.handle, ptr(flat_params), ptr(flat_bn), self._stream()), "pcseg_prepare_eval")
                b.eval_key = weights_key
            logits = torch.empty((B, N, self.C), dtype=torch.float32, device=self.device)
            labels = torch.empty((B, N), dtype=torch.int64, device=self.device) if want_labels else None
            if lengths is not None:
                check(lib.pcseg_forward_eval_ragged(b.handle, ptr(x), lengths, N, ptr(logits), ptr(labels), self._stream()),
                      "pcseg_forward_eval_ragged")
            else:
                check(lib.pcseg_forward_eval(b.handle, ptr(x), ptr(logits), ptr(labels), self._stream()), "pcseg_forward_eval")
        return (logits, labels) if want_labels else logits

    def forward_eval_sharded(self, x, flat_params, flat_bn, weights_key, reduce_max, want_labels=False, x3=False):
        """Inference on this rank's slice of the points of B clouds: trunk, `reduce_max(pooled)` (the caller's MAX
        all-reduce over the ranks that hold the other slices, in place on a (B, 1024) fp32 tensor), head."""
        B, N, _ = x.shape
        b = self.binding(B, N, False, x3=x3)
        with torch.cuda.device(self.device):
            if b.eval_key != weights_key:
                check(lib.pcseg_prepare_eval(b.handle, ptr(flat_params), ptr(flat_bn), self._stream()), "pcseg_prepare_eval")
                b.eval_key = weights_key
            check(lib.pcseg_forward_eval_part(b.handle, ptr(x), None, None, 1, self._stream()), "pcseg_forward_eval_part")
            pooled_ptr = C.c_void_p()
            check(lib.pcseg_pooled_feature(b.handle, C.byref(pooled_ptr)), "pcseg_pooled_feature")
            storage = self._ws[False][0]
            off = pooled_ptr.value - storage.data_ptr()
            pooled = storage[off:off + B * 1024 * 4].view(torch.float32).view(B, 1024)      # a view of the workspace
            reduce_max(pooled)
            logits = torch.empty((B, N, self.C), dtype=torch.float32, device=self.device)
            labels = torch.empty((B, N), dtype=torch.int64, device=self.device) if want_labels else None
            check(lib.pcseg_forward_eval_part(b.handle, None, ptr(logits), ptr(labels), 2, self._stream()), "pcseg_forward_eval_part")
        return (logits, labels) if want_labels else logits

    # ---- train
    def forward_train(self, x, flat_params, flat_bn, seed, dropout_p, labels=None, class_w=None, ce=None, state=None, lengths=None):
        B, N, _ = x.shape
        lengths = host_lengths(lengths, B, N, train_ragged_min_pad())
        b = self.binding(B, N if lengths is None else ragged_capacity(N), True)
        self._train_binding = b              # backward runs on the binding of the latest training forward
        logits = torch.empty((B, N, self.C), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            if lengths is not None:
                check(lib.pcseg_forward_train_ragged(b.handle, ptr(x), lengths, N, ptr(flat_params), ptr(flat_bn),
                                                     C.c_ulonglong(seed & (2**64 - 1)), C.c_float(dropout_p), ptr(logits), ptr(labels),
                                                     ptr(class_w), ptr(ce), ptr(state), self._stream()), "pcseg_forward_train_ragged")
                return logits
            check(lib.pcseg_forward_train(b.handle, ptr(x), ptr(flat_params), ptr(flat_bn), C.c_ulonglong(seed & (2**64 - 1)),
                                          C.c_float(dropout_p), ptr(logits), ptr(labels), ptr(class_w), ptr(ce), ptr(state), self._stream()),
                  "pcseg_forward_train")
        return logits

    def backward(self, x, flat_params, flat_grads, dlogits=None, logits=None, labels=None, class_w=None, wsum=None, phase=0):
        B, N, _ = x.shape
        b = getattr(self, "_train_binding", None)
        if b is None or b.shape[0] != B or b.shape[1] < N or b.ws_ptr is None or self.bindings.get(b.shape) is not b:
            b = self.binding(B, N, True)
        with torch.cuda.device(self.device):
            check(lib.pcseg_backward(b.handle, ptr(x), ptr(flat_params), ptr(dlogits), ptr(logits), ptr(labels), ptr(class_w),
                                     ptr(wsum), ptr(flat_grads), phase, self._stream()), "pcseg_backward")

    def adam(self, flat_params, flat_grads, m, v, step, lr, betas, eps, weight_decay, grad_scale=1.0, state=None, grad_div=None):
        with torch.cuda.device(self.device):
            check(lib.pcseg_adam_step(ptr(flat_params), ptr(flat_grads), ptr(m), ptr(v), flat_params.numel(), step, lr, betas[0],
                                      betas[1], eps, weight_decay, grad_scale, ptr(state), ptr(grad_div), self._stream()), "pcseg_adam_step")

    def eval_metrics(self, logits, labels, class_w=None, want_pred=False):
        """Weighted CE sums, accuracy counters and confusion matrix of eval-mode logits (one kernel, no host sync)."""
        P = logits.shape[0] * logits.shape[1]
        ce = torch.zeros(32, dtype=torch.uint8, device=self.device)
        conf = torch.zeros((self.C, self.C), dtype=torch.int64, device=self.device)
        pred = torch.empty(logits.shape[:2], dtype=torch.int64, device=self.device) if want_pred else None
        with torch.cuda.device(self.device):
            check(lib.pcseg_eval_metrics(ptr(logits), ptr(labels), P, self.C, ptr(class_w), ptr(ce), ptr(conf), ptr(pred), self._stream()),
                  "pcseg_eval_metrics")
        return ce, conf, pred

    def step_advance(self, state, betas):
        with torch.cuda.device(self.device):
            check(lib.pcseg_step_advance(ptr(state), betas[0], betas[1], self._stream()), "pcseg_step_advance")


DEBUG_KINDS = {"y": 0, "act": 1, "dz": 2, "dy": 3, "bnp": 4, "coef": 5, "stats_f": 6, "stats_b": 7, "g": 8, "ystar": 9,
               "argidx": 10, "cb": 11, "dcb": 12, "dzv": 13, "gram": 14, "colsum": 15, "qraw": 16, "bwf": 17, "cstf": 18,
               "s5b": 19, "gc5b": 20, "side": 21, "rowslot": 22, "cloudsum6": 23, "cst6": 24, "wcat6": 25, "gram6": 26, "colsum6": 27}


def debug_tensor(engine, B, N, kind, layer=0):
    """Copy an internal training-workspace tensor out (tests only): returns a torch tensor."""
    b = engine.binding(B, N, True)
    rows, cols, eb = C.c_longlong(), C.c_longlong(), C.c_int()
    k = DEBUG_KINDS[kind]
    check(lib.pcseg_debug_copy(b.handle, k, layer, None, 0, C.byref(rows), C.byref(cols), C.byref(eb), None), "pcseg_debug_copy")
    dt = {2: torch.bfloat16, 8: torch.float64, 4: torch.int32 if kind in ("argidx", "rowslot") else torch.float32}[eb.value]
    out = torch.empty((rows.value, cols.value), dtype=dt, device=engine.device)
    with torch.cuda.device(engine.device):
        check(lib.pcseg_debug_copy(b.handle, k, layer, ptr(out), out.numel() * out.element_size(), None, None, None,
                                   engine._stream()), "pcseg_debug_copy")
    return out


def profile_enable(engine, B, N, on=True, train=True):
    b = engine.binding(B, N, train)
    check(lib.pcseg_profile_reset(b.handle))
    check(lib.pcseg_profile_enable(b.handle, int(on)))


def profile_read(engine, B, N, train=True):
    """{tag: (total_ms, launches)} for the tcgen05 GEMMs (tags: conv index + 0 / 16 / 32 / 48 / 64 = training forward / data
    gradient / weight gradient / Gram and fold GEMMs / inference trunk)."""
    b = engine.binding(B, N, train)
    out = {}
    for tag in [base + i for base in (0, 16, 32, 48, 64) for i in range(1, 9)] + list(range(80, 96)):
        ms, n = C.c_double(), C.c_longlong()
        check(lib.pcseg_profile_read(b.handle, tag, C.byref(ms), C.byref(n)))
        if n.value:
            out[tag] = (ms.value, n.value)
    return out


# tags >= 80: CUDA-core kernels of the training step (event-timed while profiling; several launches per step share a tag)
KERNEL_TAGS = {80: "k_gram_reduce", 81: "k_predict_bn", 82: "k_fold5_prep", 83: "k_pool_claim", 84: "k_pool_rows", 85: "k_gram_center",
               86: "k_fold_coef", 87: "k_fold_bwd", 88: "k_convert_multi", 89: "k_bn_relu (all layers)", 90: "k_bn_bwd_apply (all layers)",
               91: "k_head_fwd", 92: "k_head_bwd", 93: "memsets of the folded global_feat backward",
               94: "k_peer_allreduce (NVLink peer-memory gradient all-reduce, incl. waiting for the slowest rank)"}


def launch_count():
    return int(lib.pcseg_launch_count())
