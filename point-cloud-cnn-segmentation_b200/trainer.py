"""Fused training step: forward + weighted cross-entropy + backward + Adam on the flat arenas,
data-parallel over point-cloud batches with one process per GPU (replaces the reference's
nn.DataParallel, pcs.py:209-211, and the loop body pcs.py:241-255).

Data-parallel semantics follow the reference's single-process DataParallel: one global weighted-mean
loss over all points of the global batch (so every rank divides by the ALL-REDUCED sum of class weights
and gradients are SUMMED), per-replica BatchNorm batch statistics (no SyncBN), per-replica dropout masks,
replicas that start from rank 0's parameters (DataParallel re-broadcasts them every step; here they are
broadcast once at construction and stay identical because every rank applies the same summed gradient),
rank 0's running statistics are the ones that survive (broadcast on demand with `sync_bn_buffers`).
"""
import torch
import torch.distributed as dist

from .model import PointNetSegmentation, _BNS


from .trainer_protocol import GradSync, grad_buckets  # noqa: E402


class FusedTrainer:
    def __init__(self, model: PointNetSegmentation, class_weights=None, lr=1e-3, betas=(0.9, 0.999), eps=1e-8,
                 weight_decay=1e-4, process_group=None, device=None, overlap=True, use_cuda_graph=True, sm_reserve=None):
        """sm_reserve: SMs that the persistent kernels leave free for the collectives that overlap backward (default:
        PCSEG_SM_RESERVE or 0).  Measured on 2 x B200 (tools/gpu_r2_mg_ab.sh): reserving 4 / 8 / 16 SMs costs 2.2 / 3.3 / 5.8 %
        of the step, more than the interference it removes, so the default is no reservation."""
        self.model = model
        self.device = torch.device(device if device is not None else torch.device("cuda", torch.cuda.current_device()))
        self.model.to(self.device)
        self.flat = model._ensure_flat(self.device)
        self.engine = model._get_engine(self.device)
        n = self.flat["params"].numel()
        self.exp_avg = torch.zeros(n, dtype=torch.float32, device=self.device)
        self.exp_avg_sq = torch.zeros(n, dtype=torch.float32, device=self.device)
        self.step_count = 0
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        self.class_w = None if class_weights is None else torch.as_tensor(class_weights, dtype=torch.float32, device=self.device).contiguous()
        self.pg = process_group
        self.sync = GradSync(self.flat["grads"], process_group)
        self.world = self.sync.world
        self.distributed = self.world > 1
        self.rank = dist.get_rank(process_group) if self.distributed else 0
        self.overlap = overlap
        if sm_reserve is None:
            import os
            sm_reserve = int(os.environ.get("PCSEG_SM_RESERVE", "0"))
        if sm_reserve > 0:
            from ._lib import lib
            lib.pcseg_set_sm_limit(torch.cuda.get_device_properties(self.device).multi_processor_count - sm_reserve)
        if self.distributed:
            # replicas = rank 0's model (nn.DataParallel replicates module 0, pcs.py:211): do not rely on identical seeding
            dist.broadcast(self.flat["params"], src=0, group=self.pg)
            dist.broadcast(self.flat["bn"], src=0, group=self.pg)
            for name in _BNS:
                dist.broadcast(getattr(self.model, name).num_batches_tracked, src=0, group=self.pg)
            self.model._manual_version += 1
        # 32-byte CE accumulator {loss_num f64, w_sum f64, correct u64, valid u64}
        self.ce_raw = torch.zeros(32, dtype=torch.uint8, device=self.device)
        self.ce_f64 = self.ce_raw.view(torch.float64)
        self.ce_i64 = self.ce_raw.view(torch.int64)
        # {loss numerator, sum of class weights} of the step: both are known after the forward and all-reduced together
        self.lw = torch.zeros(2, dtype=torch.float64, device=self.device)
        self.loss_num, self.wsum = self.lw[0:1], self.lw[1:2]
        # device-resident step state (pcseg_step_state): dropout seed, Adam step / bias corrections, learning rate
        self.state = torch.zeros(32, dtype=torch.uint8, device=self.device)
        seed_t = torch.empty((), dtype=torch.int64).random_().to(self.device)
        if self.distributed:
            dist.broadcast(seed_t, src=0, group=self.pg)
        # one base seed, a different stream per rank: DataParallel replicas draw independent dropout masks
        seed0 = (int(seed_t.item()) + self.rank * 0x632BE59BD9B4E019) & (2 ** 63 - 1)
        self.state.view(torch.int64)[0] = seed0
        self.one = torch.ones(1, dtype=torch.float64, device=self.device)
        self.state.view(torch.float32)[4] = lr
        # static outputs of a step
        self.out_loss = torch.zeros(1, dtype=torch.float64, device=self.device)
        self.out_counts = torch.zeros(2, dtype=torch.int64, device=self.device)
        self.early, self.late = grad_buckets(self.flat["offs"], n)
        self.last_logits = None
        # CUDA graph of the whole step (captured on the third step of a given batch shape)
        self.use_cuda_graph = use_cuda_graph
        self.profiling = False          # set by the bench's per-kernel event timing: forces eager launches
        self._graph = None
        self._graph_key = None
        self._eager_steps_at_key = 0
        self._static_x = None
        self._static_labels = None
        self._lengths = None            # per-cloud point counts of the current batch (ragged execution) or None
        import os as _os
        self.one_graph = self.distributed and _os.environ.get("PCSEG_DDP_ONE_GRAPH", "0") == "1"
        # Gradient exchange: "peer" = one NVLink peer-memory kernel on the compute stream (the whole step is ONE CUDA graph,
        # csrc/peer_allreduce.cuh); "nccl" = bucketed NCCL all-reduce between graph segments.  PCSEG_COMM overrides.
        self.comm = _os.environ.get("PCSEG_COMM", "peer") if self.distributed else "none"
        self.peer = None
        self.peer_events = []
        self._peer_skip = _os.environ.get("PCSEG_PEER_SKIP", "0") == "1"     # experiments only: no gradient exchange at all
        if self.comm == "peer":
            if self.world not in (2, 4, 8):
                self.comm = "nccl"
            else:
                from .peer_ar import PeerAllReduce
                self.lw_global = torch.zeros(2, dtype=torch.float64, device=self.device)
                self.peer = PeerAllReduce(self.flat["grads"], self.lw, self.lw_global, group=self.pg)

    def set_lr(self, lr):
        self.lr = lr
        self.state.view(torch.float32)[4] = lr

    # ------------------------------------------------------------------ one step = 4 capturable segments + collectives
    # Collectives are never captured (they run eagerly between the graph segments), so the same code serves one GPU and
    # data-parallel ranks:  [forward] -> async all-reduce(loss num, sum w) -> [backward phase 1] -> async all-reduce(bucket 1)
    #                       -> [backward phase 2] -> async all-reduce(bucket 2), wait for all three -> [Adam + outputs]
    # Data-parallel ranks back-propagate the UN-normalised loss (wsum = 1) and Adam divides by the all-reduced sum of class
    # weights, so no collective sits between forward and backward (`deferred`); a single GPU normalises in backward as
    # before and keeps true gradients in the arena.
    @property
    def deferred(self):
        return self.distributed and (self.overlap or self.peer is not None)

    def _seg_forward(self, x, labels):
        self.engine.step_advance(self.state, self.betas)
        self.last_logits = self.model._run_train_forward(x, labels=labels, class_w=self.class_w, ce=self.ce_raw, state=self.state,
                                                         lengths=self._lengths)
        self.lw.copy_(self.ce_f64[0:2])

    def _seg_backward(self, x, labels, phase):
        f = self.flat
        self.engine.backward(x, f["params"], f["grads"], phase=phase, logits=self.last_logits, labels=labels, class_w=self.class_w,
                             wsum=self.one if self.deferred else self.wsum)

    def _seg_tail(self, x, labels):
        f = self.flat
        self.engine.adam(f["params"], f["grads"], self.exp_avg, self.exp_avg_sq, 1, self.lr, self.betas, self.eps, self.weight_decay,
                         state=self.state, grad_div=self.wsum if self.deferred else None)
        torch.div(self.loss_num, self.wsum, out=self.out_loss)
        self.out_counts.copy_(self.ce_i64[2:4])

    @property
    def global_wsum(self):
        """sum of class weights over the valid points of the GLOBAL batch of the latest step (device, fp64)"""
        return self.lw_global[1:2] if self.peer is not None else self.wsum

    def _seg_peer_step(self, x, labels):
        """data-parallel step with the peer-memory all-reduce: forward, un-normalised backward, ONE all-reduce kernel (gradient
        arena + {loss numerator, sum w}), Adam dividing by the global sum of class weights -- all on the compute stream"""
        self._seg_forward(x, labels)
        self._seg_backward(x, labels, 0)
        if self.profiling:                       # (eager, event-timed pass of the bench: time the exchange kernel as well)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            self.peer.run()
            e1.record()
            self.peer_events.append((e0, e1))
        elif not self._peer_skip:
            self.peer.run()
        else:
            self.lw_global.copy_(self.lw)
        f = self.flat
        self.engine.adam(f["params"], f["grads"], self.exp_avg, self.exp_avg_sq, 1, self.lr, self.betas, self.eps, self.weight_decay,
                         state=self.state, grad_div=self.lw_global[1:2])
        torch.div(self.lw_global[0:1], self.lw_global[1:2], out=self.out_loss)
        self.out_counts.copy_(self.ce_i64[2:4])

    def _segments(self):
        if self.peer is not None:
            return [self._seg_peer_step, None, None, None]
        if self.distributed and self.overlap:
            return [self._seg_forward, lambda x, l: self._seg_backward(x, l, 1), lambda x, l: self._seg_backward(x, l, 2), self._seg_tail]
        return [self._seg_forward, lambda x, l: self._seg_backward(x, l, 0), None, self._seg_tail]

    def _collective_after(self, i):
        """eager collectives that follow segment i"""
        if self.peer is not None:
            return
        f = self.flat
        self.sync.g = f["grads"]
        if i == 0:
            if self.deferred:
                self.sync.launch_tensor(self.lw)            # joined by the wait() ahead of Adam
            else:
                self.sync.reduce_normaliser(self.lw)
        elif i == 1:
            if self.distributed and self.overlap:
                self.sync.launch(self.early)                # NCCL runs on its own stream while phase 2 computes
            else:
                self.sync.launch([(0, f["grads"].numel())])
                self.sync.wait()
        elif i == 2:
            if self.distributed and self.overlap:
                self.sync.launch(self.late)
                self.sync.wait()

    def _run_step(self, x, labels, graphs=None):
        if graphs is not None and len(graphs) == 1:      # the whole step, collectives included, is ONE graph
            graphs[0].replay()
            return
        for i, seg in enumerate(self._segments()):
            if seg is not None:
                if graphs is not None:
                    graphs[i].replay()
                else:
                    seg(x, labels)
            self._collective_after(i)

    def _try_capture(self, x, labels):
        """Record the four segments (no execution).  Returns the list of graphs or None."""
        try:
            self._static_x = x.clone()
            self._static_labels = labels.clone()
            torch.cuda.synchronize(self.device)
            if self.one_graph:
                # compute segments AND the NCCL collectives between them in a single graph: no host launches, no idle gaps
                # between segments (the collectives' side-stream events become graph edges)
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._run_step(self._static_x, self._static_labels)
                return [g]
            graphs, pool = [], None
            for seg in self._segments():
                if seg is None:
                    graphs.append(None)
                    continue
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, pool=pool):
                    seg(self._static_x, self._static_labels)
                pool = g.pool()
                graphs.append(g)
            return graphs
        except Exception as e:                      # stay eager
            self.use_cuda_graph = False
            self._capture_error = repr(e)
            torch.cuda.synchronize(self.device)
            return None

    @torch.no_grad()
    def step(self, points, labels, lengths=None):
        """One optimizer step on this rank's shard.  points (B,N,4) fp32 and labels (B,N) int64 (-1 = pad) on
        the device.  Returns dict of device tensors: loss (global weighted mean), correct, valid.
        lengths (optional, B host ints = real points per cloud, `masks.sum(1)` of pcs.py:58-63) runs the step on the
        un-padded points only (ragged execution; launched eagerly, packed shapes change from batch to batch)."""
        m = self.model
        if not m.training:
            raise RuntimeError("FusedTrainer.step needs model.train()")
        x = m._check_input(points)
        labels = labels.contiguous()
        if not m._flat_quick_ok(self.device):
            self.flat = m._ensure_flat(self.device)
        if lengths is not None:
            from .engine import host_lengths, train_ragged_min_pad
            if host_lengths(lengths, x.shape[0], x.shape[1], train_ragged_min_pad()) is None:
                lengths = None                                  # nothing / little padded: dense (CUDA-graph) path
        self._lengths = lengths
        if lengths is not None:
            # packed step: launched eagerly on a capacity-bucketed binding (engine.ragged_capacity); graphs are left alone
            self._run_step(x, labels)
            self.step_count += 1
            m._fwd_token += 1
            m._manual_version += 1
            loss, counts = self.out_loss.clone(), self.out_counts.clone()
            return dict(loss=loss[0], correct=counts[0], valid=counts[1])
        ws_ptr = self.engine.binding(x.shape[0], x.shape[1], True).ws_ptr       # (also keeps the binding hot in the LRU)
        key = (tuple(x.shape), float(m.dropout.p) if m.dropout.training else 0.0, self.flat["params"].data_ptr(), ws_ptr)
        use_graph = self.use_cuda_graph and not self.profiling
        if key != self._graph_key:
            self._graph, self._graph_key, self._eager_steps_at_key = None, key, 0
        if use_graph and self._graph is None and self._eager_steps_at_key >= 2:
            self._graph = self._try_capture(x, labels)          # recording only; the replay below runs this step
        if use_graph and self._graph is not None:
            self._static_x.copy_(x, non_blocking=True)
            self._static_labels.copy_(labels, non_blocking=True)
            self._run_step(self._static_x, self._static_labels, self._graph)
        else:
            self._run_step(x, labels)
            self._eager_steps_at_key += 1
        self.step_count += 1
        m._fwd_token += 1
        m._manual_version += 1
        # clones: the static output buffers are overwritten by the next step (graph replay)
        loss, counts = self.out_loss.clone(), self.out_counts.clone()
        return dict(loss=loss[0], correct=counts[0], valid=counts[1])

    # ------------------------------------------------------------------ host-fed training (pinned memory -> device)
    def prefetch(self, points_host, labels_host):
        """Start the host->device copy of the NEXT batch on a side stream (double buffered) so that it overlaps the
        current step; replaces the synchronous `.to(device)` of pcs.py:237-238.  Returns a ticket for `step_prefetched`."""
        if not hasattr(self, "_h2d_stream"):
            self._h2d_stream = torch.cuda.Stream(device=self.device)
            self._h2d_slots = [None, None]
            self._h2d_next = 0
        slot = self._h2d_next
        self._h2d_next ^= 1
        buf = self._h2d_slots[slot]
        if buf is None or buf[0].shape != points_host.shape or buf[1].shape != labels_host.shape:
            buf = [torch.empty(points_host.shape, dtype=torch.float32, device=self.device),
                   torch.empty(labels_host.shape, dtype=torch.int64, device=self.device), torch.cuda.Event(), torch.cuda.Event()]
            self._h2d_slots[slot] = buf
        with torch.cuda.stream(self._h2d_stream):
            self._h2d_stream.wait_event(buf[3])          # the step that last read this slot has finished
            buf[0].copy_(points_host, non_blocking=True)
            buf[1].copy_(labels_host, non_blocking=True)
            buf[2].record(self._h2d_stream)
        return slot

    def step_prefetched(self, ticket):
        buf = self._h2d_slots[ticket]
        torch.cuda.current_stream(self.device).wait_event(buf[2])
        out = self.step(buf[0], buf[1])
        buf[3].record(torch.cuda.current_stream(self.device))
        return out

    def sync_bn_buffers(self, src=0):
        """DataParallel keeps replica 0's running statistics; broadcast them when a checkpoint is written."""
        if self.distributed:
            dist.broadcast(self.flat["bn"], src=src, group=self.pg)
            for n in _BNS:
                dist.broadcast(getattr(self.model, n).num_batches_tracked, src=src, group=self.pg)

    def optimizer_state_dict(self):
        """`torch.optim.Adam(model.parameters(), lr, weight_decay).state_dict()` of the fused optimizer (what pcs.py:376
        saves): per-parameter `step` / `exp_avg` / `exp_avg_sq` keyed by the position in `model.parameters()` (= the
        arena order) and one param group.  Tensors are clones."""
        ps = self.model._param_list()
        state = {}
        if self.step_count > 0:
            for i, (p, (o, n)) in enumerate(zip(ps, self.flat["offs"])):
                state[i] = {"step": torch.tensor(float(self.step_count)), "exp_avg": self.exp_avg[o:o + n].view(p.shape).clone(),
                            "exp_avg_sq": self.exp_avg_sq[o:o + n].view(p.shape).clone()}
        group = {"lr": self.lr, "betas": tuple(self.betas), "eps": self.eps, "weight_decay": self.weight_decay, "amsgrad": False,
                 "maximize": False, "foreach": None, "capturable": False, "differentiable": False, "fused": None,
                 "decoupled_weight_decay": False, "params": list(range(len(ps)))}
        return {"state": state, "param_groups": [group]}

    def load_optimizer_state(self, sd):
        """Restore the optimizer from `optimizer_state_dict()` or from a `torch.optim.Adam.state_dict()` of the reference
        loop (pcs.py:217, 376): moments, step count (host and device copies) and the hyper-parameters of the group."""
        group = sd["param_groups"][0]
        self.betas, self.eps, self.weight_decay = tuple(group["betas"]), group["eps"], group["weight_decay"]
        self.set_lr(group["lr"])
        ps = self.model._param_list()
        if len(group["params"]) != len(ps):
            raise ValueError(f"optimizer state has {len(group['params'])} parameters, the model has {len(ps)}")
        self.exp_avg.zero_()
        self.exp_avg_sq.zero_()
        step = 0
        for i, (p, (o, n)) in enumerate(zip(ps, self.flat["offs"])):
            st = sd["state"].get(group["params"][i])
            if st is None:
                continue
            self.exp_avg[o:o + n].copy_(st["exp_avg"].reshape(-1))
            self.exp_avg_sq[o:o + n].copy_(st["exp_avg_sq"].reshape(-1))
            step = max(step, int(float(st["step"])))
        self.step_count = step
        self.state.view(torch.int64)[1] = step
        self._graph, self._graph_key = None, None          # hyper-parameters may be baked into a captured step

    def checkpoint_dict(self, epoch=0, **extra):
        """Same dict as the reference writes to best_model.pth (pcs.py:373-382): `optimizer_state_dict` has the
        torch.optim.Adam layout (loadable by the reference's optimizer and by `load_optimizer_state`)."""
        d = {"epoch": epoch, "model_state_dict": {k: v.detach().clone() for k, v in self.model.state_dict().items()},
             "optimizer_state_dict": self.optimizer_state_dict(), "num_classes": self.model.num_classes}
        d.update(extra)
        return d
