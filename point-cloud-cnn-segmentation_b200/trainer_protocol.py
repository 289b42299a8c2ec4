"""Backend-agnostic collective protocol of the data-parallel training step (no CUDA dependency, so the
CPU test-suite can exercise it with gloo)."""
import torch.distributed as dist


def grad_buckets(offs, total):
    """Two all-reduce buckets matching pcseg_backward's phases: ranges of the flat gradient arena that
    are final after phase 1 (global_feat + seg head + their BNs) and after phase 2 (the rest)."""
    conv_split = offs[10][0]          # first element of global_feat.weight
    bn_start = offs[20][0]            # first BN tensor
    bn_split = offs[30][0]            # bn_global.weight
    early = [(conv_split, bn_start), (bn_split, total)]
    late = [(0, conv_split), (bn_start, bn_split)]
    return early, late


class GradSync:
    """Collective protocol of the data-parallel step (backend-agnostic: NCCL on GPUs, gloo in the CPU tests).

    * The reference computes ONE weighted-mean loss over the gathered logits of all replicas (pcs.py:244-251 under
      nn.DataParallel): every gradient is divided by the GLOBAL sum of class weights.  Backward is linear in the loss
      scale, so the ranks back-propagate the un-normalised loss and the division happens in the optimizer:
      `launch_tensor(lw)` starts the asynchronous SUM all-reduce of {loss numerator, sum of class weights} right after the
      forward and nobody waits for it before the optimizer step (no collective on the critical path ahead of backward).
      `reduce_normaliser` is the blocking form (normalise before backward) kept for callers that want true gradients in
      the arena.
    * `launch(ranges)` / `wait()`: asynchronous SUM all-reduce of slices of the flat gradient arena (summing
      equals DataParallel's reduce-add of replica gradients); launched per bucket so that the first bucket's
      transfer overlaps the rest of backward.
    """

    def __init__(self, flat_grads, process_group=None):
        self.g = flat_grads
        self.pg = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.works = []
        # slices of a bucket go out as ONE collective launch on NCCL (ncclGroupStart/End); other backends (gloo in the CPU
        # tests) get one call per slice -- a coalescing attempt that fails half-way leaves the group unusable
        self._coalesce = (self.world > 1 and hasattr(dist, "_coalescing_manager")
                          and "nccl" in str(dist.get_backend(process_group)).lower())

    @property
    def active(self):
        return self.world > 1

    def reduce_normaliser(self, t):
        if self.active:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.pg)
        return t

    def launch_tensor(self, t):
        """asynchronous SUM all-reduce of a small tensor (joined by `wait`)"""
        if self.active:
            self.works.append(dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.pg, async_op=True))

    def launch(self, ranges):
        """one asynchronous SUM all-reduce of the given slices of the arena (the slices of a bucket are coalesced into one
        collective launch where the backend supports it)"""
        if not self.active:
            return
        views = [self.g[a:b] for a, b in ranges if b > a]
        if len(views) > 1 and self._coalesce:
            with dist._coalescing_manager(group=self.pg, device=self.g.device, async_ops=True) as cm:
                for v in views:
                    dist.all_reduce(v, op=dist.ReduceOp.SUM, group=self.pg)
            self.works.append(cm)
            return
        for v in views:
            self.works.append(dist.all_reduce(v, op=dist.ReduceOp.SUM, group=self.pg, async_op=True))

    def wait(self):
        for w in self.works:
            w.wait()
        self.works = []


