"""Gradient all-reduce over NVLink peer memory (csrc/peer_allreduce.cuh): every rank maps the gradient arenas and signal
blocks of its peers through CUDA IPC and ONE kernel per step reduce-scatters / all-gathers the arena with peer loads; the
{loss numerator, sum of class weights} pair of the deferred loss normalisation rides along.  Replaces the NCCL all-reduce
between CUDA-graph segments (and DataParallel's reduce-add of pcs.py:209-211): the kernel runs on the compute stream, so the
whole data-parallel step is one CUDA graph.  One node only (2, 4 or 8 ranks)."""
import ctypes as C

import torch
import torch.distributed as dist

from ._lib import lib, check, ptr


def _export(t):
    h = C.create_string_buffer(64)
    off = C.c_longlong()
    check(lib.pcseg_ipc_export(ptr(t), h, C.byref(off)), "pcseg_ipc_export")
    return bytes(h.raw), int(off.value)


class PeerAllReduce:
    def __init__(self, arena, lw_in, lw_out, group=None):
        """arena: this rank's flat fp32 gradient tensor (storage padded to a multiple of 4 floats); lw_in / lw_out: 2-element
        fp64 device tensors (local / global {loss numerator, sum of class weights})."""
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        dev = arena.device
        n = (arena.numel() + 3) // 4 * 4
        if arena.untyped_storage().nbytes() < arena.storage_offset() * 4 + n * 4:
            raise ValueError("gradient arena storage is not padded to a multiple of 4 floats")
        self.signals = torch.zeros(int(lib.pcseg_peer_ar_signal_bytes(n, self.world)) // 4 + 64, dtype=torch.int32, device=dev)
        self.counters = torch.zeros(4, dtype=torch.int32, device=dev)
        self.keep = (arena, lw_in, lw_out)
        torch.cuda.synchronize(dev)
        mine = (_export(arena), _export(self.signals))
        everyone = [None] * self.world
        dist.all_gather_object(everyone, mine, group=group)
        self.handle = C.c_void_p()
        with torch.cuda.device(dev):
            check(lib.pcseg_peer_ar_create(C.byref(self.handle), self.rank, self.world, ptr(arena), n, ptr(self.signals), ptr(self.counters),
                                           ptr(lw_in), ptr(lw_out)), "pcseg_peer_ar_create")
            for p, ((ha, oa), (hs, os_)) in enumerate(everyone):
                if p != self.rank:
                    check(lib.pcseg_peer_ar_open(self.handle, p, ha, oa, hs, os_), "pcseg_peer_ar_open")
        dist.barrier(group=group)          # every rank has mapped every peer before the first kernel spins on a flag
        self.device = dev

    def run(self):
        """enqueue the all-reduce kernel on the current stream (capturable)"""
        with torch.cuda.device(self.device):
            check(lib.pcseg_peer_ar_run(self.handle, C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)), "pcseg_peer_ar_run")

    def __del__(self):
        try:
            if self.handle:
                lib.pcseg_peer_ar_destroy(self.handle)
        except Exception:
            pass
