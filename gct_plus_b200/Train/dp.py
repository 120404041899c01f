"""Host-side data-parallel helpers (one process per GPU, torch.distributed for the plumbing).

Training follows train1.py:111-112 (DDP): every rank computes the gradient of ITS sum-loss, gradients are
summed over ranks and divided by the world size (the division is folded into the fused Adam kernel).
Sampling shards independent latent draws with no data-path collective.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch
import torch.distributed as dist


def world_info(group=None):
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def shard_range(n_total: int, rank: int, world: int):
    """Contiguous share [lo, hi) of n_total independent work items for `rank` (sizes differ by at most 1)."""
    base, rem = divmod(n_total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def rank_seed(base_seed: int, rank: int) -> int:
    """Distinct, reproducible RNG stream per rank for the latent draws."""
    return (int(base_seed) * 1000003 + 7919 * int(rank)) % (2 ** 31 - 1)


def allreduce_sum_(flat: torch.Tensor, group=None) -> float:
    """In-place sum over ranks of the flat gradient buffer; returns the scale (1/world) that turns it into DDP's mean."""
    _, world = world_info(group)
    if world > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    return 1.0 / world


def gradient_buckets(offsets: np.ndarray, n_layers: int, total: int, num_global_slots: int, enc_slots: int, dec_slots: int):
    """Cuts the flat gradient buffer into the buckets of the overlapped exchange: one per decoder layer (final after backward
    stage N-1-l), one per encoder layer (stage 2N-1-l), and the rest -- embeddings, latent heads, final norms, the
    vocabulary projection, parameters the kernels never touch -- at stage 2N.  Returns [(stage, offset, count)] covering
    [0, total) exactly once.  Relies on the layout of engine._flatten: every layer's slots are contiguous and layers follow
    each other in order (encoder 0..N-1, then decoder 0..N-1)."""
    N = n_layers
    enc0 = [int(offsets[num_global_slots + l * enc_slots]) for l in range(N)]
    dec0 = [int(offsets[num_global_slots + N * enc_slots + l * dec_slots]) for l in range(N)]
    # end of the last decoder layer = end of its last slot (ff.linear_2.bias, d_model elements, padded to the 64-element slot
    # alignment of engine._flatten); whatever follows (parameters no kernel reads, e.g. vaetf's encoder.fc_mu) is "the rest"
    last = int(offsets[num_global_slots + N * enc_slots + N * dec_slots - 1])
    d_model = int(offsets[num_global_slots + 1] - offsets[num_global_slots])      # layer 0: norm_1.alpha -> norm_1.bias
    dec_end = min(total, last + d_model)
    assert all(int(o) < dec_end for o in offsets if o >= 0), "a slot is laid out behind the decoder layers"
    starts = enc0 + dec0 + [dec_end]
    assert starts == sorted(starts), "layer slots are not laid out in order"
    out = []
    for l in range(N):
        out.append((2 * N - 1 - l, enc0[l], starts[l + 1] - enc0[l]))
    for l in range(N):
        out.append((N - 1 - l, dec0[l], starts[N + l + 1] - dec0[l]))
    if enc0[0] > 0:
        out.append((2 * N, 0, enc0[0]))
    if total > dec_end:
        out.append((2 * N, dec_end, total - dec_end))
    assert sum(c for _, _, c in out) == total
    return out


class GradExchange:
    """NCCL communicator owned by the trainer + the bucket table of gct_backward_dp.  Rendezvous: rank 0 draws the 128-byte
    NCCL id, torch.distributed (whatever backend the host initialised) broadcasts it, every rank calls ncclCommInitRank
    through the library.  One communicator per trainer, destroyed with it."""

    def __init__(self, model, group=None):
        from .. import _lib as L
        self.L = L
        lib = L.lib()
        rank, world = world_info(group)
        self.rank, self.world = rank, world
        ident = (C.c_char * 128)()
        if rank == 0:
            L.check(lib.gct_nccl_unique_id(C.addressof(ident)), "gct_nccl_unique_id")
        if world > 1:
            box = [bytes(ident)]
            dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
            ident = (C.c_char * 128).from_buffer_copy(box[0])
        self.comm = C.c_void_p()
        L.check(lib.gct_nccl_comm_init(C.byref(self.comm), world, rank, C.addressof(ident)), "gct_nccl_comm_init")
        self.stream = torch.cuda.Stream(device=model._flat.device)
        cfg = model._cfg()
        bl = gradient_buckets(model._offsets, cfg.n_layers, model._flat.numel(), L.NUM_GLOBAL_SLOTS, L.ENC_LAYER_SLOTS, L.DEC_LAYER_SLOTS)
        self.buckets = (L.GctBucket * len(bl))(*[L.GctBucket(stage=s, reserved=0, offset=o, count=c) for s, o, c in bl])
        self.n_buckets = len(bl)

    def allreduce_(self, flat: torch.Tensor):
        """Un-overlapped exchange of the whole buffer on the current stream (the A/B baseline of the bucketed form)."""
        self.L.check(self.L.lib().gct_allreduce_grads(self.comm, self.L.ptr(flat), flat.numel(), self.L.stream_ptr()), "gct_allreduce_grads")

    def close(self):
        if self.comm:
            self.L.lib().gct_nccl_comm_destroy(self.comm)
            self.comm = C.c_void_p()

    def __del__(self):
        import sys
        if sys.is_finalizing():      # the CUDA context / NCCL may already be gone at interpreter shutdown: leave the handle to the OS
            return
        try:
            self.close()
        except Exception:
            pass
