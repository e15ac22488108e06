"""Multi-GPU layer: batch sharding + the one collective of the path.

Sequences are independent (each batch element owns its map and pose chain), so a (B_total, L) job shards by
contiguous blocks of B_total / world sequences per rank, one process per GPU, and NO traffic crosses GPUs
while the L frames are fused.  The only exchange is at the end: an all-gather of the per-sequence sizes
followed by a variable-length all-gather of the fused maps (points 3 + normals 3 + colours 3 + confidence 1
floats per point), after which every rank holds all B_total maps.
"""
from typing import Optional

import torch
import torch.distributed as dist

from .structures.pointclouds import Pointclouds, _ATTRS

__all__ = ["shard_batch", "gather_maps"]


def shard_batch(total: int, rank: Optional[int] = None, world: Optional[int] = None):
    """Contiguous block [lo, hi) of the batch owned by `rank` (blocks differ by at most one element)."""
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    base, extra = divmod(total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_maps(pointclouds: Pointclouds, group=None) -> Pointclouds:
    """All-gathers the maps of every rank (equal local batch size).  Returns a Pointclouds with world*B maps,
    ordered by rank.  Works on NCCL (CUDA tensors) and on gloo (CPU tensors, used by the CPU tests)."""
    world = dist.get_world_size(group)
    if world == 1:
        return pointclouds
    dev = pointclouds.device
    B = len(pointclouds)
    local = pointclouds._counts_dev[pointclouds._cur].to(torch.int64)
    all_counts = torch.empty(world * B, dtype=torch.int64, device=dev)
    _all_gather(all_counts, local.contiguous(), group)
    counts = [int(c) for c in all_counts.tolist()]  # the one host sync of the whole job
    nmax = max(max(counts), 1)
    pointclouds._counts_host = counts[dist.get_rank(group) * B: (dist.get_rank(group) + 1) * B]
    pointclouds.reserve(nmax)
    pointclouds._zero_rows_upto(nmax)
    out = Pointclouds(device=dev)
    out._B = world * B
    for key in _ATTRS:
        st = pointclouds._store[key]
        if st is None:
            continue
        send = st[:, :nmax].contiguous()
        recv = torch.empty((world * B, nmax, st.shape[2]), dtype=st.dtype, device=dev)
        _all_gather(recv, send, group)
        out._store[key] = recv
    out._set_counts(counts)
    return out


def _all_gather(recv, send, group):
    if send.is_cuda:
        dist.all_gather_into_tensor(recv, send, group=group)
    else:
        chunks = list(recv.chunk(dist.get_world_size(group), dim=0))
        dist.all_gather(chunks, send, group=group)
