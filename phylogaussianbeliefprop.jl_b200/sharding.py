"""Batch sharding across ranks (one process per GPU).

The path shards by batch element only: rank r owns the contiguous slice
`shard_slice(B, r, world)` of the replicate / parameter axis, the plan is replicated, and nothing is
exchanged during message passing.  The single collective is the gather of per-element results
(log-likelihoods, energies, status) at the end of a step; it goes through `torch.distributed`
(NCCL on GPUs over NVLink; gloo in the CPU test of this module).
"""
from __future__ import annotations

import numpy as np


def shard_size(B: int, world: int) -> int:
    """Elements per rank: ceil(B / world) (the last ranks may own fewer, possibly zero)."""
    return (int(B) + world - 1) // world


def shard_slice(B: int, rank: int, world: int) -> slice:
    bg = shard_size(B, world)
    lo = min(rank * bg, B)
    return slice(lo, min(lo + bg, B))


def allgather_elements(local, B: int, group=None):
    """Gather per-element results of all ranks into the global element order.

    local: torch tensor whose LAST axis is this rank's elements (length shard_slice(B, rank, world)
    length); every rank gets the tensor over all B elements.  Pads to the common shard size so that
    one all_gather_into_tensor suffices (B_g * 8 bytes per rank and row: latency-bound)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    bg = shard_size(B, world)
    lead = tuple(local.shape[:-1])
    buf = torch.zeros(lead + (bg,), dtype=local.dtype, device=local.device)
    buf[..., :local.shape[-1]] = local
    flat = buf.contiguous().reshape(-1)
    out = torch.empty(world * flat.numel(), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, flat, group=group)
    # [world, ..., bg] -> [..., world*bg] -> trim
    out = out.reshape((world,) + lead + (bg,)).movedim(0, -2).reshape(lead + (world * bg,))
    return out[..., :B]


def shard_inputs(arrays, B: int, rank: int, world: int):
    """Slice the leading (element) axis of each array whose leading length is B; arrays with a
    leading length of 1 (shared parameters / shared data) are passed through."""
    sl = shard_slice(B, rank, world)
    out = []
    for a in arrays:
        a = np.asarray(a)
        out.append(a[sl] if a.shape[0] == B else a)
    return out
