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


class PeerGather:
    """Gather of per-element results fused into the producing kernel (pgbp_comm_*, include/pgbp_b200.h).

    Every rank owns a window [nbuffers][world][ld] in HBM, exported with CUDA IPC and mapped by all peers;
    `integrate_gather(batch, j, k)` runs integratebelief! and stores each log-likelihood into row `rank` of buffer k
    on EVERY rank through NVLink peer stores -- no collective launch, nothing on the step's critical path.
    The 64-byte IPC handles are exchanged once, at construction, through `torch.distributed`
    (`handles=` takes them directly instead: tests, other launchers)."""

    def __init__(self, lib, device, rank, world, ld, nbuffers=2, group=None, handles=None, connect=True):
        import ctypes as C
        self.lib, self.rank, self.world, self.ld, self.nbuffers = lib, int(rank), int(world), int(ld), int(nbuffers)
        h = C.c_void_p()
        lib.check(lib.pgbp_comm_create(int(device), self.rank, self.world, self.ld, self.nbuffers, C.byref(h)))
        self.handle = h
        mine = (C.c_uint8 * 64)()
        lib.check(lib.pgbp_comm_handle(h, mine))
        self.ipc_handle = bytes(mine)
        if handles is None and connect:
            import torch
            import torch.distributed as dist
            if self.world == 1:
                handles = [self.ipc_handle]
            else:
                dev = torch.device("cuda", int(device)) if dist.get_backend(group) == "nccl" else torch.device("cpu")
                t = torch.tensor(list(self.ipc_handle), dtype=torch.uint8, device=dev)
                out = torch.empty(self.world * 64, dtype=torch.uint8, device=dev)
                dist.all_gather_into_tensor(out, t, group=group)
                flat = bytes(out.cpu().tolist())
                handles = [flat[64 * r:64 * (r + 1)] for r in range(self.world)]
        if handles is not None:
            self.connect(handles)

    def connect(self, handles):
        import ctypes as C
        buf = (C.c_uint8 * (64 * self.world)).from_buffer_copy(b"".join(handles))
        self.lib.check(self.lib.pgbp_comm_connect(self.handle, buf))

    def window_ptr(self, buffer):
        import ctypes as C
        p, ld = C.c_void_p(), C.c_int64()
        self.lib.check(self.lib.pgbp_comm_window(self.handle, buffer, C.byref(p), C.byref(ld)))
        return p.value, ld.value

    def integrate_gather(self, batch, j, buffer):
        """integratebelief!(beliefs, j) (1-based belief index) + store into every rank's window; enqueue only."""
        self.lib.check(self.lib.pgbp_integrate_gather(batch.handle, j - 1, self.handle, buffer))

    def put(self, batch, buffer, d_src_ptr):
        import ctypes as C
        self.lib.check(self.lib.pgbp_comm_put(self.handle, batch.handle, buffer, C.c_void_p(int(d_src_ptr))))

    def wait(self, batch, buffer, timeout_ms=2000):
        self.lib.check(self.lib.pgbp_comm_wait(self.handle, batch.handle, buffer, timeout_ms))

    def check(self, batch):
        self.lib.check(self.lib.pgbp_comm_check(self.handle, batch.handle))

    def read(self, batch, buffer):
        """Host copy [world, ld] of this rank's window (synchronous on the batch's stream)."""
        import ctypes as C
        out = np.empty((self.world, self.ld))
        self.lib.check(self.lib.pgbp_comm_read(self.handle, batch.handle, buffer, out.ctypes.data_as(C.POINTER(C.c_double))))
        return out

    def close(self, barrier=None):
        """Teardown: unmap the peers, `barrier()` (all ranks have unmapped), free the local window."""
        if getattr(self, "handle", None):
            self.lib.pgbp_comm_disconnect(self.handle)
            if barrier is not None:
                barrier()
            self.lib.pgbp_comm_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
