"""Multi-GPU plumbing for the traversal path: one process per GPU, rays sharded, the finished CWBVH replicated.

The reference has no multi-device path at all (rayon threads inside one process). Each `CwBvh::ray_traverse` call only
reads `&self` (src/cwbvh/mod.rs:169), so rays shard trivially: the build runs on ONE GPU (PLOC iterations are globally
ordered), the finished tree (80-byte nodes, primitive_indices, permuted triangles) is broadcast once with NCCL over
NVLink/NVSwitch, and every rank traverses its own contiguous range of rays against its replica. There is no exchange step
during traversal, hence no collective on the traversal path.

`torch.distributed` launches the ranks and ships the communicator id (backend "nccl" on GPUs, "gloo" in the CPU tests); the tree
itself travels through the library's own NCCL communicator (include/obvhs_cuda.h: obvhs_cuda_comm_init / _cwbvh_broadcast).
"""
from __future__ import annotations

import numpy as np


def shard_range(n: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous range [lo, hi) of `n` rays owned by `rank`: rays[g*n/G .. (g+1)*n/G)."""
    assert 0 <= rank < world
    return (n * rank) // world, (n * (rank + 1)) // world


class _DevView:
    """A raw device allocation exposed through __cuda_array_interface__ so torch can wrap it without a copy."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


def device_bytes_tensor(ptr: int, nbytes: int, device: int):
    import torch

    return torch.as_tensor(_DevView(ptr, nbytes), device=f"cuda:{device}")


def broadcast_meta(meta, src: int = 0):
    """Broadcast a small picklable object (tree sizes, scene AABB) from `src` to every rank."""
    import torch.distributed as dist

    box = [meta]
    dist.broadcast_object_list(box, src=src)
    return box[0]


def init_comm(ctx):
    """Bind the library's own NCCL communicator to `ctx` (obvhs_cuda_comm_init). torch.distributed only ships the 128-byte id from
    rank 0; every later transfer runs below the C ABI, on the context's stream. Idempotent."""
    import torch.distributed as dist

    from . import api

    if getattr(ctx, "comm_world", None) is not None:
        return
    box = [api.nccl_unique_id() if dist.get_rank() == 0 else None]
    dist.broadcast_object_list(box, src=0)
    ctx.comm_init(box[0], dist.get_rank(), dist.get_world_size())


def broadcast_cwbvh(bvh, ctx, src: int = 0):
    """Replicate a finished CwBvh (built on rank `src`; pass None -- or an earlier replica to refill -- elsewhere) to every rank's
    GPU: obvhs_cuda_cwbvh_broadcast, i.e. a 64-byte header plus ONE grouped NCCL launch for nodes (80*M B), primitive_indices
    (4*N B) and the permuted triangles (64*N B, RtTriangle form), straight out of / into the library's device buffers and ordered
    on the context's stream (whatever follows on that stream sees the replica). Returns this rank's CwBvh handle."""
    from . import api

    init_comm(ctx)
    return api.CwBvh.broadcast(bvh, ctx, src)


def broadcast_arrays_cpu(arrays, src: int = 0):
    """Same replication for host numpy arrays (gloo): used by the CPU tests of the host-side sharding logic."""
    import torch
    import torch.distributed as dist

    rank = dist.get_rank()
    meta = [(a.shape, a.dtype.str) for a in arrays] if rank == src else None
    meta = broadcast_meta(meta, src)
    out = []
    for i, (shape, dt) in enumerate(meta):
        a = np.ascontiguousarray(arrays[i]) if rank == src else np.zeros(shape, dtype=np.dtype(dt))
        t = torch.from_numpy(a.view(np.uint8).reshape(-1)) if a.size else torch.zeros(0, dtype=torch.uint8)
        if a.size:
            dist.broadcast(t, src=src)
        out.append(a)
    return out
