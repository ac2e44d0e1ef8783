"""Multi-GPU plumbing for the traversal path: one process per GPU, rays sharded, the finished CWBVH replicated.

The reference has no multi-device path at all (rayon threads inside one process). Each `CwBvh::ray_traverse` call only
reads `&self` (src/cwbvh/mod.rs:169), so rays shard trivially: the build runs on ONE GPU (PLOC iterations are globally
ordered), the finished tree (80-byte nodes, primitive_indices, permuted triangles) is broadcast once with NCCL over
NVLink/NVSwitch, and every rank traverses its own contiguous range of rays against its replica. There is no exchange step
during traversal, hence no collective on the traversal path.

`torch.distributed` is the plumbing (backend "nccl" on GPUs, "gloo" in the CPU tests).
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


def broadcast_cwbvh(bvh, ctx, src: int = 0):
    """Replicate a finished CwBvh (built on rank `src`; pass None elsewhere) to every rank's GPU.

    Three NCCL broadcasts straight out of / into the library's device buffers: nodes (80*M B), primitive_indices (4*N B),
    permuted triangles (64*N B, RtTriangle form). Returns this rank's CwBvh handle."""
    import torch.distributed as dist

    from . import api

    import torch

    rank = dist.get_rank()
    # sizes + scene AABB travel as one small NCCL broadcast (11 float64: exact for counts < 2^53 and for f32 bounds)
    meta = torch.zeros(11, dtype=torch.float64, device=f"cuda:{ctx.device}")
    if rank == src:
        nodes_p, prims_p, tris_p = bvh.device_ptrs()
        meta[:3] = torch.tensor([bvh.node_count, bvh.prim_count, 1.0 if tris_p else 0.0], dtype=torch.float64)
        meta[3:] = torch.from_numpy(bvh.total_aabb().astype(np.float64))
    dist.broadcast(meta, src=src)
    m = meta.tolist()
    node_count, prim_count, has_tris, total = int(m[0]), int(m[1]), bool(m[2]), m[3:]
    tri_bytes = int(ctx.lib.obvhs_cuda_cwbvh_triangle_bytes())  # the handle keeps 64-byte RtTriangles
    if rank != src:
        bvh = api.CwBvh.alloc(node_count, prim_count, has_tris, np.asarray(total, np.float32), ctx=ctx)
        nodes_p, prims_p, tris_p = bvh.device_ptrs()
    ctx.synchronize()
    for ptr, nbytes in ((nodes_p, node_count * 80), (prims_p, prim_count * 4), (tris_p if has_tris else 0, prim_count * tri_bytes)):
        if ptr and nbytes:
            dist.broadcast(device_bytes_tensor(ptr, nbytes, ctx.device), src=src)
    return bvh


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
