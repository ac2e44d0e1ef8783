"""Host-side multi-GPU logic on CPU: world_size-2 gloo processes shard rays and replicate a tree (no GPU)."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_partitions_exactly():
    from obvhs_b200.sharding import shard_range

    for n in (0, 1, 7, 100, 2073600, 100_320_000):
        for world in (1, 2, 4, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, tmp):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist

    import oracle_bind as ob
    from obvhs_b200 import sharding, test_util as tu
    from obvhs_b200.types import make_rays

    dist.init_process_group("gloo", rank=rank, world_size=world)
    tris = tu.demoscene(16, 0)
    rng = np.random.default_rng(0)
    o = rng.random((1001, 3), dtype=np.float32) * 2 - 1 + np.array([0, 2, 0], np.float32)
    d = np.tile(np.array([[0.0, -1.0, 0.0]], np.float32), (1001, 1))
    rays = make_rays(o, d, 0.0, np.inf)
    # rank 0 "builds" (CPU oracle stands in for the GPU builder: this test is about the plumbing), everyone receives
    arrays = None
    if rank == 0:
        c = ob.build_cwbvh_from_tris(tris, "fast_build")
        nodes, prims, total = c.get()
        arrays = [nodes.view(np.uint8).reshape(-1), prims, c.bvh_tris(tris)]
    nodes_b, prims, bt = sharding.broadcast_arrays_cpu(arrays, src=0)
    from obvhs_b200.types import CWBVH_NODE

    c = ob.cwbvh_from(nodes_b.view(CWBVH_NODE), prims)
    lo, hi = sharding.shard_range(rays.shape[0], rank, world)
    hits = c.ray_traverse(bt, rays[lo:hi], threads=1)
    np.save(os.path.join(tmp, f"hits_{rank}.npy"), hits)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_broadcast_and_sharded_traversal(tmp_path):
    import socket

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_bind as ob
    from obvhs_b200 import test_util as tu
    from obvhs_b200.types import make_rays

    tris = tu.demoscene(16, 0)
    rng = np.random.default_rng(0)
    o = rng.random((1001, 3), dtype=np.float32) * 2 - 1 + np.array([0, 2, 0], np.float32)
    d = np.tile(np.array([[0.0, -1.0, 0.0]], np.float32), (1001, 1))
    rays = make_rays(o, d, 0.0, np.inf)
    c = ob.build_cwbvh_from_tris(tris, "fast_build")
    want = c.ray_traverse(c.bvh_tris(tris), rays, threads=1)
    got = np.concatenate([np.load(tmp_path / f"hits_{r}.npy") for r in range(2)])
    assert np.array_equal(got["primitive_id"], want["primitive_id"])
    assert np.array_equal(got["t"].view(np.uint32), want["t"].view(np.uint32))
    assert (want["t"] < np.inf).sum() > 500
