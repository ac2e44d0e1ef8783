"""Two ranks on two GPUs: the tree built on rank 0 is replicated with obvhs_cuda_cwbvh_broadcast (NCCL below the C ABI), every
replica is byte-compared with the oracle's tree, and the concatenated sharded hits with the oracle's hits. Skipped below 2 GPUs."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, tmp):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch
    import torch.distributed as dist

    from obvhs_b200 import api, sharding, test_util as tu
    from obvhs_b200.types import make_rays

    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{rank}"))
    ctx = api.Context(rank)
    tris = np.concatenate([tu.demoscene(48, 0), tu.triangle_soup(4096, 3)], axis=0)
    rng = np.random.default_rng(0)
    o = rng.random((20011, 3), dtype=np.float32) * 2 - 1 + np.array([0, 2, 0], np.float32)
    d = rng.random((20011, 3), dtype=np.float32) - np.array([0.5, 1.5, 0.5], np.float32)
    rays = make_rays(o, d / np.linalg.norm(d, axis=1, keepdims=True).astype(np.float32), 0.0, np.inf)
    replica = None
    for preset in ("fast_build", "medium_build"):  # the second round refills / replaces the first replica
        bvh = api.build_cwbvh_from_tris(tris, api.BvhBuildParams.preset(preset), ctx=ctx) if rank == 0 else replica
        replica = sharding.broadcast_cwbvh(bvh, ctx, src=0)
        nodes, prims, total = replica.download()
        lo, hi = sharding.shard_range(rays.shape[0], rank, world)
        hits = replica.ray_traverse(rays[lo:hi])
        np.save(os.path.join(tmp, f"nodes_{preset}_{rank}.npy"), nodes.view(np.uint8))
        np.save(os.path.join(tmp, f"prims_{preset}_{rank}.npy"), prims)
        np.save(os.path.join(tmp, f"total_{preset}_{rank}.npy"), total)
        np.save(os.path.join(tmp, f"hits_{preset}_{rank}.npy"), hits)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.gpu
def test_two_gpu_broadcast_replica_and_sharded_hits(tmp_path):
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    import oracle_bind as ob
    from obvhs_b200 import test_util as tu
    from obvhs_b200.types import RAY_HIT, make_rays

    tris = np.concatenate([tu.demoscene(48, 0), tu.triangle_soup(4096, 3)], axis=0)
    rng = np.random.default_rng(0)
    o = rng.random((20011, 3), dtype=np.float32) * 2 - 1 + np.array([0, 2, 0], np.float32)
    d = rng.random((20011, 3), dtype=np.float32) - np.array([0.5, 1.5, 0.5], np.float32)
    rays = make_rays(o, d / np.linalg.norm(d, axis=1, keepdims=True).astype(np.float32), 0.0, np.inf)
    for preset in ("fast_build", "medium_build"):
        c = ob.build_cwbvh_from_tris(tris, preset)
        wn, wp, wt = c.get()
        for r in range(2):
            assert np.load(tmp_path / f"nodes_{preset}_{r}.npy").tobytes() == wn.tobytes(), f"replica on rank {r} differs from the oracle tree"
            assert np.array_equal(np.load(tmp_path / f"prims_{preset}_{r}.npy"), wp)
            assert np.array_equal(np.load(tmp_path / f"total_{preset}_{r}.npy")[[0, 1, 2, 4, 5, 6]], np.asarray(wt, np.float32)[[0, 1, 2, 4, 5, 6]])
        want = c.ray_traverse(c.bvh_tris(tris), rays)
        got = np.concatenate([np.load(tmp_path / f"hits_{preset}_{r}.npy") for r in range(2)]).view(RAY_HIT).reshape(-1)
        assert np.array_equal(got["primitive_id"], want["primitive_id"])
        assert np.array_equal(got["t"].view(np.uint32), want["t"].view(np.uint32))
        assert (want["t"] < np.inf).sum() > 1000
