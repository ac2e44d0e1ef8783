"""obvhs_b200.device_rays (torch, meant for the GPU) against the numpy recipes of obvhs_b200.camera, on the CPU."""
import numpy as np
import torch

from obvhs_b200 import camera, device_rays, test_util as tu
from obvhs_b200.types import ray_args_of


def test_hash_noise_matches_numpy():
    x = np.arange(0, 5000, dtype=np.uint32)
    y = (x * np.uint32(7) + np.uint32(3)) % np.uint32(475)
    for frame in (0, 1, 164, 512 + 17, 1024 + 164):
        want = tu.hash_noise(x, y, np.uint32(frame))
        got = device_rays.hash_noise(torch.from_numpy(x.astype(np.int64)), torch.from_numpy(y.astype(np.int64)), frame).numpy()
        assert np.array_equal(want.view(np.uint32), got.view(np.uint32))


def test_primary_rays_match_numpy():
    cam = camera.demoscene_camera(160)
    dcam = device_rays.DeviceCamera(cam, torch.device("cpu"))
    for s in (0, 3):
        want = ray_args_of(camera.demoscene_primary(cam, s))
        got = device_rays.demoscene_primary(dcam, s).numpy()
        assert got.shape == want.shape
        assert np.array_equal(got[:, [3, 7]], want[:, [3, 7]])  # tmin, tmax
        np.testing.assert_allclose(got[:, 0:3], want[:, 0:3], rtol=0, atol=1e-7)
        np.testing.assert_allclose(got[:, 4:7], want[:, 4:7], rtol=0, atol=2e-6)


def test_bounce_rays_match_numpy():
    import oracle_bind as ob

    tris = tu.demoscene(48, 0)
    c = ob.build_cwbvh_from_tris(tris, "fast_build")
    bt = c.bvh_tris(tris)
    cam = camera.demoscene_camera(96)
    prim = camera.demoscene_primary(cam, 1)
    hits = c.ray_traverse(bt, prim)
    nrm = camera.shading_normals(bt, hits["primitive_id"], prim[:, 4:7])
    want, _ = camera.diffuse_bounce_rays(prim, hits["t"], nrm, cam, 1)
    want = ray_args_of(want)
    # the device form reads the tree's RtTriangle records {v0, e1, e2, ng}
    t = bt.reshape(-1, 12)
    v0, v1, v2 = t[:, 0:3], t[:, 4:7], t[:, 8:11]
    e1, e2 = v0 - v1, v2 - v0
    rt = np.zeros((t.shape[0], 16), np.float32)
    rt[:, 0:3], rt[:, 4:7], rt[:, 8:11] = v0, e1, e2
    rt[:, 12:15] = np.stack([e1[:, 1] * e2[:, 2] - e2[:, 1] * e1[:, 2], e1[:, 2] * e2[:, 0] - e2[:, 2] * e1[:, 0], e1[:, 0] * e2[:, 1] - e2[:, 0] * e1[:, 1]], axis=1)
    dcam = device_rays.DeviceCamera(cam, torch.device("cpu"))
    h = torch.from_numpy(np.ascontiguousarray(hits).view(np.int32).reshape(-1, 4))
    got = device_rays.diffuse_bounce(dcam, 1, torch.from_numpy(ray_args_of(prim)), h, torch.from_numpy(rt)).numpy()
    assert got.shape == want.shape and got.shape[0] > 1000
    np.testing.assert_allclose(got[:, 0:3], want[:, 0:3], rtol=0, atol=1e-6)
    np.testing.assert_allclose(got[:, 4:7], want[:, 4:7], rtol=0, atol=1e-5)
