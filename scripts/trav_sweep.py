"""python scripts/trav_sweep.py <workload> [tris] [samples] -- device time of the persistent traversal kernel variants on one workload.
VARIANTS = comma list of "<variant>[/<refill>]" (obvhs_cuda_set_option traverse_variant / traverse); all must return
identical hits. PACKED=1 feeds the 32-byte Ray::new records instead of the 64-byte Ray structs."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from obvhs_b200 import api, camera  # noqa: E402
from obvhs_b200.types import ray_args_of  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "soup"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000_000
samples = int(sys.argv[3]) if len(sys.argv) > 3 else 8
variants = os.environ.get("VARIANTS", "0,1,2,3,4").split(",")
tris, rays, desc, preset = bench.make_workload(wl, n)
ctx0 = api.Context(0)
d_tris = torch.from_numpy(tris).cuda()
bvh = api.build_cwbvh_from_tris(d_tris, api.BvhBuildParams.preset(preset), ctx=ctx0)
nodes, prim_idx, total = bvh.download()
if rays is None:
    rays, _ = camera.demoscene_bounce_set(camera.demoscene_camera(1280), range(samples), tris[prim_idx], lambda r: bvh.ray_traverse(r))
print(desc, rays.shape[0], "rays", flush=True)
if os.environ.get("PACKED") == "1":
    rays = ray_args_of(rays)
d_rays = torch.from_numpy(rays).cuda()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ref = None
stream = torch.cuda.Stream()
ctx = api.Context(0, stream=stream.cuda_stream)
b = api.CwBvh.upload(nodes, prim_idx, total, ctx=ctx)
b.set_triangles(d_tris)
d_hits = torch.empty((rays.shape[0], 4), dtype=torch.int32, device="cuda")
for v in variants:
    vv, _, refill = v.partition("/")
    ctx.set_option("traverse", "persistent:" + (refill or "4"))
    ctx.set_option("traverse_variant", vv)
    ts = []
    for it in range(6):
        flush.zero_()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        b.ray_traverse(d_rays, out=d_hits)
        e1.record(stream)
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    h = d_hits.cpu().numpy()
    if ref is None:
        ref = h
    same = bool(np.array_equal(ref, h))
    t = float(np.median(ts[1:]))
    print(f"{v:12s} {t:8.3f} ms  {rays.shape[0] / t / 1e3:9.1f} Mrays/s  identical={same}", flush=True)
    assert same
