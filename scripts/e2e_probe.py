"""Where does a host-to-host traversal call spend its time? (B200 box; tuning aid, prints ms per variant)"""
import sys, time, os
import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from obvhs_b200 import api
from obvhs_b200.types import RAY_HIT, ray_args_of

wl = sys.argv[1] if len(sys.argv) > 1 else "kitchen"
tris, rays, desc, preset = bench.make_workload(wl, None)
ctx = api.Context(0)
bvh = api.build_cwbvh_from_tris(tris, api.BvhBuildParams.preset(preset), ctx=ctx)
n = rays.shape[0]
h_args = torch.from_numpy(ray_args_of(rays)).pin_memory()
h_rays = torch.from_numpy(rays).pin_memory()
h_hits = torch.empty((n, 4), dtype=torch.int32).pin_memory()
hits_np = h_hits.numpy().view(RAY_HIT).reshape(-1)
d_args = h_args.cuda()
d_rays = h_rays.cuda()
d_hits = torch.empty((n, 4), dtype=torch.int32, device="cuda")


def t(name, fn, reps=8):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / reps * 1e3
    print(f"{name:46s} {ms:8.3f} ms  {n / ms / 1e3:8.0f} Mrays/s")


def sync(fn):
    def g():
        fn()
        ctx.synchronize()
    return g


print(desc, n, "rays")
t("host args -> host hits", lambda: bvh.ray_traverse(h_args.numpy(), out=hits_np))
t("host args -> device hits", sync(lambda: bvh.ray_traverse(h_args.numpy(), out=d_hits)))
t("device args -> host hits", lambda: bvh.ray_traverse(d_args, out=hits_np))
t("device args -> device hits", sync(lambda: bvh.ray_traverse(d_args, out=d_hits)))
t("device rays -> device hits", sync(lambda: bvh.ray_traverse(d_rays, out=d_hits)))
t("host rays -> host hits", lambda: bvh.ray_traverse(h_rays.numpy(), out=hits_np))
t("memcpy H2D args (32 B/ray)", lambda: d_args.copy_(h_args, non_blocking=True))
t("memcpy D2H hits (16 B/ray)", lambda: h_hits.copy_(d_hits, non_blocking=True))
s2 = torch.cuda.Stream()


def both():
    d_args.copy_(h_args, non_blocking=True)
    with torch.cuda.stream(s2):
        h_hits.copy_(d_hits, non_blocking=True)


t("memcpy H2D args || D2H hits", both)
for sl in [int(x) for x in os.environ.get('SLICES', '32768,65536,131072,262144,524288').split(',')]:
    ctx.set_option("host_slice", str(sl))
    t(f"host args -> host hits, host_slice {sl}", lambda: bvh.ray_traverse(h_args.numpy(), out=hits_np))
