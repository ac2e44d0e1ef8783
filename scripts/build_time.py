"""python scripts/build_time.py <workload> [tris] -- device time (CUDA events) and core_build_time of build_cwbvh_from_tris, no tracing."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import bench
from obvhs_b200 import api

wl = sys.argv[1] if len(sys.argv) > 1 else "terrain"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000_000
tris, _, desc, preset = bench.make_workload(wl, n)
stream = torch.cuda.Stream()
ctx = api.Context(0, stream=stream.cuda_stream)
d_tris = torch.from_numpy(tris).cuda()
params = api.BvhBuildParams.preset(preset)
ev, core = [], []
for it in range(13):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    bvh = api.build_cwbvh_from_tris(d_tris, params, ctx=ctx)
    b.record(stream)
    torch.cuda.synchronize()
    if it >= 3:
        ev.append(a.elapsed_time(b)); core.append(bvh.core_build_seconds * 1e3)
print(f"{desc}: events {np.mean(ev):.3f} ms (min {np.min(ev):.3f}), core_build_time {np.mean(core):.3f} ms")
