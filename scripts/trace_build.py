"""OBVHS_TRACE=1 python scripts/trace_build.py <workload> [tris] -- per-stage wall times of build_cwbvh_from_tris on cuda:0."""
import os
import sys
import time

os.environ["OBVHS_TRACE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from obvhs_b200 import api  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "kitchen"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000_000
preset = sys.argv[3] if len(sys.argv) > 3 else None
tris, rays, desc, p = bench.make_workload(wl, n)
preset = preset or p
print(desc, preset, flush=True)
ctx = api.Context(0)
d_tris = torch.from_numpy(tris).cuda()
for it in range(3):
    print(f"--- build {it}", file=sys.stderr, flush=True)
    t0 = time.perf_counter()
    bvh = api.build_cwbvh_from_tris(d_tris, api.BvhBuildParams.preset(preset), ctx=ctx)
    print(f"--- total {1e3 * (time.perf_counter() - t0):.3f} ms, core {bvh.core_build_seconds * 1e3:.3f} ms", file=sys.stderr, flush=True)
