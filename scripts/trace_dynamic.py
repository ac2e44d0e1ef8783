"""OBVHS_TRACE=1 python scripts/trace_dynamic.py [tris] -- per-phase times of one dynamic frame (set_leaf_aabbs + refit_all + run(0.01))."""
import os, sys
os.environ["OBVHS_TRACE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import bench
from obvhs_b200 import api, test_util as tu

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
res = int(round((n / 2) ** 0.5))
tris = tu.demoscene(res, 0)
ctx = api.Context(0)
d_tris = torch.from_numpy(tris).cuda()
d_frames = bench.dynamic_frames_device(d_tris, 4)
aabbs0 = d_frames[0]
bvh = api.PlocBuilder(ctx).build(api.PlocSearchDistance.Low if hasattr(api.PlocSearchDistance, "Low") else 6, aabbs0, None, 64, 2)
opt = api.ReinsertionOptimizer()
opt.run(bvh, 0.02)
for it in range(3):
    print(f"--- frame {it}", file=sys.stderr, flush=True)
    bvh.set_leaf_aabbs(d_frames[(it + 1) % 4])
    opt.run(bvh, 0.01)
