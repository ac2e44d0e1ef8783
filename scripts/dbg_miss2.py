import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oracle_bind as ob
from obvhs_b200 import api, test_util as tu
import test_gpu_parity as T
tris = np.ascontiguousarray(np.concatenate([tu.icosphere(1), tu.plane()], axis=0), dtype=np.float32)
c = ob.build_cwbvh_from_tris(tris, "fast_build")
nodes, prims, total = c.get()
bt = c.bvh_tris(tris)
rays = T.rays_for(tris)
want = c.ray_traverse(bt, rays)
srays = rays.copy()
finite = np.isfinite(want["t"])
srays[:, 13] = np.where(finite, want["t"] * np.float32(0.999), np.float32(5.0))
srays[::3, 13] = np.float32(1e30)
wm = c.ray_traverse_miss(bt, srays)
wc = c.ray_traverse_anyhit_count(bt, srays)
for rep in range(int(sys.argv[1])):
    for mode in (sys.argv[2:] or ["persistent:32:32", "persistent:8:32", "persistent:1:32", "persistent:32:64", "persistent:16:128", "auto"]):
        g = api.CwBvh.upload(nodes, prims, total, ctx=api.Context(0, traverse=mode))
        g.set_triangles(tris)
        t0 = time.time()
        print(mode, "closest", flush=True)
        a = g.ray_traverse(rays)
        print(mode, "miss", flush=True)
        b = g.ray_traverse_miss(srays)
        print(mode, "count", flush=True)
        d = g.ray_traverse_anyhit_count(srays)
        dt = time.time() - t0
        ok = np.array_equal(a["primitive_id"], want["primitive_id"]) and np.array_equal(b, wm) and np.array_equal(d, wc)
        if not ok or dt > 0.5:
            print("rep", rep, mode, "ok", ok, "dt", dt, flush=True)
print("finished", flush=True)
