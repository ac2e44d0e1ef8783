"""python scripts/refit_modes_10m.py -- release- vs debug-build refit_from_fast (bvh2/mod.rs:722-751) on the 10 M-triangle scene, oracle only
(CPU, ~1 min): counts the Bvh2 nodes and CWBVH bytes whose BITS differ between the two modes (values never differ)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

import bench  # noqa: E402
import oracle_bind as ob  # noqa: E402

tris = bench.cached_demoscene(int(sys.argv[1]) if len(sys.argv) > 1 else 2237)
aabbs = ob.tri_aabbs(tris)
res = []
for full in (0, 1):
    ob.lib().orc_set_refit_full(full)
    b = ob.ploc_build(aabbs, None, 6, 64, 2, threads=os.cpu_count())
    applied = b.reinsertion_run(0.02, threads=os.cpu_count())
    res.append((applied, b.get()[0].copy(), b.to_cwbvh(3, True).get()[0].copy()))
ob.lib().orc_set_refit_full(0)
(a0, n0, c0), (a1, n1, c1) = res
assert a0 == a1 and np.array_equal(n0["aabb"], n1["aabb"]) and np.array_equal(n0["first_index"], n1["first_index"])
bits = n0["aabb"].view(np.uint32) != n1["aabb"].view(np.uint32)
d = c0.view(np.uint8).reshape(-1, 80) != c1.view(np.uint8).reshape(-1, 80)
print(f"{tris.shape[0]} tris, {a0} reinsertions: {int(bits.any(axis=1).sum())} of {n0.shape[0]} Bvh2 nodes and {int(d.sum())} bytes in "
      f"{int(d.any(axis=1).sum())} of {c0.shape[0]} CwBvh nodes differ in bits (zero signs only: {bool(np.all(n0['aabb'][bits] == 0.0))})")
