"""Broad-phase queries in the CPU oracle, pinned by the reference's own tests `traverse_aabb` and `traverse_point`
(tests/mod.rs:178-297): the primitives found through Bvh2::aabb_traverse / point_traverse and through the traverse! macro
over CwBvhNode::intersect_aabb / contains_point equal a brute-force scan (count and index sum). CPU only."""
import numpy as np

import oracle_bind as ob
from obvhs_b200 import test_util as tu


def brute_aabb(aabbs, q):
    return np.nonzero(np.all(q[0:3] <= aabbs[:, 4:7], axis=1) & np.all(q[4:7] >= aabbs[:, 0:3], axis=1))[0]


def brute_point(aabbs, p):
    return np.nonzero(np.all(p[0:3] >= aabbs[:, 0:3], axis=1) & np.all(p[0:3] <= aabbs[:, 4:7], axis=1))[0]


def bvh2_prims(bvh, leaf_ids):
    nodes, prims = bvh.get()
    out = []
    for leaf in leaf_ids:
        f, c = int(nodes["first_index"][leaf]), int(nodes["prim_count"][leaf])
        out.extend(prims[f:f + c])
    return np.asarray(out, dtype=np.int64)


def test_traverse_aabb_reference_test():
    # tests/mod.rs:178-249
    tris = tu.demoscene(201, 0)
    aabbs = ob.tri_aabbs(tris)
    q = np.array([[0.511, -1.0, 0.511, 0, 0.611, 1.0, 0.611, 0]], np.float32)
    ref = brute_aabb(aabbs, q[0])
    assert ref.size > 0
    bvh2 = ob.build_bvh2_from_tris(tris, "fast_build", threads=4)
    counts, leaves = bvh2.aabb_traverse(q)
    cand = bvh2_prims(bvh2, leaves)
    found = cand[np.isin(cand, ref)]  # the test's eval re-checks every primitive of a reported leaf
    assert found.size == ref.size and int(found.sum()) == int(ref.sum())
    assert counts[0] == leaves.size
    cw = ob.build_cwbvh_from_tris(tris, "fast_build", threads=4)
    _, prims, _ = cw.get()
    ccounts, slots = cw.aabb_traverse(q)  # new_traversal(Vec3A::ZERO)
    cand = prims[slots].astype(np.int64)
    found = cand[np.isin(cand, ref)]
    assert found.size == ref.size and int(found.sum()) == int(ref.sum())
    assert np.unique(slots).size == slots.size


def test_traverse_point_reference_test():
    # tests/mod.rs:251-325: 512 points on the unit sphere against icosphere(0)
    tris = tu.icosphere(0)
    aabbs = ob.tri_aabbs(tris)
    i = np.arange(512, dtype=np.uint32)
    z = np.zeros(512, np.uint32)
    pts = tu.uniform_sample_sphere(tu.hash_noise(z, z, i), tu.hash_noise(z, z + np.uint32(1), i))
    cw = ob.build_cwbvh_from_tris(tris, "fast_build")
    bvh2 = ob.build_bvh2_from_tris(tris, "fast_build")
    _, prims, _ = cw.get()
    ccounts, cslots = cw.point_traverse(pts)
    bcounts, bleaves = bvh2.point_traverse(pts)
    co = np.concatenate([[0], np.cumsum(ccounts, dtype=np.int64)]).astype(np.int64)
    bo = np.concatenate([[0], np.cumsum(bcounts, dtype=np.int64)]).astype(np.int64)
    some = 0
    for k in range(512):
        ref = brute_point(aabbs, pts[k])
        some += ref.size
        cand = prims[cslots[co[k]:co[k + 1]]].astype(np.int64)
        found = cand[np.isin(cand, ref)]
        assert found.size == ref.size and int(found.sum()) == int(ref.sum()), k
        cand = bvh2_prims(bvh2, bleaves[bo[k]:bo[k + 1]])
        found = cand[np.isin(cand, ref)]
        assert found.size == ref.size and int(found.sum()) == int(ref.sum()), k
    assert some > 512


def test_queries_on_degenerate_trees():
    q = np.array([[-1, -1, -1, 0, 1, 1, 1, 0]], np.float32)
    empty = ob.ploc_build(np.zeros((0, 8), np.float32), None, 1, 64, 0)
    assert empty.aabb_traverse(q)[0][0] == 0 and empty.point_traverse(np.zeros((1, 3)))[0][0] == 0
    assert empty.to_cwbvh(3).aabb_traverse(q)[0][0] == 0
    one = ob.ploc_build(ob.tri_aabbs(tu.plane()[:1]), None, 1, 64, 0)  # the root is a leaf (bvh2/mod.rs:370-376)
    counts, leaves = one.aabb_traverse(np.concatenate([q, q + np.float32(100.0)]))
    assert list(counts) == [1, 0] and list(leaves) == [0]
    counts, slots = one.to_cwbvh(3).aabb_traverse(q)
    assert list(counts) == [1] and list(slots) == [0]


def test_query_order_depends_on_traversal_direction_only_as_a_permutation():
    tris = tu.triangle_soup(3000, 4)
    cw = ob.build_cwbvh_from_tris(tris, "fast_build")
    q = np.array([[0.2, 0.2, 0.2, 0, 0.6, 0.5, 0.7, 0]], np.float32)
    a = cw.aabb_traverse(q, (1.0, 1.0, 1.0))[1]
    b = cw.aabb_traverse(q, (-1.0, -1.0, -1.0))[1]
    assert a.size > 10 and not np.array_equal(a, b) and np.array_equal(np.sort(a), np.sort(b))
