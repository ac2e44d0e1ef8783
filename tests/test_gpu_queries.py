"""GPU parity for the batched broad-phase queries (Bvh2::aabb_traverse / point_traverse, and the traverse! macro over
CwBvhNode::intersect_aabb / contains_point) through the C ABI against the CPU oracle: per-query counts and the reported ids
in the reference's call order, bit for bit. Includes the reference's own traverse_aabb / traverse_point cases."""
import ctypes as C

import numpy as np
import pytest

import oracle_bind as ob
from obvhs_b200 import test_util as tu
from test_gpu_parity import api, oracle_refit_semantics  # noqa: F401  (fixtures)
from test_oracle_queries import brute_aabb, brute_point

pytestmark = pytest.mark.gpu


def query_boxes(tris, n, seed, size):
    lo = tris.reshape(-1, 4)[:, :3].min(axis=0)
    hi = tris.reshape(-1, 4)[:, :3].max(axis=0)
    rng = np.random.default_rng(seed)
    c = lo + rng.random((n, 3), dtype=np.float32) * (hi - lo)
    h = rng.random((n, 3), dtype=np.float32) * np.float32(size) * (hi - lo)
    q = np.zeros((n, 8), np.float32)
    q[:, 0:3] = c - h
    q[:, 4:7] = c + h
    q[0, 0:3], q[0, 4:7] = lo - 1, hi + 1          # everything
    q[1, 0:3], q[1, 4:7] = hi + 1, hi + 2          # nothing
    q[2, 0:3], q[2, 4:7] = c[2], c[2]              # a degenerate box
    q[3, 0:3], q[3, 4:7] = c[3] + 1, c[3] - 1      # an inverted box
    return q


def pair(api, tris, preset="fast_build"):
    wb = ob.build_bvh2_from_tris(tris, preset)
    wc = ob.build_cwbvh_from_tris(tris, preset)
    gb = api.Bvh2.upload(*wb.get(), max_depth=wb.max_depth)
    nodes, prims, total = wc.get()
    gc = api.CwBvh.upload(nodes, prims, total)
    return wb, wc, gb, gc


@pytest.mark.parametrize("scene", ["cornell", "terrain32", "soup4k", "kitchen"])
def test_aabb_and_point_queries_bit_exact(api, scenes, scene):
    tris = scenes[scene]
    wb, wc, gb, gc = pair(api, tris)
    q = query_boxes(tris, 3000, 1, 0.08)
    for w, g in ((wb, gb), (wc, gc)):
        wcounts, wids = w.aabb_traverse(q)
        gcounts, gids = g.aabb_traverse(q)
        assert np.array_equal(gcounts, wcounts) and np.array_equal(gids, wids)
        assert wcounts[1] == 0 and wcounts[0] > 0
    pts = (q[:, 0:3] + q[:, 4:7]) * np.float32(0.5)
    pts[:8] = tris[:8, 0:3]  # exactly on vertices: the closed interval tests of contains_point
    for w, g in ((wb, gb), (wc, gc)):
        wcounts, wids = w.point_traverse(pts)
        gcounts, gids = g.point_traverse(pts)
        assert np.array_equal(gcounts, wcounts) and np.array_equal(gids, wids)
    # child order follows the traversal direction (CwBvh::new_traversal)
    for d in ((1.0, -1.0, 1.0), (-1.0, -1.0, -1.0), (-0.0, 2.0, -3.0)):
        wcounts, wids = wc.aabb_traverse(q, d)
        gcounts, gids = gc.aabb_traverse(q, d)
        assert np.array_equal(gcounts, wcounts) and np.array_equal(gids, wids), d


def test_reference_traverse_aabb_on_gpu(api):
    # tests/mod.rs:178-249
    tris = tu.demoscene(201, 0)
    aabbs = ob.tri_aabbs(tris)
    q = np.array([[0.511, -1.0, 0.511, 0, 0.611, 1.0, 0.611, 0]], np.float32)
    ref = brute_aabb(aabbs, q[0])
    bvh2 = api.build_bvh2_from_tris(tris, api.BvhBuildParams.fast_build())
    nodes, prims = bvh2.download()
    _, leaves = bvh2.aabb_traverse(q)
    cand = np.concatenate([prims[nodes["first_index"][l]: nodes["first_index"][l] + nodes["prim_count"][l]] for l in leaves]).astype(np.int64)
    found = cand[np.isin(cand, ref)]
    assert found.size == ref.size and int(found.sum()) == int(ref.sum())
    cw = api.build_cwbvh_from_tris(tris, api.BvhBuildParams.fast_build())
    _, cprims, _ = cw.download()
    _, slots = cw.aabb_traverse(q)
    cand = cprims[slots].astype(np.int64)
    found = cand[np.isin(cand, ref)]
    assert found.size == ref.size and int(found.sum()) == int(ref.sum())


def test_reference_traverse_point_on_gpu(api):
    # tests/mod.rs:251-325
    tris = tu.icosphere(0)
    aabbs = ob.tri_aabbs(tris)
    i = np.arange(512, dtype=np.uint32)
    z = np.zeros(512, np.uint32)
    pts = tu.uniform_sample_sphere(tu.hash_noise(z, z, i), tu.hash_noise(z, z + np.uint32(1), i))
    cw = api.build_cwbvh_from_tris(tris, api.BvhBuildParams.fast_build())
    bvh2 = api.build_bvh2_from_tris(tris, api.BvhBuildParams.fast_build())
    _, cprims, _ = cw.download()
    nodes, prims = bvh2.download()
    ccounts, cslots = cw.point_traverse(pts)
    bcounts, bleaves = bvh2.point_traverse(pts)
    co = np.concatenate([[0], np.cumsum(ccounts, dtype=np.int64)]).astype(np.int64)
    bo = np.concatenate([[0], np.cumsum(bcounts, dtype=np.int64)]).astype(np.int64)
    for k in range(512):
        ref = brute_point(aabbs, pts[k])
        cand = cprims[cslots[co[k]:co[k + 1]]].astype(np.int64)
        found = cand[np.isin(cand, ref)]
        assert found.size == ref.size and int(found.sum()) == int(ref.sum()), k
        ls = bleaves[bo[k]:bo[k + 1]]
        cand = np.concatenate([prims[nodes["first_index"][l]: nodes["first_index"][l] + nodes["prim_count"][l]] for l in ls] + [np.zeros(0, np.uint32)]).astype(np.int64)
        found = cand[np.isin(cand, ref)]
        assert found.size == ref.size and int(found.sum()) == int(ref.sum()), k


def test_query_degenerate_trees_capacity_and_device_buffers(api):
    import torch

    ctx = api.default_context()
    q = np.array([[-1, -1, -1, 0, 1, 1, 1, 0], [5, 5, 5, 0, 6, 6, 6, 0]], np.float32)
    empty = api.PlocBuilder().build(1, np.zeros((0, 8), np.float32))
    assert list(empty.aabb_traverse(q)[0]) == [0, 0]
    assert list(api.bvh2_to_cwbvh(empty, 3).point_traverse(np.zeros((2, 3), np.float32))[0]) == [0, 0]
    one = api.PlocBuilder().build(1, ob.tri_aabbs(tu.plane()[:1]))  # the root is a leaf
    counts, leaves = one.aabb_traverse(q)
    assert list(counts) == [1, 0] and list(leaves) == [0]
    counts, slots = api.bvh2_to_cwbvh(one, 3).aabb_traverse(q)
    assert list(counts) == [1, 0] and list(slots) == [0]
    # capacity too small: nothing written, total reported
    tris = tu.triangle_soup(2000, 1)
    bvh = api.PlocBuilder().build_tris(6, tris)
    qs = query_boxes(tris, 64, 2, 0.2)
    want_counts, want_ids = bvh.aabb_traverse(qs)
    ids = np.full(8, 0xABCDEF, np.uint32)
    counts = np.zeros(64, np.uint32)
    total = C.c_size_t(0)
    rc = ctx.lib.obvhs_cuda_bvh2_aabb_traverse_batch(ctx.h, bvh.h, api._ptr(qs), 64, api._ptr(counts), api._ptr(ids), 8, C.byref(total))
    assert rc == api.ERR_CAPACITY and total.value == want_ids.size and np.all(ids == 0xABCDEF) and np.array_equal(counts, want_counts)
    # everything resident on the device
    d_q = torch.from_numpy(qs).cuda()
    d_counts = torch.zeros(64, dtype=torch.int32, device="cuda")
    d_ids = torch.zeros(want_ids.size, dtype=torch.int32, device="cuda")
    ctx.check(ctx.lib.obvhs_cuda_bvh2_aabb_traverse_batch(ctx.h, bvh.h, api._ptr(d_q), 64, api._ptr(d_counts), api._ptr(d_ids), want_ids.size,
                                                          C.byref(total)))
    assert np.array_equal(d_counts.cpu().numpy().view(np.uint32), want_counts) and np.array_equal(d_ids.cpu().numpy().view(np.uint32), want_ids)


def test_collision_broad_phase_like_the_physics_example(api):
    # examples/physics.rs:566-588: every item queries the tree with its own box; pairs = overlapping boxes
    tris = tu.triangle_soup(20000, 6)
    aabbs = ob.tri_aabbs(tris)
    bvh = api.PlocBuilder().build(1, aabbs)
    nodes, prims = bvh.download()
    counts, leaves = bvh.aabb_traverse(aabbs)
    owner = np.repeat(np.arange(len(aabbs)), counts)
    other = prims[nodes["first_index"][leaves]]  # single-primitive leaves
    got = set(zip(owner[owner != other].tolist(), other[owner != other].tolist()))
    # brute force on a sample of the items
    for s1 in range(0, len(aabbs), 997):
        ref = brute_aabb(aabbs, aabbs[s1])
        assert {(s1, int(s2)) for s2 in ref if s2 != s1} == {p for p in got if p[0] == s1}


@pytest.mark.parametrize("n,g", [(300, 0.05), (2000, 0.005)])
def test_deep_bvh2_traversal_and_queries(api, n, g):
    # max_depth 156 (the 192-entry fixed stack) and ~1000 (beyond every fixed stack: the reference's HeapStack, here a global arena):
    # Bvh2 ray traversal (closest / miss / all-hit count, both kernels) and the box / point queries against the oracle
    from test_gpu_parity import graded_boxes
    from obvhs_b200.types import make_rays

    a = graded_boxes(n, g)
    tris = np.zeros((n, 12), np.float32)  # one thin triangle per box, same AABB
    tris[:, 0:3] = a[:, 0:3]
    tris[:, 4:7] = np.stack([a[:, 4], a[:, 1], a[:, 2]], axis=1)
    tris[:, 8:11] = np.stack([a[:, 0], a[:, 5], a[:, 6]], axis=1)
    aabbs = ob.tri_aabbs(tris)
    w = ob.ploc_build(aabbs, None, 6, 64, 2)
    assert w.max_depth > (192 if n == 2000 else 96)
    wn, wp = w.get()
    gb = api.Bvh2.upload(wn, wp, max_depth=w.max_depth)
    gb.set_triangles(tris)
    bt = w.bvh_tris(tris)
    rng = np.random.default_rng(5)
    m = 40000
    x = (aabbs[rng.integers(0, n, m), 0] + aabbs[rng.integers(0, n, m), 4]) * np.float32(0.5)
    o = np.stack([x, np.full(m, 0.5, np.float32), rng.random(m, dtype=np.float32) * np.float32(0.02) - np.float32(0.01)], axis=1).astype(np.float32)
    d = np.stack([rng.random(m, dtype=np.float32) * np.float32(0.4) - np.float32(0.2), np.full(m, -1.0, np.float32),
                  rng.random(m, dtype=np.float32) * np.float32(0.02) - np.float32(0.01)], axis=1).astype(np.float32)
    rays = make_rays(o, d / np.linalg.norm(d, axis=1, keepdims=True).astype(np.float32), 0.0, np.inf)
    rays[m // 2:, 0:3] = np.stack([np.full(m - m // 2, -1.0, np.float32), np.zeros(m - m // 2, np.float32), np.zeros(m - m // 2, np.float32)], axis=1)
    rays[m // 2:, 4:7] = np.array([1.0, 0.0, 0.0], np.float32)  # along the whole chain: deep stacks
    rays[m // 2:, 8:11] = np.array([1.0, 8388608.0, 8388608.0], np.float32)
    want = w.ray_traverse(bt, rays)
    assert (want["t"] < np.inf).sum() > 100
    for mode in ("static", "persistent"):
        ctx = api.Context(0, traverse=mode)
        g2 = api.Bvh2.upload(wn, wp, max_depth=w.max_depth, ctx=ctx)
        g2.set_triangles(tris)
        got = g2.ray_traverse(rays)
        assert np.array_equal(got["primitive_id"], want["primitive_id"]) and np.array_equal(got["t"].view(np.uint32), want["t"].view(np.uint32))
        assert np.array_equal(g2.ray_traverse_miss(rays), w.ray_traverse_miss(bt, rays))
        assert np.array_equal(g2.ray_traverse_anyhit_count(rays), w.ray_traverse_anyhit_count(bt, rays))
    q = query_boxes(tris, 2000, 2, 0.2)
    wc, wi = w.aabb_traverse(q)
    gc, gi = gb.aabb_traverse(q)
    assert np.array_equal(gc, wc) and np.array_equal(gi, wi) and wc[0] == n
    pts = (q[:, 0:3] + q[:, 4:7]) * np.float32(0.5)
    wc, wi = w.point_traverse(pts)
    gc, gi = gb.point_traverse(pts)
    assert np.array_equal(gc, wc) and np.array_equal(gi, wi)
