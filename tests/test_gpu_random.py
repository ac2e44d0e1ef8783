"""Randomised end-to-end parity (GPU path vs the CPU oracle through the C ABI): scene generators chosen to provoke the tie rules
the reference leaves to its unstable sorts (duplicated triangles = equal Morton codes AND equal merge costs, coplanar grids =
equal surface areas everywhere, a huge coordinate offset = few distinct Morton cells, degenerate zero-area triangles), random
sizes and every preset, for both tree types. Bit-exact nodes, index maps and hits."""
import os

import numpy as np
import pytest

import oracle_bind as ob
from obvhs_b200 import test_util as tu
from obvhs_b200.types import make_rays

from test_gpu_parity import api, assert_nodes_equal, oracle_refit_semantics  # noqa: F401  (fixtures)

pytestmark = pytest.mark.gpu

PRESET_NAMES = list(ob.PRESETS)


def random_scene(seed):
    rng = np.random.default_rng(1000 + seed)
    n = int(np.exp(rng.uniform(np.log(30), np.log(40000))))
    kind = seed % 6
    if kind == 0:  # every triangle four times: equal keys, equal AABBs, equal merge costs
        base = tu.triangle_soup(max(1, n // 4), seed)
        tris = np.concatenate([base] * 4, axis=0)[rng.permutation(4 * base.shape[0])]
    elif kind == 1:  # coplanar regular grid: identical areas and extents
        tris = tu.flat_plane(max(1, int(np.sqrt(n / 2))))
    elif kind == 2:  # a small soup far from the origin: f32 spacing collapses many centres onto few Morton cells
        tris = tu.triangle_soup(n, seed) * np.float32(0.01)
        tris[:, [0, 1, 2, 4, 5, 6, 8, 9, 10]] += np.float32(4096.0)
    elif kind == 3:  # degenerate triangles (points and segments) mixed into a soup
        tris = tu.triangle_soup(n, seed)
        k = rng.random(tris.shape[0]) < 0.3
        tris[k, 4:7] = tris[k, 0:3]
        k2 = rng.random(tris.shape[0]) < 0.15
        tris[k2, 8:11] = tris[k2, 0:3]
    elif kind == 4:  # soup with long slivers (pre-splits do real work in the slow presets)
        tris = tu.soup_with_large_triangles(n, max(1, n // 200), seed)
    else:  # terrain
        tris = tu.demoscene(max(2, int(np.sqrt(n / 2))), seed)
    return np.ascontiguousarray(tris, dtype=np.float32)


def random_rays(tris, seed, m=4000):
    rng = np.random.default_rng(77 + seed)
    v = tris.reshape(-1, 4)[:, :3]
    lo, hi = v.min(axis=0), v.max(axis=0)
    c, ext = (lo + hi) * 0.5, max(float(np.max(hi - lo)), 1e-3)
    o = (c + (rng.random((m, 3)) - 0.5) * ext * 2.5).astype(np.float32)
    t = (c + (rng.random((m, 3)) - 0.5) * ext).astype(np.float32)
    d = t - o
    d /= np.maximum(np.linalg.norm(d, axis=1, keepdims=True), 1e-20)
    return make_rays(o, d.astype(np.float32), 0.0, np.inf)


N_SEEDS = int(os.environ.get("OBVHS_RANDOM_SEEDS", "24"))  # a one-off `OBVHS_RANDOM_SEEDS=400` run is a cheap fuzzing pass


@pytest.mark.parametrize("seed", range(N_SEEDS))
def test_random_scene_cwbvh_and_bvh2_end_to_end(api, seed):
    tris = random_scene(seed)
    preset = PRESET_NAMES[(seed * 5 + seed // 6) % len(PRESET_NAMES)]
    rays = random_rays(tris, seed)
    # CwBvh
    want = ob.build_cwbvh_from_tris(tris, preset)
    wn, wp, _ = want.get()
    got = api.build_cwbvh_from_tris(tris, api.BvhBuildParams.preset(preset))
    gn, gp, _ = got.download()
    assert gn.shape == wn.shape, (preset, tris.shape[0])
    assert np.array_equal(gp, wp), (preset, tris.shape[0])
    assert gn.tobytes() == wn.tobytes(), (preset, tris.shape[0])
    wh = want.ray_traverse(want.bvh_tris(tris), rays)
    gh = got.ray_traverse(rays)
    assert np.array_equal(gh["primitive_id"], wh["primitive_id"])
    assert np.array_equal(gh["t"].view(np.uint32), wh["t"].view(np.uint32))
    # Bvh2 (collapsed leaves, second reinsertion pass)
    want2 = ob.build_bvh2_from_tris(tris, preset)
    wn2, wp2 = want2.get()
    got2 = api.build_bvh2_from_tris(tris, api.BvhBuildParams.preset(preset))
    gn2, gp2 = got2.download()
    assert np.array_equal(gp2, wp2), (preset, tris.shape[0])
    assert_nodes_equal(gn2, wn2, f"seed {seed} {preset} {tris.shape[0]} tris")
    wh2 = want2.ray_traverse(want2.bvh_tris(tris), rays)
    gh2 = got2.ray_traverse(rays)
    assert np.array_equal(gh2["primitive_id"], wh2["primitive_id"])
    assert np.array_equal(gh2["t"].view(np.uint32), wh2["t"].view(np.uint32))


@pytest.mark.parametrize("seed", range(N_SEEDS))
def test_random_scene_rebuilds_queries_and_dynamic_updates(api, seed):
    # the remaining stages on the same scenes: PLOC with a random search distance / precision / depth threshold, reinsertion
    # at a random ratio, a partial rebuild of a random leaf set and a full rebuild (ploc/rebuild.rs), box and point queries on
    # both tree types, then two frames of moved leaves (refit_all + reinsertion, examples/physics.rs)
    from test_gpu_parity import displaced_aabbs
    from test_gpu_queries import query_boxes
    from test_oracle_rebuild import leaf_ids

    rng = np.random.default_rng(5000 + seed)
    tris = random_scene(seed)
    aabbs = ob.tri_aabbs(tris)
    sd = int(rng.choice([1, 2, 6, 14, 24, 32]))
    prec = int(rng.choice([64, 128]))
    thr = int(rng.integers(0, 4))
    ratio = float(rng.choice([0.01, 0.1, 0.5, 1.0]))
    what = f"seed {seed} n {tris.shape[0]} sd {sd} prec {prec} thr {thr} ratio {ratio}"
    want = ob.ploc_build(aabbs, None, sd, prec, thr)
    got = api.PlocBuilder().build(sd, aabbs, None, api.SortPrecision(prec), thr)
    assert_nodes_equal(got.download()[0], want.get()[0], what + " ploc")
    assert api.ReinsertionOptimizer().run(got, ratio) == want.reinsertion_run(ratio), what
    assert_nodes_equal(got.download()[0], want.get()[0], what + " reinsertion")
    # queries on the Bvh2 and on the CwBvh made from it
    q = query_boxes(tris, 300, seed, 0.1)
    pts = (q[:, 0:3] + q[:, 4:7]) * np.float32(0.5)
    wc = want.to_cwbvh(int(rng.integers(1, 4)), bool(rng.integers(0, 2)))
    gc = api.CwBvh.upload(*wc.get())
    for w, g in ((want, got), (wc, gc)):
        for fn, arg in (("aabb_traverse", q), ("point_traverse", pts)):
            wcounts, wids = getattr(w, fn)(arg)
            gcounts, gids = getattr(g, fn)(arg)
            assert np.array_equal(gcounts, wcounts) and np.array_equal(gids, wids), what + " " + fn
    # partial rebuild of a random set of leaves, then a full rebuild
    want.compute_parents()
    got.compute_parents()
    ids = leaf_ids(want)
    ids = ids[rng.random(len(ids)) < rng.choice([0.02, 0.3, 1.0])]
    if len(ids):
        wflags = want.rebuild_path_flags(ids)
        assert np.array_equal(api.compute_rebuild_path_flags(got, ids), wflags), what
        want.partial_rebuild(wflags, sd, prec, thr)
        api.PlocBuilder().partial_rebuild(got, wflags, sd, api.SortPrecision(prec), thr)
        assert_nodes_equal(got.download()[0], want.get()[0], what + " partial_rebuild")
        assert np.array_equal(got.download()[1], want.get()[1]), what
    want.full_rebuild(sd, prec, thr)
    api.PlocBuilder().full_rebuild(got, sd, api.SortPrecision(prec), thr)
    assert_nodes_equal(got.download()[0], want.get()[0], what + " full_rebuild")
    # moved leaves
    for frame in range(2):
        moved = displaced_aabbs(tris, frame + seed)
        want.set_leaf_aabbs(moved)
        want.refit_all()
        got.set_leaf_aabbs(moved)
        assert_nodes_equal(got.download()[0], want.get()[0], what + f" frame {frame} refit")
        assert api.ReinsertionOptimizer().run(got, 0.05) == want.reinsertion_run(0.05), what
        assert_nodes_equal(got.download()[0], want.get()[0], what + f" frame {frame} reinsertion")


@pytest.mark.parametrize("seed", range(N_SEEDS))
def test_random_scene_traversal_flavours_and_kernel_variants(api, seed):
    # closest hit / miss / all-hit counts and the visit counters, through a random kernel variant (one ray per thread, persistent
    # warps at a random refill threshold and fetch size, or the per-block choice), for rays with random tmin / tmax windows,
    # axis-parallel directions and origins inside the scene; CwBvh and Bvh2
    rng = np.random.default_rng(9000 + seed)
    tris = random_scene(seed)
    rays = random_rays(tris, seed, m=3000)
    v = tris.reshape(-1, 4)[:, :3]
    ext = max(float(np.max(v.max(axis=0) - v.min(axis=0))), 1e-3)
    k = rng.random(rays.shape[0])
    rays[k < 0.3, 12] = (rng.random(int((k < 0.3).sum())) * ext).astype(np.float32)           # tmin > 0
    rays[k > 0.6, 13] = (rng.random(int((k > 0.6).sum())) * 2.0 * ext).astype(np.float32)     # finite tmax (some below tmin)
    ax = rng.integers(0, rays.shape[0], 200)
    d = np.zeros((200, 3), np.float32)
    d[np.arange(200), rng.integers(0, 3, 200)] = rng.choice(np.array([-1.0, 1.0], np.float32), 200)
    rays[ax] = make_rays(rays[ax, 0:3], d, 0.0, np.inf)
    mode = ["static", "auto", f"persistent:{rng.choice([1, 4, 8, 16, 32])}:{rng.choice([32, 64, 128])}"][seed % 3]
    ctx = api.Context(0, traverse=mode)
    preset = PRESET_NAMES[seed % 4]
    wc = ob.build_cwbvh_from_tris(tris, preset)
    gc = api.CwBvh.upload(*wc.get(), ctx=ctx)
    gc.set_triangles(tris)
    wb = ob.build_bvh2_from_tris(tris, preset)
    gb = api.Bvh2.upload(*wb.get(), max_depth=wb.max_depth, ctx=ctx)
    gb.set_triangles(tris)
    for w, g in ((wc, gc), (wb, gb)):
        bt = w.bvh_tris(tris)
        cw, cg = np.zeros(2, np.uint64), np.zeros(2, np.uint64)
        want = w.ray_traverse(bt, rays, counters=cw)
        got = g.ray_traverse(rays, counters=cg)
        assert np.array_equal(got["primitive_id"], want["primitive_id"]), (seed, mode)
        assert np.array_equal(got["t"].view(np.uint32), want["t"].view(np.uint32)), (seed, mode)
        assert np.array_equal(cg, cw), (seed, mode)
        assert np.array_equal(g.ray_traverse_miss(rays), w.ray_traverse_miss(bt, rays)), (seed, mode)
        assert np.array_equal(g.ray_traverse_anyhit_count(rays), w.ray_traverse_anyhit_count(bt, rays)), (seed, mode)
