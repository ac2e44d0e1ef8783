"""GPU parity for the Bvh2 side of the path, through the C ABI, against the CPU oracle:
leaf collapse (nodes + primitive_indices bit-exact), build_bvh2_from_tris, and Bvh2 ray traversal (closest / miss / any-hit
count, every kernel variant; hit ids, hit t bits and the visit counters identical)."""
import numpy as np
import pytest

import oracle_bind as ob
from obvhs_b200 import camera, test_util as tu
from obvhs_b200.types import make_rays

from test_gpu_parity import TRAVERSE_MODES, assert_nodes_equal, rays_for

pytestmark = pytest.mark.gpu
F32_MAX = np.float32(3.4028235e38)


@pytest.fixture(scope="module", autouse=True)
def oracle_refit_semantics():
    ob.lib().orc_set_refit_full(1)  # see tests/test_gpu_parity.py
    yield
    ob.lib().orc_set_refit_full(0)


@pytest.fixture(scope="module")
def api():
    from obvhs_b200 import api as a

    a.default_context(0)
    return a


@pytest.mark.parametrize("scene", ["cornell", "ico_plane", "terrain32", "soup4k", "kitchen"])
@pytest.mark.parametrize("max_prims,cost,with_parents", [(8, 3.0, False), (8, 1.0, True), (2, 0.0, True), (255, 3.0, False), (1, 3.0, True)])
def test_collapse_bit_exact(api, scenes, scene, max_prims, cost, with_parents):
    aabbs = ob.tri_aabbs(scenes[scene])
    want = ob.ploc_build(aabbs, None, 6, 64, 2)
    wn, wp = want.get()
    got = api.Bvh2.upload(wn, wp, want.max_depth, True)
    if with_parents:
        want.compute_parents()
        got.compute_parents()
    want.collapse(max_prims, cost)
    got.collapse(max_prims, cost)
    wn2, wp2 = want.get()
    gn2, gp2 = got.download()
    assert gn2.shape == wn2.shape
    assert_nodes_equal(gn2, wn2, f"{scene} collapse({max_prims},{cost})")
    assert np.array_equal(gp2, wp2)
    if with_parents:
        _, _, wpar = want.get(with_parents=True)
        _, _, gpar = got.download(with_parents=True)
        assert np.array_equal(gpar[1:], wpar[1:])
    rc, msg = ob.bvh2_from(gn2, gp2, want.max_depth).validate(aabbs, tight_fit=False)
    assert rc == 0, msg


@pytest.mark.parametrize("preset", ["fastest_build", "very_fast_build", "fast_build", "medium_build"])
@pytest.mark.parametrize("scene", ["cornell", "terrain32", "kitchen"])
def test_build_bvh2_from_tris_end_to_end(api, scenes, scene, preset):
    tris = scenes[scene]
    want = ob.build_bvh2_from_tris(tris, preset)
    got = api.build_bvh2_from_tris(tris, api.BvhBuildParams.preset(preset))
    wn, wp = want.get()
    gn, gp = got.download()
    assert_nodes_equal(gn, wn, f"{scene} {preset}")
    assert np.array_equal(gp, wp)
    rays = rays_for(tris, n_side=64)
    wh = want.ray_traverse(want.bvh_tris(tris), rays)
    gh = got.ray_traverse(rays)
    assert np.array_equal(gh["primitive_id"], wh["primitive_id"])
    assert np.array_equal(gh["t"].view(np.uint32), wh["t"].view(np.uint32))
    assert got.core_build_seconds > 0


@pytest.mark.parametrize("mode", TRAVERSE_MODES)
@pytest.mark.parametrize("scene", ["cornell", "ico_plane", "flat4", "terrain32", "soup4k", "kitchen"])
def test_bvh2_traversal_bit_exact(api, scenes, scene, mode):
    tris = scenes[scene]
    want = ob.build_bvh2_from_tris(tris, "medium_build")
    wn, wp = want.get()
    bt = want.bvh_tris(tris)
    rays = rays_for(tris)
    wc = np.zeros(2, np.uint64)
    wh = want.ray_traverse(bt, rays, counters=wc)
    g = api.Bvh2.upload(wn, wp, want.max_depth, False, ctx=api.Context(0, traverse=mode))
    g.set_triangles(tris)
    gc = np.zeros(2, np.uint64)
    gh = g.ray_traverse(rays, counters=gc)
    assert np.array_equal(gh["primitive_id"], wh["primitive_id"])
    assert np.array_equal(gh["t"].view(np.uint32), wh["t"].view(np.uint32))  # includes the `hit.t = ray.tmax` no-hit convention
    assert np.array_equal(gh["geometry_id"], wh["geometry_id"]) and np.array_equal(gh["instance_id"], wh["instance_id"])
    assert np.array_equal(gc, wc), "node tests / triangle tests differ: the visit order is not the reference's"
    assert np.array_equal(g.ray_traverse(rays)["t"].view(np.uint32), wh["t"].view(np.uint32))  # uncounted kernel variant
    srays = rays.copy()
    hit = wh["t"] < F32_MAX
    srays[:, 13] = np.where(hit, wh["t"] * np.float32(0.999), np.float32(5.0))
    srays[::3, 13] = np.float32(1e30)
    assert np.array_equal(g.ray_traverse_miss(srays), want.ray_traverse_miss(bt, srays))
    assert np.array_equal(g.ray_traverse_anyhit_count(srays), want.ray_traverse_anyhit_count(bt, srays))


def test_kitchen_golden_hash_through_gpu_bvh2(api, kitchen_tris):
    # examples/obj_cwbvh.rs:142-181 rendered through build_bvh2_from_tris + Bvh2::ray_traverse on the GPU
    rays = camera.primary_rays(camera.kitchen_camera(32))
    for preset in ("fast_build", "medium_build"):
        bvh = api.build_bvh2_from_tris(kitchen_tris, api.BvhBuildParams.preset(preset))
        hits = bvh.ray_traverse(rays)
        _, prims = bvh.download()
        with np.errstate(invalid="ignore"):
            nrm = ob.triangle_normals(kitchen_tris[prims])
        hit = hits["t"] < F32_MAX
        out = np.zeros((rays.shape[0], 3), np.float32)
        nn = nrm[hits["primitive_id"][hit]]
        d = rays[hit, 4:7]
        s = np.sign((nn[:, 0] * -d[:, 0] + nn[:, 1] * -d[:, 1]) + nn[:, 2] * -d[:, 2]).astype(np.float32)
        out[hit] = nn * s[:, None]
        assert tu.hash_vec3a_vec(out) == 1343358762


def test_bvh2_degenerate_and_errors(api):
    from obvhs_b200.types import make_rays

    ray = make_rays(np.array([[0.0, 0.0, 1.0]], np.float32), np.array([[0.0, 0.0, -1.0]], np.float32), 0.0, np.inf)
    empty = api.build_bvh2_from_tris(np.zeros((0, 12), np.float32), api.BvhBuildParams.fast_build())
    assert empty.node_count == 0
    assert not (empty.ray_traverse(ray)["t"][0] < np.inf)
    one = tu.flat_plane(4)[:1]
    b = api.build_bvh2_from_tris(one, api.BvhBuildParams.medium_build())
    w = ob.build_bvh2_from_tris(one, "medium_build")
    down = make_rays(np.array([[0.1, 1.0, 0.1]], np.float32), np.array([[0.0, -1.0, 0.0]], np.float32), 0.0, np.inf)
    assert np.array_equal(b.ray_traverse(down)["t"].view(np.uint32), w.ray_traverse(w.bvh_tris(one), down)["t"].view(np.uint32))
    # traversal without attached triangles is an error, not a crash
    aabbs = ob.tri_aabbs(tu.flat_plane(4))
    bare = api.PlocBuilder().build(api.PlocSearchDistance.Low, aabbs)
    with pytest.raises(api.ObvhsError):
        bare.ray_traverse(ray)


def test_large_collapse_and_bvh2_traversal_properties(api):
    # 0.5M-triangle terrain: collapse and the whole Bvh2 build bit-exact vs the oracle, traversal agrees with the CwBvh
    tris = tu.demoscene(500, 0)
    want = ob.build_bvh2_from_tris(tris, "fast_build", threads=ob.lib().orc_max_threads())
    got = api.build_bvh2_from_tris(tris, api.BvhBuildParams.fast_build())
    wn, wp = want.get()
    gn, gp = got.download()
    assert_nodes_equal(gn, wn, "terrain 0.5M build_bvh2_from_tris")
    assert np.array_equal(gp, wp)
    rays = camera.demoscene_primary(camera.demoscene_camera(640), 0)
    hb = got.ray_traverse(rays)
    cw = api.build_cwbvh_from_tris(tris, api.BvhBuildParams.fast_build())
    hc = cw.ray_traverse(rays)
    hit = hc["t"] < np.inf
    assert hit.sum() > 1000
    assert np.array_equal(hb["t"][hit].view(np.uint32), hc["t"][hit].view(np.uint32))
    assert np.array_equal(np.isinf(hb["t"]), ~hit)


PRESET_NAMES = ["fastest_build", "very_fast_build", "fast_build", "medium_build", "slow_build", "very_slow_build"]


@pytest.mark.parametrize("preset", PRESET_NAMES)
def test_reference_degenerate_builds_over_aabbs(api, preset):
    # tests/mod.rs:35-88: build_bvh2 / build_cwbvh over one empty AABB, ten infinite, ten LARGEST, and nothing; a ray must never
    # report a hit (the test's closure returns +inf for every primitive: here triangles that cannot be hit)
    inf, mx = np.float32(np.inf), np.float32(3.4028235e38)
    cases = {
        "empty": np.array([[mx, mx, mx, 0, -mx, -mx, -mx, 0]], np.float32),
        "inf": np.tile(np.array([[-inf, -inf, -inf, 0, inf, inf, inf, 0]], np.float32), (10, 1)),
        "max": np.tile(np.array([[-mx, -mx, -mx, 0, mx, mx, mx, 0]], np.float32), (10, 1)),
        "nothing": np.zeros((0, 8), np.float32),
    }
    ray = make_rays(np.array([[0.0, 0.0, 1.0]], np.float32), np.array([[0.0, 0.0, -1.0]], np.float32), 0.0, np.inf)
    p = api.BvhBuildParams.preset(preset)
    sd, thr, ratio, mult, prec, mp, cost = ob.BVH2_PRESETS[preset][:7]
    for name, aabbs in cases.items():
        far = np.zeros((max(1, aabbs.shape[0]), 12), np.float32)
        far[:, [0, 4, 8]] = 1e30
        b = api.build_bvh2(aabbs, p)
        want = ob.ploc_build(aabbs, None, sd, prec, thr)
        want.reinsertion_run(ratio)
        want.collapse(mp, cost)
        want.reinsertion_run(ratio * mult)
        gn, gp = b.download()
        wn, wp = want.get()
        assert np.array_equal(gp, wp), (name, preset)
        assert_nodes_equal(gn, wn, f"build_bvh2 {name} {preset}")
        if aabbs.shape[0]:
            b.set_triangles(far)
        assert not (b.ray_traverse(ray)["t"][0] < np.inf), name
        if name in ("inf", "max"):
            # the reference only builds CwBvhs over the empty box and over nothing (tests/mod.rs:62-88); with infinite boxes its
            # child ordering panics (bvh2_to_cwbvh.rs order_children, NaN centres). The library reports an error instead.
            with pytest.raises(api.ObvhsError):
                api.build_cwbvh(aabbs, p)
            continue
        c = api.build_cwbvh(aabbs, p)
        wc = ob.ploc_build(aabbs, None, sd, prec, thr)
        wc.reinsertion_run(ratio)
        wc = wc.to_cwbvh(min(max(mp, 1), 3))
        assert c.download()[0].tobytes() == wc.get()[0].tobytes(), (name, preset)
        if aabbs.shape[0]:
            c.set_triangles(far)
        assert not (c.ray_traverse(ray)["t"][0] < np.inf), name


def test_build_over_aabbs_equals_component_calls(api, scenes):
    tris = scenes["kitchen"]
    aabbs = ob.tri_aabbs(tris)
    for preset in ("fast_build", "medium_build"):
        p = api.BvhBuildParams.preset(preset)
        t = [0.0]
        a = api.build_cwbvh(aabbs, p, t).download()
        b = api.build_cwbvh_from_tris(tris, p).download()
        assert t[0] > 0 and a[0].tobytes() == b[0].tobytes() and np.array_equal(a[1], b[1])
        a = api.build_bvh2(aabbs, p).download()
        b = api.build_bvh2_from_tris(tris, p).download()
        assert_nodes_equal(a[0], b[0], preset)


@pytest.mark.parametrize("scene", ["terrain32", "soup4k", "kitchen"])
@pytest.mark.parametrize("with_parents", [False, True])
def test_reorder_in_stack_traversal_order_bit_exact(api, scenes, scene, with_parents):
    # the reference's test_reinsertion sequence (bvh2/reinsertion.rs:394-436): VeryLow PLOC, run(0.25),
    # reorder_in_stack_traversal_order (bvh2/mod.rs:462-500), run(0.5) -- every step bit-exact against the oracle, and the
    # reordered tree validates with parents before children
    aabbs = ob.tri_aabbs(scenes[scene])
    want = ob.ploc_build(aabbs, None, 2, 64, 0)
    got = api.PlocBuilder().build(2, aabbs, None, 64, 0)
    if with_parents:
        want.compute_parents()
        got.compute_parents()
    want.reinsertion_run(0.25)
    api.ReinsertionOptimizer().run(got, 0.25)
    want.reorder_in_stack_traversal_order()
    got.reorder_in_stack_traversal_order()
    assert got.children_are_ordered_after_parents
    gn, gp, gpar = got.download(with_parents=True)
    wn, wp, wpar = want.get(with_parents=True)
    assert_nodes_equal(gn, wn, f"{scene} after reorder")
    assert np.array_equal(gp, wp) and np.array_equal(gpar[1:], wpar[1:])
    inner = gn["prim_count"] == 0
    assert np.all(gn["first_index"][inner] > np.nonzero(inner)[0])  # children after parents
    rc, msg = ob.bvh2_from(gn, gp, want.max_depth).validate(aabbs)
    assert rc == 0, msg
    want.reinsertion_run(0.5)
    api.ReinsertionOptimizer().run(got, 0.5)
    assert_nodes_equal(got.download()[0], want.get()[0], f"{scene} reinsertion after reorder")


def test_reorder_degenerate_and_deep(api):
    from test_gpu_parity import graded_boxes

    for n in (1, 2, 3):
        aabbs = graded_boxes(n, 0.05)
        got = api.PlocBuilder().build(6, aabbs, None, 64, 2)
        want = ob.ploc_build(aabbs, None, 6, 64, 2)
        got.reorder_in_stack_traversal_order()
        want.reorder_in_stack_traversal_order()
        assert_nodes_equal(got.download()[0], want.get()[0], f"{n} leaves")
    aabbs = graded_boxes(2000, 0.005)  # a chain ~1000 levels deep: one grid barrier per level
    got = api.PlocBuilder().build(6, aabbs, None, 64, 2)
    want = ob.ploc_build(aabbs, None, 6, 64, 2)
    api.ReinsertionOptimizer().run(got, 0.3)
    want.reinsertion_run(0.3)
    got.reorder_in_stack_traversal_order()
    want.reorder_in_stack_traversal_order()
    assert_nodes_equal(got.download()[0], want.get()[0], "deep chain")
