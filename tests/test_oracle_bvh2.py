"""Pins the CPU oracle for the Bvh2 side of the path: SAH leaf collapse
(bvh2/leaf_collapser.rs), build_bvh2_from_tris (bvh2/builder.rs:17-91) and Bvh2 ray traversal (bvh2/mod.rs:148-334).
Known answers: the reference's own collapse test (leaf_collapser.rs:401-439: validate before/after, with and without
parents), tests/mod.rs degenerate + varying-prim-count builds, and the kitchen golden hash of examples/obj_cwbvh.rs (the
hash depends only on WHICH triangle every pixel sees, so any exact closest-hit traversal of any valid tree reproduces it).
CPU only."""
import numpy as np
import pytest

import oracle_bind as ob
from obvhs_b200 import camera, test_util as tu
from obvhs_b200.types import make_rays

from test_oracle_golden import shade_normals

F32_MAX = np.float32(3.4028235e38)


def test_collapse_reference_test():
    # leaf_collapser.rs:401-439
    tris = tu.demoscene(32, 0)
    aabbs = ob.tri_aabbs(tris)
    for with_parents in (False, True):
        bvh = ob.ploc_build(aabbs, None, 2, 64, 1)
        rc, msg = bvh.validate(aabbs, tight_fit=False)
        assert rc == 0, msg
        if with_parents:
            bvh.compute_parents()
        n_before = bvh.node_count
        bvh.collapse(8, 1.0)
        rc, msg = bvh.validate(aabbs, tight_fit=False)
        assert rc == 0, msg
        assert bvh.node_count < n_before and bvh.prim_count == tris.shape[0]
        assert bvh.has_parents == with_parents  # :183-190 parents recomputed only if they existed
        nodes, prims = bvh.get()
        leaves = nodes[nodes["prim_count"] > 0]
        assert leaves["prim_count"].max() <= 8 and int(leaves["prim_count"].sum()) == tris.shape[0]
        assert np.array_equal(np.sort(prims), np.arange(tris.shape[0], dtype=np.uint32))


def test_collapse_early_outs():
    tris = tu.demoscene(8, 0)
    aabbs = ob.tri_aabbs(tris)
    bvh = ob.ploc_build(aabbs, None, 2, 64, 1)
    n0 = bvh.get()[0].tobytes()
    bvh.collapse(1, 1.0)  # max_prims <= 1 (:25)
    assert bvh.get()[0].tobytes() == n0
    bvh.collapse(200, 1.0)  # nodes <= 2*max_prims+1 (:25) / prims <= max_prims (:29)
    assert bvh.get()[0].tobytes() == n0


@pytest.mark.parametrize("preset", list(ob.BVH2_PRESETS))
def test_kitchen_golden_hash_through_bvh2(kitchen_tris, preset):
    rays = camera.primary_rays(camera.kitchen_camera(32))
    b = ob.build_bvh2_from_tris(kitchen_tris, preset)
    rc, msg = b.validate(ob.tri_aabbs(kitchen_tris), tight_fit=False)
    assert rc == 0, msg
    normals, hits = shade_normals(b, kitchen_tris, rays)
    assert tu.hash_vec3a_vec(normals) == 1343358762
    # the CwBvh built from the same triangles sees the same distances
    c = ob.build_cwbvh_from_tris(kitchen_tris, preset)
    ch = c.ray_traverse(c.bvh_tris(kitchen_tris), rays)
    hit = ch["t"] < F32_MAX
    assert np.array_equal(hits["t"][hit].view(np.uint32), ch["t"][hit].view(np.uint32))
    # bvh2/mod.rs:326: a ray that runs out of stack without a hit leaves hit.t = ray.tmax (f32::MAX here), not +inf
    assert np.all((hits["t"][~hit] == F32_MAX) | np.isinf(hits["t"][~hit]))


def test_bvh2_degenerate_builds_never_hit():
    # tests/mod.rs:35-88 (the Bvh2 flavours)
    ray = make_rays(np.array([[0.0, 0.0, 1.0]], np.float32), np.array([[0.0, 0.0, -1.0]], np.float32), 0.0, np.inf)
    inf = np.float32(np.inf)
    mx = np.float32(3.4028235e38)
    cases = {
        "empty": np.array([[inf, inf, inf, 0, -inf, -inf, -inf, 0]], np.float32),
        "inf": np.tile(np.array([[-inf, -inf, -inf, 0, inf, inf, inf, 0]], np.float32), (10, 1)),
        "max": np.tile(np.array([[-mx, -mx, -mx, 0, mx, mx, mx, 0]], np.float32), (10, 1)),
        "nothing": np.zeros((0, 8), np.float32),
    }
    for name, aabbs in cases.items():
        for sd, thr, ratio, mult, prec, mp, cost in (cfg[:7] for cfg in ob.BVH2_PRESETS.values()):
            b = ob.ploc_build(aabbs, None, sd, prec, thr)
            b.reinsertion_run(ratio)
            b.collapse(mp, cost)
            b.reinsertion_run(ratio * mult)
            # the closure of the reference test never reports a hit: traverse over triangles that cannot be hit
            far = np.zeros((max(1, aabbs.shape[0]), 12), np.float32)
            far[:, [0, 4, 8]] = 1e30
            hits = b.ray_traverse(far, ray)
            assert not (hits["t"][0] < np.inf), name


def test_bvh2_varying_prim_counts_validate():
    # tests/mod.rs:91-102
    tris = tu.flat_plane(4)
    while tris.shape[0] > 1:
        tris = tris[:-1]
        aabbs = ob.tri_aabbs(tris)
        for preset in ob.BVH2_PRESETS:
            b = ob.build_bvh2_from_tris(tris, preset)
            rc, msg = b.validate(aabbs, tight_fit=False)
            assert rc == 0, f"{tris.shape[0]} tris {preset}: {msg}"


def test_bvh2_and_cwbvh_agree_on_incoherent_rays():
    tris = tu.triangle_soup(4096, 3)
    rng = np.random.default_rng(2)
    o = rng.random((20000, 3), dtype=np.float32) * np.float32(1.4) - np.float32(0.2)
    d = rng.standard_normal((20000, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays = make_rays(o, d.astype(np.float32), 0.0, np.inf)
    b = ob.build_bvh2_from_tris(tris, "medium_build")
    c = ob.build_cwbvh_from_tris(tris, "medium_build")
    hb = b.ray_traverse(b.bvh_tris(tris), rays)
    hc = c.ray_traverse(c.bvh_tris(tris), rays)
    hit = hc["t"] < np.inf
    assert hit.sum() > 1000
    assert np.array_equal(hb["t"][hit].view(np.uint32), hc["t"][hit].view(np.uint32))
    assert np.array_equal(np.isinf(hb["t"]), ~hit)  # tmax = inf: the no-hit value is ray.tmax = inf
    _, pb = b.get()
    _, pc, _ = c.get()
    assert np.array_equal(pb[hb["primitive_id"][hit]], pc[hc["primitive_id"][hit]])  # same original triangle (no exact ties here)
    # miss / any-hit flavours agree with the closest-hit result
    srays = rays.copy()
    srays[:, 13] = np.where(hit, hc["t"] * np.float32(1.001), np.float32(5.0))
    assert np.array_equal(b.ray_traverse_miss(b.bvh_tris(tris), srays).astype(bool), ~hit)
    assert np.array_equal(b.ray_traverse_anyhit_count(b.bvh_tris(tris), srays) > 0, hit)


def test_reference_test_reinsertion():
    # bvh2/reinsertion.rs:394-436 as written: VeryLow PLOC over demoscene(32), validate, run(0.25), validate,
    # reorder_in_stack_traversal_order, run(0.5), validate -- with and without parents computed up front
    tris = tu.demoscene(32, 0)
    aabbs = ob.tri_aabbs(tris)
    for with_parents in (False, True):
        bvh = ob.ploc_build(aabbs, np.arange(tris.shape[0], dtype=np.uint32), 2, 64, 1)
        assert bvh.validate(aabbs, tight_fit=False)[0] == 0
        if with_parents:
            bvh.compute_parents()
            assert bvh.validate(aabbs, tight_fit=False)[0] == 0
        n0 = bvh.get()[0].tobytes()
        for ratio in (0.25, 0.5):
            bvh.reinsertion_run(ratio)
            rc, msg = bvh.validate(aabbs, tight_fit=False)
            assert rc == 0, msg
            if ratio == 0.25:
                nodes_before = bvh.get()[0]
                bvh.reorder_in_stack_traversal_order()
                nodes_after = bvh.get()[0]
                # a re-indexing: same multiset of boxes, parents now precede their children
                assert sorted(nodes_before["aabb"].tobytes()[i:i + 32] for i in range(0, nodes_before.shape[0] * 32, 32)) == \
                    sorted(nodes_after["aabb"].tobytes()[i:i + 32] for i in range(0, nodes_after.shape[0] * 32, 32))
                inner = nodes_after["prim_count"] == 0
                assert np.all(nodes_after["first_index"][inner] > np.nonzero(inner)[0])
                assert bvh.validate(aabbs, tight_fit=False)[0] == 0
        assert bvh.get()[0].tobytes() != n0  # the optimizer did move nodes
        assert np.array_equal(np.sort(bvh.get()[1]), np.arange(tris.shape[0], dtype=np.uint32))


def test_reference_test_refit_all():
    # bvh2/mod.rs:1110-1142: move every triangle (scale 1.3, rotate 0.1 rad about y, translate), rewrite the leaf boxes through
    # primitives_to_nodes, refit_all, validate with tight_fit = true
    tris = tu.demoscene(32, 0)
    aabbs = ob.tri_aabbs(tris)
    bvh = ob.ploc_build(aabbs, np.arange(tris.shape[0], dtype=np.uint32), 2, 64, 1)
    c, s = np.float32(np.cos(0.1)), np.float32(np.sin(0.1))
    rot = np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]], np.float32)  # Quat::from_rotation_y
    v = tris.reshape(-1, 3, 4)[:, :, :3] * np.float32(1.3)
    v = (v @ rot.T + np.array([0.33, 0.3, 0.37], np.float32)).astype(np.float32)
    moved = np.zeros_like(tris).reshape(-1, 3, 4)
    moved[:, :, :3] = v
    moved_aabbs = ob.tri_aabbs(np.ascontiguousarray(moved.reshape(-1, 12)))
    assert bvh.validate(moved_aabbs, tight_fit=True)[0] != 0  # stale boxes no longer fit
    bvh.set_leaf_aabbs(moved_aabbs)
    bvh.refit_all()
    rc, msg = bvh.validate(moved_aabbs, tight_fit=True)
    assert rc == 0, msg


def test_reference_test_reinsert_node():
    # bvh2/mod.rs:1144-1163: build_bvh2_from_tris(demoscene(32), fastest_build), then Bvh2::reinsert_node(node_id) for every node
    # id 1..len, validate. reinsert_node (bvh2/mod.rs:763-773) = find_reinsertion + apply when area_diff > 0, which is
    # run_with_candidates over one candidate for one iteration.
    tris = tu.demoscene(32, 0)
    aabbs = ob.tri_aabbs(tris)
    bvh = ob.build_bvh2_from_tris(tris, "fastest_build")
    n0 = bvh.get()[0].tobytes()
    applied = 0
    for node_id in range(1, bvh.node_count):
        applied += bvh.reinsertion_run_with_candidates(np.array([node_id], np.uint32), 1)
    rc, msg = bvh.validate(aabbs, tight_fit=False)
    assert rc == 0, msg
    assert applied > 0 and bvh.get()[0].tobytes() != n0


def test_compute_primitives_to_nodes_host_helper():
    # bvh2/mod.rs:647-665 against the reference's loop written out, on plain, collapsed and pre-split trees
    from obvhs_b200.types import compute_primitives_to_nodes

    tris = tu.soup_with_large_triangles(2000, 30, 4)
    for preset in ("fastest_build", "medium_build", "slow_build"):
        b = ob.build_bvh2_from_tris(tris, preset)
        nodes, prims = b.get()
        want = np.full(prims.shape[0], 0xFFFFFFFF, np.uint32)
        for node_id in range(nodes.shape[0]):
            if nodes["prim_count"][node_id] != 0:
                s = int(nodes["first_index"][node_id])
                for k in range(s, s + int(nodes["prim_count"][node_id])):
                    want[prims[k]] = node_id
        got = compute_primitives_to_nodes(nodes, prims)
        assert np.array_equal(got, want), preset
        held = got[got != 0xFFFFFFFF]
        assert np.all(nodes["prim_count"][held] != 0)
    assert compute_primitives_to_nodes(np.zeros(0, ob.BVH2_NODE), np.zeros(0, np.uint32)).shape == (0,)
