"""Pins the CPU oracle (oracle/obvhs_oracle.cpp) against every known answer the reference's own tests hold for the
hot path. CPU only."""
import numpy as np
import pytest

import oracle_bind as ob
from obvhs_b200 import camera, test_util as tu
from obvhs_b200.types import make_rays

F32_MAX = np.float32(3.4028235e38)


def shade_normals(c, tris, rays):
    """examples/obj_cwbvh.rs:94-117: double-sided geometric normal of the closest hit, (0,0,0) on a miss."""
    bt = c.bvh_tris(tris)
    hits = c.ray_traverse(bt, rays)
    with np.errstate(invalid="ignore"):
        nrm = ob.triangle_normals(bt)
    hit = hits["t"] < F32_MAX
    out = np.zeros((rays.shape[0], 3), np.float32)
    nn = nrm[hits["primitive_id"][hit]]
    d = rays[hit, 4:7]
    s = np.sign((nn[:, 0] * -d[:, 0] + nn[:, 1] * -d[:, 1]) + nn[:, 2] * -d[:, 2]).astype(np.float32)
    out[hit] = nn * s[:, None]
    return out, hits


@pytest.mark.parametrize("preset", ["fastest_build", "very_fast_build", "fast_build", "medium_build", "slow_build", "very_slow_build"])
def test_kitchen_golden_hash(kitchen_tris, preset):
    # examples/obj_cwbvh.rs:142-181: width 32 render, hash of per-pixel normals == 1343358762
    rays = camera.primary_rays(camera.kitchen_camera(32))
    c = ob.build_cwbvh_from_tris(kitchen_tris, preset)
    normals, _ = shade_normals(c, kitchen_tris, rays)
    assert tu.hash_vec3a_vec(normals) == 1343358762
    rc, msg = c.validate(ob.tri_aabbs(kitchen_tris))
    assert rc == 0, msg


def test_icosphere_plane_hits_primitive_62():
    # src/cwbvh/traverse_macro.rs:33-56 and src/lib.rs:16-67
    tris = np.concatenate([tu.icosphere(1), tu.plane()], axis=0)
    c = ob.build_cwbvh_from_tris(tris, "medium_build")
    _, prims, _ = c.get()
    rays = make_rays(np.array([[0.1, 0.1, 4.0]], np.float32), np.array([[0.0, 0.0, -1.0]], np.float32), 0.0, np.inf)
    hits = c.ray_traverse(c.bvh_tris(tris), rays)
    assert hits["t"][0] < np.inf
    assert prims[hits["primitive_id"][0]] == 62


def test_flat_plane_all_hit_normal_up():
    # tests/mod.rs:105-124: 256x256 render of a flat 4x4 plane, every pixel hits and the normal is exactly +Y
    tris = tu.flat_plane(4)
    cam = camera.Camera(256, 256, 90.0, (0.0, 0.9, 0.0), (0.0, 0.0, 0.0), up=(1.0, 0.0, 0.0))
    rays = camera.primary_rays(cam, tmax=np.inf, column_major=True)
    for preset in ob.PRESETS:
        c = ob.build_cwbvh_from_tris(tris, preset)
        bt = c.bvh_tris(tris)
        hits = c.ray_traverse(bt, rays)
        assert np.all(hits["t"] < np.inf)
        n = ob.triangle_normals(bt)[hits["primitive_id"]]
        assert np.all(n == np.array([0.0, 1.0, 0.0], np.float32))


def test_degenerate_builds_do_not_hit():
    # tests/mod.rs:65-88: one empty AABB, and nothing at all
    ray = make_rays(np.array([[0.0, 0.0, 1.0]], np.float32), np.array([[0.0, 0.0, -1.0]], np.float32), 0.0, np.inf)
    empty = np.array([[F32_MAX, F32_MAX, F32_MAX, 0, -F32_MAX, -F32_MAX, -F32_MAX, 0]], np.float32)
    for sd, thr, ratio, prec, mp in (cfg[:5] for cfg in ob.PRESETS.values()):
        b = ob.ploc_build(empty, None, sd, prec, thr)
        b.reinsertion_run(ratio)
        c = b.to_cwbvh(min(max(mp, 1), 3))
        assert c.node_count == 1
        # the closure of the reference test returns +inf for every primitive: use a far-away degenerate triangle
        tri = np.zeros((1, 12), np.float32)
        hits = c.ray_traverse(tri, ray)
        assert not (hits["t"][0] < np.inf)
        b0 = ob.ploc_build(np.zeros((0, 8), np.float32), None, sd, prec, thr)
        b0.reinsertion_run(ratio)
        c0 = b0.to_cwbvh(3)
        assert c0.node_count == 0
        assert not (c0.ray_traverse(np.zeros((0, 12), np.float32), ray)["t"][0] < np.inf)


def test_varying_prim_counts_validate():
    # tests/mod.rs:91-102: 31 .. 1 triangles x presets, Bvh2 + CwBvh validate
    tris = tu.flat_plane(4)
    assert tris.shape[0] == 32
    for n in range(31, 0, -1):
        t = tris[:n]
        aabbs = ob.tri_aabbs(t)
        for name, (sd, thr, ratio, prec, mp) in ((k, cfg[:5]) for k, cfg in ob.PRESETS.items()):
            b = ob.ploc_build(aabbs, None, sd, prec, thr)
            rc, msg = b.validate(aabbs)
            assert rc == 0, (n, msg)
            b.reinsertion_run(ratio)
            rc, msg = b.validate(aabbs)
            assert rc == 0, (n, msg)
            c = b.to_cwbvh(min(max(mp, 1), 3))
            rc, msg = c.validate(aabbs)
            assert rc == 0, (n, msg)
            # the reference's loop goes through the triangle builders, pre-splits included (tests/mod.rs:95-98)
            b2 = ob.build_bvh2_from_tris(t, name)
            rc, msg = b2.validate(aabbs, tight_fit=name not in ("slow_build", "very_slow_build"))
            assert rc == 0, (n, name, msg)
            rc, msg = ob.build_cwbvh_from_tris(t, name).validate(aabbs)
            assert rc == 0, (n, name, msg)


def test_morton_split_matches_naive_interleave():
    # src/ploc/morton.rs:35-58 magic masks == bit by bit interleave
    rng = np.random.default_rng(1)
    for a in list(rng.integers(0, 1 << 21, 200)) + [0, 1, (1 << 21) - 1]:
        want = 0
        for i in range(21):
            want |= ((int(a) >> i) & 1) << (3 * i)
        assert ob.lib().orc_test_split3_64(int(a)) == want


def test_sort_is_stable_by_index(kitchen_tris):
    # ties keep ascending original index
    aabbs = ob.tri_aabbs(kitchen_tris)
    lo, hi, order, total = ob.morton_sort(aabbs)
    keys = lo[order]
    assert np.all(keys[:-1] <= keys[1:])
    same = keys[:-1] == keys[1:]
    assert same.sum() > 1000  # the kitchen has thousands of tied codes
    assert np.all(order[:-1][same] < order[1:][same])


def test_refit_fast_vs_full_differ_only_in_zero_sign(scenes):
    # bvh2/mod.rs:722-751: the release build returns early after two unchanged ancestors, the debug build walks to the
    # root. Both give the same tree VALUES; the early return can only keep the old sign of a zero lane. The GPU path
    # matches the walk-to-the-root bits (a full bottom-up refit of the dirty paths).
    for name in ("cornell", "terrain32", "soup4k", "kitchen"):
        aabbs = ob.tri_aabbs(scenes[name])
        for ratio in (0.02, 0.5):
            res, applied = [], []
            for full in (0, 1):
                ob.lib().orc_set_refit_full(full)
                b = ob.ploc_build(aabbs, None, 6, 64, 2)
                applied.append(b.reinsertion_run(ratio))
                res.append(b.get()[0])
            ob.lib().orc_set_refit_full(0)
            assert applied[0] == applied[1]
            assert np.array_equal(res[0]["first_index"], res[1]["first_index"]) and np.array_equal(res[0]["prim_count"], res[1]["prim_count"])
            assert np.array_equal(res[0]["aabb"], res[1]["aabb"]), name  # value compare: -0.0 == +0.0
            bits_differ = res[0]["aabb"].view(np.uint32) != res[1]["aabb"].view(np.uint32)
            assert np.all(res[0]["aabb"][bits_differ] == 0.0)


def test_exact_node_aabbs_bound_by_quantised_boxes(scenes):
    # bvh2_to_cwbvh(.., include_exact_node_aabbs = true) (bvh2_to_cwbvh.rs:60-80): entry i is the unquantised box of wide node i;
    # the node's quantisation frame (p .. p + extent * 255, cwbvh/node.rs:261-265) must contain it, the root's is the scene box
    aabbs = ob.tri_aabbs(scenes["kitchen"])
    b = ob.ploc_build(aabbs, None, 6, 64, 2)
    c = b.to_cwbvh(3, True, True)
    nodes, _, total = c.get()
    ex = c.exact_node_aabbs()
    assert ex.shape[0] == b.node_count
    m = nodes.shape[0]
    assert np.array_equal(ex[0, [0, 1, 2, 4, 5, 6]], total[[0, 1, 2, 4, 5, 6]])
    extent = (nodes["e"].astype(np.uint32) << 23).view(np.float32)
    lo, hi = nodes["p"], nodes["p"] + extent * np.float32(255.0)
    assert np.all(ex[:m, 0:3] >= lo) and np.all(ex[:m, 4:7] <= hi + np.abs(hi) * 1e-6 + 1e-6)
    assert np.all(ex[m:, 0] == np.float32(3.4028235e38)) and np.all(ex[m:, 4] == np.float32(-3.4028235e38))  # Aabb::empty()
    assert ob.ploc_build(aabbs, None, 6, 64, 2).to_cwbvh(3, True, False).exact_node_aabbs() is None


def ray_new_cases():
    """Directions around every branch of safe_inverse (ray.rs:6-12): +-0, denormals, |x| == EPSILON and its neighbours, huge, inf."""
    eps = np.float32(1.1920929e-07)
    special = np.array([0.0, -0.0, 1e-45, -1e-45, eps, -eps, np.nextafter(eps, np.float32(1)), np.nextafter(eps, np.float32(0)),
                        -np.nextafter(eps, np.float32(1)), 1.0, -1.0, 3.0, 1e-7, -1e-7, 1e30, -1e30, 3.4028235e38, np.inf, -np.inf], np.float32)
    rng = np.random.default_rng(11)
    d = np.concatenate([np.stack([special, np.roll(special, 1), np.roll(special, 7)], axis=1),
                        rng.standard_normal((5000, 3)).astype(np.float32) * np.float32(1e-6),
                        rng.standard_normal((5000, 3)).astype(np.float32)], axis=0)
    o = rng.standard_normal(d.shape).astype(np.float32)
    return o, d


def test_host_ray_helpers_match_the_oracle_ray_new():
    # obvhs_b200.types.make_rays (numpy) is what the tests and bench.py build rays with: pin it to the oracle's Ray::new, and
    # make_ray_args / ray_args_of to the 32-byte argument record of include/obvhs_cuda.h (ObvhsRayNew)
    from obvhs_b200.types import make_ray_args, make_rays, ray_args_of

    o, d = ray_new_cases()
    want = ob.make_rays(o, d, 0.25, 7.5)
    with np.errstate(all="ignore"):
        got = make_rays(o, d, 0.25, 7.5)
    assert got.view(np.uint32).tobytes() == want.view(np.uint32).tobytes()
    a = make_ray_args(o, d, 0.25, 7.5)
    assert a.shape == (o.shape[0], 8) and a.dtype == np.float32
    assert np.array_equal(a[:, 0:3], o) and np.array_equal(a[:, 4:7].view(np.uint32), d.view(np.uint32))
    assert np.all(a[:, 3] == np.float32(0.25)) and np.all(a[:, 7] == np.float32(7.5))
    assert ray_args_of(want).view(np.uint32).tobytes() == a.view(np.uint32).tobytes()


def _aabb(mn, mx):
    a = np.zeros(8, np.float32)
    a[0:3], a[4:7] = mn, mx
    return a


def test_reference_aabb_unit_tests():
    # src/aabb.rs:222-360: the known answers of the reference's own Aabb tests, through the primitives the oracle builds on
    L = ob.lib()
    p = lambda a: a.ctypes.data_as(ob.C.c_void_p)  # noqa: E731
    unit = _aabb([0, 0, 0], [1, 1, 1])
    assert L.orc_aabb_half_area(p(unit)) == 3.0                                   # test_half_area (:313-317)
    out = np.zeros(8, np.float32)
    L.orc_aabb_union(p(unit), p(_aabb([0.5, 0.5, 0.5], [1.5, 1.5, 1.5])), p(out))    # test_union (:252-259)
    assert np.array_equal(out[[0, 1, 2, 4, 5, 6]], np.array([0, 0, 0, 1.5, 1.5, 1.5], np.float32))
    assert L.orc_aabb_intersect_aabb(p(unit), p(_aabb([0.5] * 3, [1.5] * 3))) == 1  # test_intersect_aabb (:341-348)
    assert L.orc_aabb_intersect_aabb(p(unit), p(_aabb([1.5] * 3, [2.5] * 3))) == 0
    pt = lambda x: np.array(x, np.float32)  # noqa: E731
    assert L.orc_aabb_contains_point(p(unit), p(pt([0.5, 0.5, 0.5]))) == 1         # test_contains_point (:237-242)
    assert L.orc_aabb_contains_point(p(unit), p(pt([1.5, 1.5, 1.5]))) == 0
    assert L.orc_aabb_contains_point(p(unit), p(pt([1.0, 0.0, 1.0]))) == 1         # closed interval (cmpge / cmple)
    ray = ob.make_rays(np.array([[-1, -1, -1]], np.float32), np.array([[1, 1, 1]], np.float32), 0.0, 3.4028235e38)
    assert L.orc_aabb_intersect_ray(p(unit), p(ray)) == 1.0                         # test_intersect_ray (:350-357)
    ray2 = ob.make_rays(np.array([[2, 2, 2]], np.float32), np.array([[1, 1, 1]], np.float32), 0.0, 3.4028235e38)
    assert L.orc_aabb_intersect_ray(p(unit), p(ray2)) == np.inf


def cwbvh_parents_numpy(nodes):
    """CwBvh::compute_parents (cwbvh/mod.rs:494-509) restated with numpy over CWBVH_NODE records."""
    parents = np.zeros(nodes.shape[0], np.uint32)
    for ch in range(8):
        inner = ((nodes["imask"] & (1 << ch)) != 0) & (nodes["child_meta"][:, ch] != 0)
        idx = np.nonzero(inner)[0]
        slot = (nodes["child_meta"][idx, ch] & 0b11111).astype(np.int64) - 24
        below = nodes["imask"][idx].astype(np.int64) & ((1 << slot) - 1)
        rel = np.array([bin(int(b)).count("1") for b in below], np.int64)
        parents[nodes["child_base_idx"][idx].astype(np.int64) + rel] = idx
    return parents


def test_reference_exact_aabbs_cwbvh_and_compute_parents():
    # tests/mod.rs:387-446 (exact_aabbs_cwbvh): every inner child's exact box lies inside the parent's quantised child box AND inside
    # the child node's own quantised frame; tests/mod.rs:325-349 (compute_parents_cwbvh): every node but the root is an inner child of
    # exactly the node that references it. Accessors restated from cwbvh/node.rs:251-304 (child_aabb, aabb, is_leaf, child_node_index).
    tris = tu.demoscene(100, 0)
    aabbs = ob.tri_aabbs(tris)
    bvh2 = ob.ploc_build(aabbs, np.arange(tris.shape[0], dtype=np.uint32), 1, 64, 0)  # very_fast_build (lib.rs:246-256)
    cw = bvh2.to_cwbvh(3, True, True)
    assert cw.validate(aabbs)[0] == 0
    nodes, _, _ = cw.get()
    exact = cw.exact_node_aabbs()
    assert exact is not None and exact.shape[0] >= nodes.shape[0]
    e = (nodes["e"].astype(np.uint32) << 23).view(np.float32)                      # compute_extent
    p = nodes["p"]
    own_min, own_max = p, p + e * np.float32(255.0)                                 # CwBvhNode::aabb, NQ_SCALE = 255
    parent_of = np.full(nodes.shape[0], -1, np.int64)
    checked = 0
    for ch in range(8):
        inner = (nodes["imask"] & (1 << ch)) != 0                                   # !is_leaf(ch)
        idx = np.nonzero(inner)[0]
        slot = (nodes["child_meta"][idx, ch] & 0b11111).astype(np.int64) - 24
        rel = np.array([bin(int(m) & ~(0xFFFFFFFF << int(s)) & 0xFFFFFFFF).count("1") for m, s in zip(nodes["imask"][idx], slot)], np.int64)
        child = nodes["child_base_idx"][idx].astype(np.int64) + rel                 # child_node_index
        cmin = np.stack([nodes[f"child_min_{a}"][idx, ch] for a in "xyz"], axis=1).astype(np.float32) * e[idx] + p[idx]   # child_aabb
        cmax = np.stack([nodes[f"child_max_{a}"][idx, ch] for a in "xyz"], axis=1).astype(np.float32) * e[idx] + p[idx]
        ex_min, ex_max = exact[child][:, 0:3], exact[child][:, 4:7]
        assert np.all(ex_min >= cmin) and np.all(ex_max <= cmax)
        assert np.all(ex_min >= own_min[child]) and np.all(ex_max <= own_max[child])
        assert np.all(parent_of[child] == -1)                                       # referenced once
        parent_of[child] = idx
        checked += idx.size
    assert checked == nodes.shape[0] - 1 and parent_of[0] == -1 and np.all(parent_of[1:] >= 0)
    assert np.array_equal(cwbvh_parents_numpy(nodes)[1:], parent_of[1:].astype(np.uint32))


def test_reference_reuse_allocs():
    # tests/mod.rs:448-500: a builder reused for a second, smaller scene (build_with_bvh) gives a tree that validates tightly.
    # (On the GPU path the reuse is the context's result cache and arena: tests/test_gpu_parity.py builds many scenes per context.)
    for side in (99, 98):
        tris = tu.demoscene(side, 0)
        aabbs = ob.tri_aabbs(tris)
        bvh = ob.ploc_build(aabbs, np.arange(tris.shape[0], dtype=np.uint32), 14, 64, 0)  # PlocSearchDistance::default() = Medium
        rc, msg = bvh.validate(aabbs, tight_fit=True)
        assert rc == 0, msg


def test_reference_order_children_cwbvh():
    # tests/mod.rs:351-385 (order_children_cwbvh) and :431-445 (the tail of exact_aabbs_cwbvh): CwBvh::order_node_children on every
    # node, then CwBvh::order_children, validating after each; with exact_node_aabbs present they move with their nodes.
    tris = tu.demoscene(100, 0)
    aabbs = ob.tri_aabbs(tris)
    rays = camera.primary_rays(camera.Camera(64, 64, 60.0, np.array([0.5, 0.9, 1.7], np.float32), np.array([0.5, 0.0, 0.5], np.float32)))
    for with_exact in (False, True):
        bvh2 = ob.ploc_build(aabbs, np.arange(tris.shape[0], dtype=np.uint32), 1, 64, 0)  # very_fast_build
        cw = bvh2.to_cwbvh(3, True, with_exact)
        assert cw.validate(aabbs)[0] == 0
        before = cw.ray_traverse(cw.bvh_tris(tris), rays)
        n0 = cw.get()[0].tobytes()
        if not with_exact:
            for node in range(cw.node_count):
                cw.order_node_children(aabbs, node, False)
            rc, msg = cw.validate(aabbs)
            assert rc == 0, msg
        cw.order_children(aabbs, False)
        rc, msg = cw.validate(aabbs)
        assert rc == 0, msg
        cw.order_children(aabbs, False)
        rc, msg = cw.validate(aabbs)
        assert rc == 0, msg
        nodes, prims, _ = cw.get()
        assert nodes.tobytes() != n0  # "a slightly different order" than the converter's (cwbvh/mod.rs:510-513)
        # the same tree in another child order: identical closest hits (distances bit for bit)
        after = cw.ray_traverse(cw.bvh_tris(tris), rays)
        assert np.array_equal(after["t"].view(np.uint32), before["t"].view(np.uint32)) and (before["t"] < 3e38).sum() > 1000
        assert np.array_equal(cwbvh_parents_numpy(nodes)[1:] < np.arange(1, nodes.shape[0]), np.ones(nodes.shape[0] - 1, bool))
        if with_exact:  # exact boxes still inside their (moved) nodes' frames
            exact = cw.exact_node_aabbs()
            e = (nodes["e"].astype(np.uint32) << 23).view(np.float32)
            assert np.all(exact[: nodes.shape[0], 0:3] >= nodes["p"]) and np.all(exact[: nodes.shape[0], 4:7] <= nodes["p"] + e * np.float32(255.0))
