"""GPU parity: every stage of the CUDA path, called through the C ABI (obvhs_b200.api -> libobvhs_cuda.so), against the
CPU oracle on the same inputs. Integer / byte / index results are compared bit for bit; hit distances too (the contract
allows 1e-6 relative, the implementation is bit-exact because it keeps the reference's operation order without FMA)."""
import numpy as np
import pytest

import oracle_bind as ob
from obvhs_b200 import camera, test_util as tu
from obvhs_b200.types import make_rays

pytestmark = pytest.mark.gpu

F32_MAX = np.float32(3.4028235e38)
SCENES = ["cornell", "ico_plane", "flat4", "terrain32", "soup4k", "kitchen"]
PLOC_CONFIGS = [(1, 0), (2, 0), (6, 2), (14, 3), (24, 1), (32, 0)]


@pytest.fixture(scope="module", autouse=True)
def oracle_refit_semantics():
    """Parity target for inner-node AABB bits after reinsertion: Bvh2::refit_from_fast WITHOUT its release-only early
    return (src/bvh2/mod.rs:732-737, `#[cfg(not(debug_assertions))]`), i.e. the reference built with debug assertions,
    where every ancestor is recomputed as first.union(second). The release early-out can keep a stale -0.0 / +0.0 sign on
    a lane whose value is unchanged (tests/test_oracle_golden.py::test_refit_fast_vs_full_differ_only_in_zero_sign)."""
    ob.lib().orc_set_refit_full(1)
    yield
    ob.lib().orc_set_refit_full(0)


@pytest.fixture(scope="module")
def api():
    from obvhs_b200 import api as a

    a.default_context(0)  # raises if the CUDA library or device is missing: no fallback
    return a


def node_fields(nodes):
    """Comparable view of Bvh2Node arrays (the Vec3A padding lanes are unspecified)."""
    a = np.ascontiguousarray(nodes["aabb"][:, [0, 1, 2, 4, 5, 6]]).view(np.uint32)
    return a, nodes["prim_count"], nodes["first_index"]


def assert_nodes_equal(got, want, what):
    ga, gp, gf = node_fields(got)
    wa, wp, wf = node_fields(want)
    assert got.shape == want.shape, what
    bad = np.nonzero((ga != wa).any(axis=1) | (gp != wp) | (gf != wf))[0]
    assert bad.size == 0, f"{what}: {bad.size} of {got.shape[0]} nodes differ, first {bad[:5]}: got {got[bad[:2]]} want {want[bad[:2]]}"


def rays_for(tris, n_side=96, seed=7):
    """Primary-style rays from a camera outside the scene box plus random incoherent rays through it."""
    lo = tris.reshape(-1, 4)[:, :3].min(axis=0)
    hi = tris.reshape(-1, 4)[:, :3].max(axis=0)
    c = (lo + hi) * 0.5
    ext = max(float(np.max(hi - lo)), 1e-3)
    eye = c + np.array([0.9, 0.7, 1.3], np.float32) * ext
    cam = camera.Camera(n_side, n_side, 60.0, eye.astype(np.float32), c.astype(np.float32))
    prim = camera.primary_rays(cam)
    rng = np.random.default_rng(seed)
    m = n_side * n_side
    o = (c + (rng.random((m, 3), dtype=np.float32) - 0.5) * ext * 2.5).astype(np.float32)
    t = (c + (rng.random((m, 3), dtype=np.float32) - 0.5) * ext).astype(np.float32)
    d = t - o
    d /= np.maximum(np.linalg.norm(d, axis=1, keepdims=True), 1e-20)
    rnd = make_rays(o, d.astype(np.float32), 0.0, np.inf)
    # axis-aligned directions exercise the safe_inverse / zero-direction paths
    axis = make_rays(np.tile(c, (6, 1)).astype(np.float32) + np.float32(0.01), np.array(
        [[1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0], [0, 0, 1], [0, 0, -1]], np.float32), 0.0, np.inf)
    return np.concatenate([prim, rnd, axis], axis=0)


@pytest.mark.parametrize("precision", [64, 128])
@pytest.mark.parametrize("scene", SCENES)
def test_morton_codes_and_sorted_order(api, scenes, scene, precision):
    aabbs = ob.tri_aabbs(scenes[scene])
    lo, hi, order, total = ob.morton_sort(aabbs, precision)
    glo, ghi, gorder, gtotal = api.PlocBuilder().morton_sort(aabbs, api.SortPrecision(precision))
    assert np.array_equal(glo, lo)
    assert np.array_equal(ghi, hi)
    assert precision == 64 or hi.any()
    assert np.array_equal(gorder, order)
    assert np.array_equal(gtotal[[0, 1, 2, 4, 5, 6]], total[[0, 1, 2, 4, 5, 6]])


def test_sort_heavy_ties_and_sizes(api):
    # every primitive identical (all codes tie), and sizes around the tile boundaries of the onesweep sort
    tri = tu.cube()[:1]
    for n in (1, 2, 3, 255, 256, 257, 4095, 4096, 4097, 8193, 70001):
        aabbs = ob.tri_aabbs(np.repeat(tri, n, axis=0))
        _, _, order, _ = api.PlocBuilder().morton_sort(aabbs)
        assert np.array_equal(order, np.arange(n, dtype=np.uint32)), n
    soup = tu.triangle_soup(70001, 11)
    aabbs = ob.tri_aabbs(soup)
    lo, _, order, _ = ob.morton_sort(aabbs)
    glo, _, gorder, _ = api.PlocBuilder().morton_sort(aabbs)
    assert np.array_equal(glo, lo) and np.array_equal(gorder, order)


@pytest.mark.parametrize("scene", SCENES)
@pytest.mark.parametrize("cfg", PLOC_CONFIGS)
def test_ploc_bvh2_bit_exact(api, scenes, scene, cfg):
    sd, thr = cfg
    aabbs = ob.tri_aabbs(scenes[scene])
    want = ob.ploc_build(aabbs, None, sd, 64, thr)
    wn, wp = want.get()
    got = api.PlocBuilder().build(sd, aabbs, None, api.SortPrecision.U64, thr)
    gn, gp = got.download()
    assert got.ploc_iterations == want.ploc_iterations
    assert got.max_depth == want.max_depth
    assert np.array_equal(gp, wp)
    assert_nodes_equal(gn, wn, f"{scene} ploc r={sd} thr={thr}")
    assert got.children_are_ordered_after_parents


@pytest.mark.parametrize("scene", SCENES)
@pytest.mark.parametrize("cfg", [(24, 2), (14, 1), (1, 0)])
def test_ploc_u128_bvh2_bit_exact(api, scenes, scene, cfg):
    # SortPrecision::U128 (ploc/mod.rs:686-701, morton.rs:65-89): the sort of the slow / very_slow presets (lib.rs:282-305)
    sd, thr = cfg
    aabbs = ob.tri_aabbs(scenes[scene])
    want = ob.ploc_build(aabbs, None, sd, 128, thr)
    wn, wp = want.get()
    got = api.PlocBuilder().build(sd, aabbs, None, api.SortPrecision.U128, thr)
    gn, gp = got.download()
    assert got.ploc_iterations == want.ploc_iterations
    assert np.array_equal(gp, wp)
    assert_nodes_equal(gn, wn, f"{scene} ploc u128 r={sd} thr={thr}")


def test_u128_sort_separates_points_that_tie_in_63_bits(api):
    # centres closer than 2^-21 of the scene extent share a 63-bit code but not a 126-bit one; order must follow the
    # 126-bit code (and stay stable for exact duplicates)
    rng = np.random.default_rng(5)
    n = 20000
    c = np.zeros((n, 3), np.float64)
    c[:] = rng.random((n // 100, 3)).repeat(100, axis=0) + rng.random((n, 3)) * 2.0 ** -24
    c[-1] = 0.0
    c[-2] = 1.0
    c[1000:1100] = c[1000]  # exact duplicates
    aabbs = np.zeros((n, 8), np.float32)
    aabbs[:, 0:3] = c
    aabbs[:, 4:7] = c
    lo, hi, order, _ = ob.morton_sort(aabbs, 128)
    glo, ghi, gorder, _ = api.PlocBuilder().morton_sort(aabbs, api.SortPrecision.U128)
    assert np.array_equal(glo, lo) and np.array_equal(ghi, hi) and np.array_equal(gorder, order)
    lo64, _, order64, _ = ob.morton_sort(aabbs, 64)
    assert not np.array_equal(order64, order)
    k = [(int(h) << 64) | int(l) for h, l in zip(ghi[gorder], glo[gorder])]
    assert all(a <= b for a, b in zip(k, k[1:]))


def test_ploc_from_triangles_equals_from_aabbs(api, scenes):
    tris = scenes["kitchen"]
    a = api.PlocBuilder().build_tris(6, tris, api.SortPrecision.U64, 2).download()
    b = api.PlocBuilder().build(6, ob.tri_aabbs(tris), None, api.SortPrecision.U64, 2).download()
    assert_nodes_equal(a[0], b[0], "tris vs aabbs")


def test_ploc_with_custom_indices(api, scenes):
    aabbs = ob.tri_aabbs(scenes["soup4k"])
    idx = np.random.default_rng(0).permutation(aabbs.shape[0]).astype(np.uint32)
    wn, wp = ob.ploc_build(aabbs, idx, 6, 64, 2).get()
    gn, gp = api.PlocBuilder().build(6, aabbs, idx, api.SortPrecision.U64, 2).download()
    assert np.array_equal(gp, wp)
    assert_nodes_equal(gn, wn, "custom indices")


def test_compute_parents_and_refit_all(api, scenes):
    aabbs = ob.tri_aabbs(scenes["kitchen"])
    want = ob.ploc_build(aabbs, None, 6, 64, 2)
    wn, wp = want.get()
    got = api.Bvh2.upload(wn, wp, want.max_depth, True)
    got.compute_parents()
    want.compute_parents()
    _, _, wpar = want.get(with_parents=True)
    gn, gp, gpar = got.download(with_parents=True)
    assert np.array_equal(gpar, wpar)
    # move the leaves, refit (config 5): bvh2/mod.rs:527-569
    rng = np.random.default_rng(5)
    moved = aabbs.copy()
    delta = (rng.random((aabbs.shape[0], 3), dtype=np.float32) - 0.5) * np.float32(0.05)
    moved[:, 0:3] += delta
    moved[:, 4:7] += delta
    want.set_leaf_aabbs(moved)
    want.refit_all()
    got.set_leaf_aabbs(moved)
    assert_nodes_equal(got.download()[0], want.get()[0], "refit_all")


@pytest.mark.parametrize("scene", SCENES)
@pytest.mark.parametrize("ratio", [0.02, 0.5, 1.0])
def test_reinsertion_bit_exact(api, scenes, scene, ratio):
    aabbs = ob.tri_aabbs(scenes[scene])
    want = ob.ploc_build(aabbs, None, 6, 64, 2)
    wn, wp = want.get()
    got = api.Bvh2.upload(wn, wp, want.max_depth, True)
    applied_want = want.reinsertion_run(ratio)
    opt = api.ReinsertionOptimizer()
    applied_got = opt.run(got, ratio)
    wn2, _, wpar = want.get(with_parents=True)
    gn2, _, gpar = got.download(with_parents=True)
    assert applied_got == applied_want
    assert_nodes_equal(gn2, wn2, f"{scene} reinsertion ratio={ratio}")
    assert np.array_equal(gpar[1:], wpar[1:])
    assert not got.children_are_ordered_after_parents
    rc, msg = ob.bvh2_from(gn2, wp, want.max_depth).validate(aabbs)
    assert rc == 0, msg


def test_reinsertion_custom_sequence(api, scenes):
    aabbs = ob.tri_aabbs(scenes["terrain32"])
    want = ob.ploc_build(aabbs, None, 1, 64, 0)
    wn, wp = want.get()
    got = api.Bvh2.upload(wn, wp, want.max_depth, True)
    seq = [1.0, 0.5, 0.25, 1.0]
    want.reinsertion_run(0.7, seq)
    api.ReinsertionOptimizer().run(got, 0.7, seq)
    assert_nodes_equal(got.download()[0], want.get()[0], "custom ratio sequence")


@pytest.mark.parametrize("scene", ["terrain32", "kitchen"])
def test_reinsertion_run_with_candidates(api, scenes, scene):
    # ReinsertionOptimizer::run_with_candidates (reinsertion.rs:66-90): caller-chosen node ids, unsorted, with duplicates
    aabbs = ob.tri_aabbs(scenes[scene])
    want = ob.ploc_build(aabbs, None, 2, 64, 0)
    wn, wp = want.get()
    got = api.Bvh2.upload(wn, wp, want.max_depth, True)
    rng = np.random.default_rng(11)
    ids = rng.integers(1, wn.shape[0], size=min(3000, wn.shape[0]), dtype=np.uint32)
    a_want = want.reinsertion_run_with_candidates(ids, 3)
    opt = api.ReinsertionOptimizer()
    a_got = opt.run_with_candidates(got, ids, 3)
    assert a_got == a_want and a_got > 0
    assert_nodes_equal(got.download()[0], want.get()[0], f"{scene} run_with_candidates")
    assert not got.children_are_ordered_after_parents
    import torch

    a_want = want.reinsertion_run_with_candidates(ids[::2], 1)
    assert opt.run_with_candidates(got, torch.from_numpy(ids[::2].astype(np.int64)).to(torch.int32).cuda(), 1) == a_want  # device-resident ids
    assert_nodes_equal(got.download()[0], want.get()[0], f"{scene} run_with_candidates (device ids)")
    with pytest.raises(api.ObvhsError):
        opt.run_with_candidates(got, np.array([0], np.uint32), 1)  # the root cannot be reinserted (the reference asserts)
    with pytest.raises(api.ObvhsError):
        opt.run_with_candidates(got, np.array([wn.shape[0]], np.uint32), 1)


def test_reference_test_reinsert_node_on_gpu(api):
    # bvh2/mod.rs:1144-1163: Bvh2::reinsert_node for every node id in turn (one candidate, one iteration per call), the tree
    # compared with the oracle's after the whole sequence and validated
    tris = tu.demoscene(16, 0)
    aabbs = ob.tri_aabbs(tris)
    want = ob.build_bvh2_from_tris(tris, "fastest_build")
    wn, wp = want.get()
    got = api.Bvh2.upload(wn, wp, want.max_depth, True)
    opt = api.ReinsertionOptimizer()
    for node_id in range(1, wn.shape[0]):
        ids = np.array([node_id], np.uint32)
        assert opt.run_with_candidates(got, ids, 1) == want.reinsertion_run_with_candidates(ids, 1), node_id
    assert_nodes_equal(got.download()[0], want.get()[0], "reinsert_node sequence")
    rc, msg = ob.bvh2_from(got.download()[0], wp, want.max_depth).validate(aabbs, tight_fit=False)
    assert rc == 0, msg


@pytest.mark.parametrize("scene", SCENES)
@pytest.mark.parametrize("max_prims,order", [(1, True), (3, True), (2, False)])
def test_cwbvh_collapse_bytes(api, scenes, scene, max_prims, order):
    aabbs = ob.tri_aabbs(scenes[scene])
    want2 = ob.ploc_build(aabbs, None, 6, 64, 2)
    want2.reinsertion_run(0.1)
    wn, wp = want2.get()
    want = want2.to_cwbvh(max_prims, order)
    wnodes, wprims, wtotal = want.get()
    got = api.bvh2_to_cwbvh(api.Bvh2.upload(wn, wp, want2.max_depth, False), max_prims, order)
    gnodes, gprims, gtotal = got.download()
    assert gnodes.shape == wnodes.shape
    assert np.array_equal(gprims, wprims)
    bad = np.nonzero(gnodes.view(np.uint8).reshape(-1, 80) != wnodes.view(np.uint8).reshape(-1, 80))[0]
    assert bad.size == 0, f"{np.unique(bad).size} CWBVH nodes differ, first: got {gnodes[bad[0]]} want {wnodes[bad[0]]}"
    assert np.array_equal(gtotal[[0, 1, 2, 4, 5, 6]], wtotal[[0, 1, 2, 4, 5, 6]])


def test_cwbvh_from_fresh_ploc_without_parents(api, scenes):
    # fastest_build: no reinsertion, Bvh2::parents is None when the converter runs
    aabbs = ob.tri_aabbs(scenes["kitchen"])
    want = ob.ploc_build(aabbs, None, 1, 64, 0).to_cwbvh(1, True).get()
    got = api.bvh2_to_cwbvh(api.PlocBuilder().build(1, aabbs), 1, True).download()
    assert got[0].tobytes() == want[0].tobytes()
    assert np.array_equal(got[1], want[1])


TRAVERSE_MODES = ["auto", "static", "persistent:8:32", "persistent:1:32", "persistent:32:64", "persistent:16:128"]


@pytest.mark.parametrize("mode", TRAVERSE_MODES)
@pytest.mark.parametrize("scene", SCENES)
def test_traversal_hits_bit_exact(api, scenes, scene, mode):
    # every kernel variant (one ray per thread, persistent warps with ray refill, and the per-block choice between
    # them) runs the reference's per-ray state machine: hits, distances AND the visit counters are identical
    tris = scenes[scene]
    c = ob.build_cwbvh_from_tris(tris, "fast_build")
    nodes, prims, total = c.get()
    bt = c.bvh_tris(tris)
    rays = rays_for(tris)
    wc = np.zeros(2, np.uint64)
    want = c.ray_traverse(bt, rays, counters=wc)
    g = api.CwBvh.upload(nodes, prims, total, ctx=api.Context(0, traverse=mode))
    g.set_triangles(tris)
    gc = np.zeros(2, np.uint64)
    got = g.ray_traverse(rays, counters=gc)
    assert np.array_equal(got["primitive_id"], want["primitive_id"])
    assert np.array_equal(got["t"].view(np.uint32), want["t"].view(np.uint32))
    assert np.array_equal(got["geometry_id"], want["geometry_id"]) and np.array_equal(got["instance_id"], want["instance_id"])
    assert np.array_equal(gc, wc), "nodes visited / triangles tested differ: the visit order is not the reference's"
    assert np.array_equal(g.ray_traverse(rays)["primitive_id"], want["primitive_id"])  # uncounted kernel variant
    # shadow / any-hit flavours with a finite tmax (cwbvh/mod.rs:201-245)
    srays = rays.copy()
    finite = np.isfinite(want["t"])
    srays[:, 13] = np.where(finite, want["t"] * np.float32(0.999), np.float32(5.0))
    srays[::3, 13] = np.float32(1e30)
    assert np.array_equal(g.ray_traverse_miss(srays), c.ray_traverse_miss(bt, srays))
    assert np.array_equal(g.ray_traverse_anyhit_count(srays), c.ray_traverse_anyhit_count(bt, srays))


@pytest.mark.parametrize("variant", ["1", "2", "3", "4"])
@pytest.mark.parametrize("scene", ["soup4k", "kitchen"])
def test_persistent_kernel_variants_bit_exact(api, scenes, scene, variant):
    # the persistent kernel's scheduling policy (which lanes move in a turn), shared-memory short stack and register cap are
    # tuning knobs: every variant must return the oracle's hits AND visit counters, for Ray structs and Ray::new records alike
    from obvhs_b200.types import ray_args_of

    tris = scenes[scene]
    c = ob.build_cwbvh_from_tris(tris, "fast_build")
    nodes, prims, total = c.get()
    bt = c.bvh_tris(tris)
    rays = rays_for(tris)
    wc = np.zeros(2, np.uint64)
    want = c.ray_traverse(bt, rays, counters=wc)
    ctx = api.Context(0, traverse="persistent:4:32")
    ctx.set_option("traverse_variant", variant)
    g = api.CwBvh.upload(nodes, prims, total, ctx=ctx)
    g.set_triangles(tris)
    gc = np.zeros(2, np.uint64)
    got = g.ray_traverse(rays, counters=gc)
    assert got.tobytes() == want.tobytes()
    assert np.array_equal(gc, wc)
    assert g.ray_traverse(ray_args_of(rays)).tobytes() == want.tobytes()  # Ray::new inside the kernel
    srays = rays.copy()
    finite = np.isfinite(want["t"])
    srays[:, 13] = np.where(finite, want["t"] * np.float32(0.999), np.float32(5.0))
    srays[::3, 13] = np.float32(1e30)
    assert np.array_equal(g.ray_traverse_miss(srays), c.ray_traverse_miss(bt, srays))
    assert np.array_equal(g.ray_traverse_miss(ray_args_of(srays)), c.ray_traverse_miss(bt, srays))
    assert np.array_equal(g.ray_traverse_anyhit_count(srays), c.ray_traverse_anyhit_count(bt, srays))


@pytest.mark.parametrize("preset", ["fastest_build", "very_fast_build", "fast_build", "medium_build"])
@pytest.mark.parametrize("scene", ["cornell", "terrain32", "kitchen"])
def test_build_cwbvh_from_tris_end_to_end(api, scenes, scene, preset):
    tris = scenes[scene]
    want = ob.build_cwbvh_from_tris(tris, preset)
    wnodes, wprims, wtotal = want.get()
    t = [0.0]
    got = api.build_cwbvh_from_tris(tris, api.BvhBuildParams.preset(preset), t)
    gnodes, gprims, gtotal = got.download()
    assert t[0] > 0.0
    assert gnodes.shape == wnodes.shape
    assert np.array_equal(gprims, wprims)
    assert gnodes.tobytes() == wnodes.tobytes()
    rc, msg = ob.cwbvh_from(gnodes, gprims, gtotal).validate(ob.tri_aabbs(tris))
    assert rc == 0, msg


def test_kitchen_golden_hash_on_gpu(api, kitchen_tris):
    # examples/obj_cwbvh.rs:142-181 through the GPU path only (build + traversal), normals on the host
    rays = camera.primary_rays(camera.kitchen_camera(32))
    for preset in ("fastest_build", "fast_build", "medium_build"):
        bvh = api.build_cwbvh_from_tris(kitchen_tris, api.BvhBuildParams.preset(preset))
        hits = bvh.ray_traverse(rays)
        _, prims, _ = bvh.download()
        with np.errstate(invalid="ignore"):
            nrm = ob.triangle_normals(kitchen_tris[prims])
        hit = hits["t"] < F32_MAX
        out = np.zeros((rays.shape[0], 3), np.float32)
        nn = nrm[hits["primitive_id"][hit]]
        d = rays[hit, 4:7]
        s = np.sign((nn[:, 0] * -d[:, 0] + nn[:, 1] * -d[:, 1]) + nn[:, 2] * -d[:, 2]).astype(np.float32)
        out[hit] = nn * s[:, None]
        assert tu.hash_vec3a_vec(out) == 1343358762, preset


def test_icosphere_known_answer_on_gpu(api):
    tris = np.concatenate([tu.icosphere(1), tu.plane()], axis=0)
    bvh = api.build_cwbvh_from_tris(tris, api.BvhBuildParams.medium_build())
    rays = api.make_rays(np.array([[0.1, 0.1, 4.0, 0.0, 0.0, -1.0]], np.float32), 0.0, np.inf)
    hits = bvh.ray_traverse(rays)
    _, prims, _ = bvh.download()
    assert hits["t"][0] < np.inf and prims[hits["primitive_id"][0]] == 62


def test_make_rays_matches_ray_new(api):
    rng = np.random.default_rng(3)
    od = rng.standard_normal((1000, 6)).astype(np.float32)
    od[:10, 3] = 0.0
    od[10:20, 4] = -0.0
    od[20:30, 5] = np.float32(1e-8)
    want = make_rays(od[:, 0:3], od[:, 3:6], 0.25, 77.0)
    got = api.make_rays(od, 0.25, 77.0)
    assert np.array_equal(got[:, [0, 1, 2, 4, 5, 6, 8, 9, 10, 12, 13]].view(np.uint32), want[:, [0, 1, 2, 4, 5, 6, 8, 9, 10, 12, 13]].view(np.uint32))


def test_degenerate_inputs(api):
    # tests/mod.rs:35-102: nothing, one empty AABB, tiny counts
    ray = make_rays(np.array([[0.0, 0.0, 1.0]], np.float32), np.array([[0.0, 0.0, -1.0]], np.float32), 0.0, np.inf)
    for preset in ("fastest_build", "fast_build", "medium_build"):
        p = api.BvhBuildParams.preset(preset)
        bvh = api.build_cwbvh_from_tris(np.zeros((0, 12), np.float32), p)
        assert bvh.node_count == 0 and bvh.prim_count == 0
        assert not (bvh.ray_traverse(ray)["t"][0] < np.inf)
        assert bvh.ray_traverse_miss(ray)[0] == 1
    empty = np.array([[F32_MAX, F32_MAX, F32_MAX, 0, -F32_MAX, -F32_MAX, -F32_MAX, 0]], np.float32)
    b = api.PlocBuilder().build(14, empty)
    assert b.node_count == 1
    c = api.bvh2_to_cwbvh(b, 3, True)
    want = ob.ploc_build(empty, None, 14, 64, 0).to_cwbvh(3, True).get()
    assert c.download()[0].tobytes() == want[0].tobytes()
    tris = tu.flat_plane(4)
    for n in range(31, 0, -3):
        for preset in ("fastest_build", "fast_build", "medium_build"):
            want = ob.build_cwbvh_from_tris(tris[:n], preset).get()
            got = api.build_cwbvh_from_tris(tris[:n], api.BvhBuildParams.preset(preset)).download()
            assert got[0].tobytes() == want[0].tobytes(), (n, preset)
            assert np.array_equal(got[1], want[1])


def test_nan_input_is_an_error_not_a_hang(api):
    aabbs = ob.tri_aabbs(tu.triangle_soup(1000, 1))
    aabbs[17, 1] = np.nan
    with pytest.raises(api.ObvhsError) as e:
        api.PlocBuilder().build(6, aabbs)
    assert e.value.code == -4


def test_unsupported_paths_fail_loudly(api):
    tris = tu.cornell_box()
    with pytest.raises(api.ObvhsError):
        api.PlocBuilder().build(7, ob.tri_aabbs(tris))  # not a PlocSearchDistance
    with pytest.raises(api.ObvhsError):
        api.PlocBuilder().build(6, ob.tri_aabbs(tris), None, 96)  # not a SortPrecision


@pytest.mark.parametrize("scene", ["cornell", "terrain32", "kitchen"])
def test_cwbvh_exact_node_aabbs(api, scenes, scene):
    # bvh2_to_cwbvh(.., include_exact_node_aabbs = true) (bvh2_to_cwbvh.rs:60-80): the unquantised box of every wide node
    aabbs = ob.tri_aabbs(scenes[scene])
    want = ob.ploc_build(aabbs, None, 6, 64, 2).to_cwbvh(3, True, True)
    got = api.bvh2_to_cwbvh(api.PlocBuilder().build(6, aabbs, None, api.SortPrecision.U64, 2), 3, True, True)
    assert got.download()[0].tobytes() == want.get()[0].tobytes()
    w, g = want.exact_node_aabbs(), got.exact_node_aabbs()
    assert g.shape == w.shape and g.shape[0] == 2 * aabbs.shape[0] - 1
    assert np.array_equal(g[:, [0, 1, 2, 4, 5, 6]].view(np.uint32), w[:, [0, 1, 2, 4, 5, 6]].view(np.uint32))
    assert api.bvh2_to_cwbvh(api.PlocBuilder().build(6, aabbs), 3, True, False).exact_node_aabbs() is None


@pytest.mark.parametrize("scene", ["cornell", "terrain32", "kitchen"])
def test_cwbvh_compute_parents(api, scenes, scene):
    # CwBvh::compute_parents (cwbvh/mod.rs:494-509) and the property the reference's test checks (tests/mod.rs:325-349): the
    # parent's inner child slots contain the child
    import torch
    from test_oracle_golden import cwbvh_parents_numpy

    tris = scenes[scene]
    bvh = api.build_cwbvh_from_tris(tris, api.BvhBuildParams.fast_build())
    nodes, _, _ = bvh.download()
    want = cwbvh_parents_numpy(nodes)
    got = bvh.compute_parents()
    assert got[0] == 0 and np.array_equal(got, want)
    d = torch.empty(nodes.shape[0], dtype=torch.int32, device="cuda")
    bvh.compute_parents(out=d)
    bvh.ctx.synchronize()
    assert np.array_equal(d.cpu().numpy().view(np.uint32), want)
    if nodes.shape[0] > 1:
        assert np.all(got[1:] < np.arange(1, nodes.shape[0]))  # parents come first in the converter's layout
    empty = api.build_cwbvh_from_tris(np.zeros((0, 12), np.float32), api.BvhBuildParams.fast_build())
    assert empty.compute_parents().shape[0] == empty.node_count


def test_large_scene_full_parity_and_properties(api):
    # 1M-triangle soup + 0.5M terrain: full byte parity against the oracle (seconds on the CPU) plus size-independent
    # properties: sorted keys, stable ties, valid trees, closest hit <= any brute-force sample
    for name, tris in (("soup1m", tu.triangle_soup(1_000_000, 2)), ("terrain500", tu.demoscene(500, 1))):
        aabbs = ob.tri_aabbs(tris)
        lo, _, order, _ = api.PlocBuilder().morton_sort(aabbs)
        k = lo[order]
        assert np.all(k[:-1] <= k[1:])
        tie = k[:-1] == k[1:]
        assert np.all(order[:-1][tie] < order[1:][tie])
        want = ob.build_cwbvh_from_tris(tris, "fast_build", threads=0)
        got = api.build_cwbvh_from_tris(tris, api.BvhBuildParams.fast_build())
        gnodes, gprims, gtotal = got.download()
        wnodes, wprims, _ = want.get()
        assert gnodes.shape == wnodes.shape, name
        assert np.array_equal(gprims, wprims), name
        assert gnodes.tobytes() == wnodes.tobytes(), name
        rc, msg = ob.cwbvh_from(gnodes, gprims, gtotal).validate(aabbs)
        assert rc == 0, msg
        rays = rays_for(tris, n_side=256)
        wh = want.ray_traverse(want.bvh_tris(tris), rays)
        gh = got.ray_traverse(rays)
        assert np.array_equal(gh["primitive_id"], wh["primitive_id"]), name
        assert np.array_equal(gh["t"].view(np.uint32), wh["t"].view(np.uint32)), name


def test_traversal_pinned_host_buffers_zero_copy(api, scenes):
    # pinned host rays / hits are read and written by the kernel directly (no staging copies): same results
    import torch

    tris = scenes["kitchen"]
    bvh = api.build_cwbvh_from_tris(tris, api.BvhBuildParams.fast_build())
    rays = rays_for(tris)
    want = bvh.ray_traverse(rays)  # pageable host path (staged)
    from obvhs_b200.types import RAY_HIT

    h_rays = torch.from_numpy(rays).pin_memory()
    h_hits = torch.empty((rays.shape[0], 4), dtype=torch.int32).pin_memory()
    got = bvh.ray_traverse(h_rays.numpy(), out=h_hits.numpy().view(RAY_HIT).reshape(-1))
    assert np.array_equal(got["primitive_id"], want["primitive_id"])
    assert np.array_equal(got["t"].view(np.uint32), want["t"].view(np.uint32))
    d_rays = h_rays.cuda()
    d_hits = torch.empty((rays.shape[0], 4), dtype=torch.int32, device="cuda")
    bvh.ray_traverse(d_rays, out=d_hits)
    bvh.ctx.synchronize()
    assert np.array_equal(d_hits.cpu().numpy().view(RAY_HIT).reshape(-1)["primitive_id"], want["primitive_id"])


@pytest.mark.parametrize("n_side", [190, 301])
def test_traversal_host_batches_are_pipelined_in_chunks(api, scenes, n_side):
    # host ray batches >= 64 Ki rays take the chunked three-stream path (H2D | kernel | D2H overlapped): results must
    # equal the single-launch device-resident path for closest hit, miss and all-hit counts, ragged tail included
    import torch

    tris = scenes["kitchen"]
    bvh = api.build_cwbvh_from_tris(tris, api.BvhBuildParams.fast_build())
    rays = rays_for(tris, n_side=n_side)
    assert rays.shape[0] >= 65536 and rays.shape[0] % 128 != 0
    d_rays = torch.from_numpy(rays).cuda()
    d_hits = torch.empty((rays.shape[0], 4), dtype=torch.int32, device="cuda")
    bvh.ray_traverse(d_rays, out=d_hits)
    bvh.ctx.synchronize()
    want = d_hits.cpu().numpy()
    got = bvh.ray_traverse(rays)  # pageable host memory
    assert np.array_equal(got["primitive_id"], want[:, 0].view(np.uint32))
    assert np.array_equal(got["t"].view(np.uint32), want[:, 3].view(np.uint32))
    from obvhs_b200.types import RAY_HIT

    h_rays = torch.from_numpy(rays).pin_memory()
    h_hits = torch.empty((rays.shape[0], 4), dtype=torch.int32).pin_memory()
    got = bvh.ray_traverse(h_rays.numpy(), out=h_hits.numpy().view(RAY_HIT).reshape(-1))
    assert np.array_equal(h_hits.numpy(), want)
    miss = bvh.ray_traverse_miss(rays)
    assert np.array_equal(miss.astype(bool), want[:, 3].view(np.float32) >= F32_MAX)


def test_ray_new_on_device_is_ray_new_of_the_reference(api):
    # ray.rs:6-12,34-52 against the oracle's restatement: safe_inverse bit for bit, including +-0, denormals, |x| == EPSILON and
    # its neighbours and inf; host and device buffers; per-ray tmin / tmax
    import torch
    from obvhs_b200.types import make_ray_args
    from test_oracle_golden import ray_new_cases

    o, d = ray_new_cases()
    rng = np.random.default_rng(5)
    tmin = rng.random(d.shape[0], dtype=np.float32)
    tmax = tmin + rng.random(d.shape[0], dtype=np.float32) * np.float32(100)
    want = ob.make_rays(o, d, 0.0, 0.0)
    want[:, 12] = tmin
    want[:, 13] = tmax
    args = make_ray_args(o, d, tmin, tmax)
    assert api.ray_new(args).view(np.uint32).tobytes() == want.view(np.uint32).tobytes()
    d_rays = torch.empty((d.shape[0], 16), dtype=torch.float32, device="cuda")
    api.ray_new(torch.from_numpy(args).cuda(), out=d_rays)
    assert d_rays.cpu().numpy().view(np.uint32).tobytes() == want.view(np.uint32).tobytes()
    assert api.ray_new(np.zeros((0, 8), np.float32)).shape == (0, 16)
    nan = api.ray_new(make_ray_args(o[:1], np.array([[np.nan, 1.0, 0.0]], np.float32)))  # (debug_assert!ed away in the reference)
    assert np.isnan(nan[0, 8]) and nan[0, 9] == 1.0 and nan[0, 10] == np.float32(8388608.0)


@pytest.mark.parametrize("n_side", [96, 190])
def test_traversal_over_ray_new_arguments(api, scenes, n_side):
    # the *_ray_new_* entry points (32-byte constructor arguments, constructor on the device) give the results of the Ray-struct
    # calls for every flavour, tree type and buffer placement: small staged batch (n_side 96) and chunked pipeline (190)
    import torch
    from obvhs_b200.types import RAY_HIT, ray_args_of

    tris = scenes["kitchen"]
    rays = rays_for(tris, n_side=n_side)
    rays[::5, 12] = np.float32(0.25)  # per-ray tmin / tmax survive the packing
    rays[::7, 13] = np.float32(2.5)
    args = ray_args_of(rays)
    assert args.shape[1] == 8
    c = ob.build_cwbvh_from_tris(tris, "fast_build")
    bt = c.bvh_tris(tris)
    want = c.ray_traverse(bt, rays)
    cw = api.build_cwbvh_from_tris(tris, api.BvhBuildParams.fast_build())
    for a in (args, torch.from_numpy(args).pin_memory().numpy()):
        got = cw.ray_traverse(a)
        assert np.array_equal(got["primitive_id"], want["primitive_id"])
        assert np.array_equal(got["t"].view(np.uint32), want["t"].view(np.uint32))
    d_hits = torch.empty((rays.shape[0], 4), dtype=torch.int32, device="cuda")
    cw.ray_traverse(torch.from_numpy(args).cuda(), out=d_hits)
    cw.ctx.synchronize()
    assert np.array_equal(d_hits.cpu().numpy().view(RAY_HIT).reshape(-1)["primitive_id"], want["primitive_id"])
    assert np.array_equal(cw.ray_traverse_miss(args), c.ray_traverse_miss(bt, rays))
    assert np.array_equal(cw.ray_traverse_anyhit_count(args), c.ray_traverse_anyhit_count(bt, rays))
    b2 = api.build_bvh2_from_tris(tris, api.BvhBuildParams.fast_build())
    assert np.array_equal(b2.ray_traverse(args).view(np.uint32), b2.ray_traverse(rays).view(np.uint32))
    assert np.array_equal(b2.ray_traverse_miss(args), b2.ray_traverse_miss(rays))
    assert cw.ray_traverse(np.zeros((0, 8), np.float32)).shape[0] == 0


@pytest.mark.parametrize("n_side", [96, 190])
@pytest.mark.parametrize("bounds", [(0.0, np.inf), (0.0, 3.4028234663852886e38), (0.3, 2.5)])
def test_traversal_over_origin_direction_records(api, scenes, n_side, bounds):
    # the *_ray_od_* entry points (24 bytes per ray, ONE tmin / tmax per batch: Ray::new_inf, ray.rs:55-57, and the
    # Ray::new(o, d, 0.0, f32::MAX) of the examples) equal the oracle on the expanded rays: host, pinned and device buffers,
    # the staged (96) and the pipelined (190) host path, closest hit / miss, both tree types, every kernel choice
    import torch
    from obvhs_b200.types import RAY_HIT

    tris = scenes["kitchen"]
    base = rays_for(tris, n_side=n_side)
    od = np.ascontiguousarray(base[:, [0, 1, 2, 4, 5, 6]])
    tmin, tmax = np.float32(bounds[0]), np.float32(bounds[1])
    rays = ob.make_rays(od[:, 0:3], od[:, 3:6], tmin, tmax)
    c = ob.build_cwbvh_from_tris(tris, "fast_build")
    bt = c.bvh_tris(tris)
    want = c.ray_traverse(bt, rays)
    want_miss = c.ray_traverse_miss(bt, rays)
    for mode in ("auto", "persistent", "static"):
        ctx = api.Context(0, traverse=mode)
        cw = api.build_cwbvh_from_tris(tris, api.BvhBuildParams.fast_build(), ctx=ctx)
        for a in (od, torch.from_numpy(od).pin_memory().numpy()):
            got = cw.ray_od_traverse(a, tmin, tmax)
            assert np.array_equal(got["primitive_id"], want["primitive_id"]), mode
            assert np.array_equal(got["t"].view(np.uint32), want["t"].view(np.uint32)), mode
        d_hits = torch.empty((od.shape[0], 4), dtype=torch.int32, device="cuda")
        cw.ray_od_traverse(torch.from_numpy(od).cuda(), tmin, tmax, out=d_hits)
        ctx.synchronize()
        assert d_hits.cpu().numpy().view(RAY_HIT).reshape(-1).tobytes() == want.tobytes(), mode
        assert np.array_equal(cw.ray_od_traverse_miss(od, tmin, tmax), want_miss), mode
        h8 = cw.ray_od_traverse(od, tmin, tmax, hit8=True)  # {primitive_id, t} records: 8 bytes per ray back
        assert h8.dtype.itemsize == 8 and np.array_equal(h8["primitive_id"], want["primitive_id"]), mode
        assert np.array_equal(h8["t"].view(np.uint32), want["t"].view(np.uint32)), mode
        d8 = torch.empty((od.shape[0], 2), dtype=torch.int32, device="cuda")
        cw.ray_od_traverse(torch.from_numpy(od).cuda(), tmin, tmax, out=d8, hit8=True)
        ctx.synchronize()
        assert d8.cpu().numpy().tobytes() == h8.tobytes(), mode
        assert cw.ray_od_traverse(np.zeros((0, 6), np.float32)).shape[0] == 0
    b2 = api.build_bvh2_from_tris(tris, api.BvhBuildParams.fast_build())
    assert np.array_equal(b2.ray_od_traverse(od, tmin, tmax).view(np.uint32), b2.ray_traverse(rays).view(np.uint32))
    # an odd-sized slice of a device buffer that is only 8-byte aligned
    d_od = torch.from_numpy(od).cuda()
    part = d_od[3:1003]
    got = cw.ray_od_traverse(part, tmin, tmax)
    assert got.tobytes() == want[3:1003].tobytes()
    lib, h = cw.ctx.lib, cw.ctx.h
    hits = np.zeros(64, dtype=np.uint8)
    assert lib.obvhs_cuda_cwbvh_ray_od_traverse_batch(h, cw.h, None, 4, 0.0, 1.0, hits.ctypes.data) < 0
    assert lib.obvhs_cuda_cwbvh_ray_od_traverse_batch(h, None, od.ctypes.data, 4, 0.0, 1.0, hits.ctypes.data) < 0
    assert lib.obvhs_cuda_cwbvh_ray_od_traverse_batch(h, cw.h, None, 0, 0.0, 1.0, None) == 0


@pytest.mark.parametrize("mode", ["persistent", "auto", "static"])
def test_host_slices_on_two_compute_streams(api, scenes, mode):
    # persistent-kernel slices of a host batch alternate between two compute streams (the tail of one hides behind the next);
    # forced here with a small host_slice: same hits, and the counters of all slices add up
    tris = scenes["kitchen"]
    rays = rays_for(tris, n_side=190)
    c = ob.build_cwbvh_from_tris(tris, "fast_build")
    nodes, prims, total = c.get()
    wc = np.zeros(2, np.uint64)
    want = c.ray_traverse(c.bvh_tris(tris), rays, counters=wc)
    ctx = api.Context(0, traverse=mode)
    ctx.set_option("host_slice", "9000")
    g = api.CwBvh.upload(nodes, prims, total, ctx=ctx)
    g.set_triangles(tris)
    from obvhs_b200.types import ray_args_of

    for r in (rays, ray_args_of(rays)):
        got = g.ray_traverse(r)
        assert np.array_equal(got["primitive_id"], want["primitive_id"])
        assert np.array_equal(got["t"].view(np.uint32), want["t"].view(np.uint32))
    gc = np.zeros(2, np.uint64)
    g.ray_traverse(rays, counters=gc)
    assert np.array_equal(gc, wc)
    import torch

    d_hits = torch.empty((rays.shape[0], 4), dtype=torch.int32, device="cuda")  # host rays, device hits: joined on the context's stream
    g.ray_traverse(rays, out=d_hits)
    ctx.synchronize()
    assert np.array_equal(d_hits.cpu().numpy()[:, 0].view(np.uint32), want["primitive_id"])
    with pytest.raises(api.ObvhsError):
        ctx.set_option("host_slice", "-1")


def test_ray_new_entry_points_edge_cases(api, scenes):
    # error behaviour of the Ray::new entry points matches the Ray-struct ones; extreme slice sizes are clamped, not rejected
    from obvhs_b200.types import ray_args_of

    tris = scenes["cornell"]
    rays = rays_for(tris, n_side=64)
    args = ray_args_of(rays)
    ctx = api.Context(0, traverse="persistent")
    bvh = api.build_cwbvh_from_tris(tris, api.BvhBuildParams.medium_build(), ctx=ctx)
    want = bvh.ray_traverse(rays)
    for sl in ("1", "1000000000", "0"):  # 1 -> 1024-ray slices (one launch each, event pool grows), huge -> a single shot
        ctx.set_option("host_slice", sl)
        got = bvh.ray_traverse(args)
        assert got.tobytes() == want.tobytes(), sl
    lib, h = ctx.lib, ctx.h
    hits = np.zeros(4, dtype=np.uint8)
    assert lib.obvhs_cuda_cwbvh_ray_new_traverse_batch(h, bvh.h, None, 4, hits.ctypes.data) < 0  # null args
    assert lib.obvhs_cuda_cwbvh_ray_new_traverse_batch(h, None, args.ctypes.data, 4, hits.ctypes.data) < 0  # null bvh
    assert lib.obvhs_cuda_cwbvh_ray_new_traverse_batch(h, bvh.h, None, 0, None) == 0  # empty batch
    assert lib.obvhs_cuda_ray_new_batch(h, None, 3, hits.ctypes.data) < 0
    bare = api.CwBvh.upload(*bvh.download(), ctx=ctx)  # no triangles attached
    with pytest.raises(api.ObvhsError):
        bare.ray_traverse(args)


def displaced_aabbs(tris, frame):
    """BASELINE config 5 / SURVEY.md 8(d) S4: every vertex moved by 0.01*(hash_noise-0.5) seeded by the frame."""
    t = tris.reshape(-1, 3, 4).copy()
    k = np.arange(t.shape[0] * 9, dtype=np.uint32)
    noise = tu.hash_noise(k, np.uint32(frame), np.uint32(17)).reshape(-1, 3, 3)
    t[:, :, 0:3] += (noise - np.float32(0.5)) * np.float32(0.01)
    return ob.tri_aabbs(np.ascontiguousarray(t.reshape(-1, 12)))


@pytest.mark.parametrize("scene", ["terrain32", "kitchen"])
def test_dynamic_frames_refit_and_reinsertion(api, scenes, scene):
    # config 5 (examples/physics.rs update loop): per frame rewrite leaf AABBs, refit_all (bvh2/mod.rs:527-569, both the
    # ordered sweep and, after the first reinsertion, the unordered path), then ReinsertionOptimizer::run(0.01)
    tris = scenes[scene]
    aabbs = ob.tri_aabbs(tris)
    want = ob.ploc_build(aabbs, None, 6, 64, 2)
    wn, wp = want.get()
    got = api.Bvh2.upload(wn, wp, want.max_depth, True)
    opt = api.ReinsertionOptimizer()
    for frame in range(4):
        moved = displaced_aabbs(tris, frame)
        want.set_leaf_aabbs(moved)
        want.refit_all()
        got.set_leaf_aabbs(moved)
        assert_nodes_equal(got.download()[0], want.get()[0], f"{scene} frame {frame} refit")
        a_want = want.reinsertion_run(0.01)
        a_got = opt.run(got, 0.01)
        assert a_got == a_want
        assert_nodes_equal(got.download()[0], want.get()[0], f"{scene} frame {frame} reinsertion")
        rc, msg = ob.bvh2_from(got.download()[0], wp, want.max_depth).validate(moved)
        assert rc == 0, msg


def graded_boxes(n, g):
    """n boxes on a line whose spacing grows by the factor (1 + g): PLOC merges about one pair per iteration, so the tree is a
    chain of depth ~n/2 (300 boxes at 5 %: max_depth 156; 2000 at 0.5 %: ~1000) -- beyond the reference's fixed 96 / 192-entry
    stacks, where it switches to HeapStack (faststack.rs:44-47)."""
    c = np.cumprod(np.full(n, 1.0 + g))
    a = np.zeros((n, 8), np.float32)
    h = (c * g * 0.25).astype(np.float32)
    a[:, 0], a[:, 4] = c - h, c + h
    a[:, 1], a[:, 5], a[:, 2], a[:, 6] = -0.01, 0.01, -0.01, 0.01
    return a


@pytest.mark.parametrize("n,g", [(300, 0.05), (2000, 0.005)])
def test_deep_trees_reinsertion_and_builders(api, n, g):
    # max_depth > 96: find_reinsertion runs on a heap stack of 2 * max_depth entries, like the reference; every preset builds
    aabbs = graded_boxes(n, g)
    want = ob.ploc_build(aabbs, None, 6, 64, 2)
    assert want.max_depth > 96
    got = api.PlocBuilder().build(6, aabbs, None, 64, 2)
    assert got.max_depth == want.max_depth
    for ratio in (0.25, 1.0):
        wa = want.reinsertion_run(ratio)
        ga = api.ReinsertionOptimizer().run(got, ratio)
        assert ga == wa
        gn, _, gpar = got.download(with_parents=True)
        wn, _, wpar = want.get(with_parents=True)
        assert_nodes_equal(gn, wn, f"graded {n} boxes, ratio {ratio}")
        assert np.array_equal(gpar[1:], wpar[1:])
    # build_cwbvh<T: Boundable> (cwbvh/builder.rs:98-123) for all six presets: (search distance, threshold, ratio, precision)
    presets = {"fastest_build": (1, 0, 0.0, 64, 1), "very_fast_build": (1, 0, 0.01, 64, 8), "fast_build": (6, 2, 0.02, 64, 8),
               "medium_build": (14, 3, 0.05, 64, 8), "slow_build": (24, 2, 0.2, 128, 8), "very_slow_build": (14, 1, 1.0, 128, 8)}
    for preset, (sd, thr, ratio, prec, mp) in presets.items():
        b = ob.ploc_build(aabbs, None, sd, prec, thr)
        b.reinsertion_run(ratio)
        w = b.to_cwbvh(min(max(mp, 1), 3), True).get()
        c = api.build_cwbvh(aabbs, api.BvhBuildParams.preset(preset)).download()
        assert c[0].tobytes() == w[0].tobytes(), preset
        assert np.array_equal(c[1], w[1])


@pytest.mark.parametrize("scene", ["cornell", "ico_plane", "terrain32", "soup4k", "kitchen"])
@pytest.mark.parametrize("exact", [False, True])
def test_cwbvh_order_children_bit_exact(api, scenes, scene, exact):
    # CwBvh::order_children as a separate pass (cwbvh/mod.rs:520-735; the reference's order_children_cwbvh test, tests/mod.rs:351-385):
    # node bytes (and exact boxes) equal to the oracle's sequential loop, for both primitive layouts; hits unchanged
    tris = scenes[scene]
    aabbs = ob.tri_aabbs(tris)
    for direct in (False, True):
        wb = ob.ploc_build(aabbs, None, 6, 64, 2)
        wb.reinsertion_run(0.02)
        w = wb.to_cwbvh(3, False, exact)  # built WITHOUT the converter's ordering, so that the pass has work to do
        gb = api.PlocBuilder().build(6, aabbs, None, 64, 2)
        api.ReinsertionOptimizer().run(gb, 0.02)
        g = api.bvh2_to_cwbvh(gb, 3, False, exact)
        assert g.download()[0].tobytes() == w.get()[0].tobytes()
        prims = w.get()[1]
        pa = aabbs[prims] if direct else aabbs
        g.set_triangles(tris)
        rays = rays_for(tris)
        before = g.ray_traverse(rays)
        w.order_children(pa, direct)
        g.order_children(pa, direct)
        gn, gp, _ = g.download()
        wn, wp, _ = w.get()
        assert gn.tobytes() == wn.tobytes(), f"{scene} direct={direct} exact={exact}"
        assert np.array_equal(gp, wp)
        if exact:
            assert np.array_equal(g.exact_node_aabbs()[:, [0, 1, 2, 4, 5, 6]], w.exact_node_aabbs()[:, [0, 1, 2, 4, 5, 6]])
        after = g.ray_traverse(rays)
        assert np.array_equal(after["primitive_id"], before["primitive_id"]) and np.array_equal(after["t"].view(np.uint32), before["t"].view(np.uint32))
        rc, msg = ob.cwbvh_from(gn, gp, g.total_aabb()).validate(aabbs)
        assert rc == 0, msg
        # a second pass over an already ordered tree is deterministic as well
        w.order_children(pa, direct)
        g.order_children(pa, direct)
        assert g.download()[0].tobytes() == w.get()[0].tobytes()
    with pytest.raises(api.ObvhsError):  # fewer boxes than the leaves refer to: reported, not read out of bounds
        g.order_children(aabbs[: max(1, aabbs.shape[0] // 2)], False)
