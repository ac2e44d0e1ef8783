"""GPU parity for spatial pre-splits (reference src/splits.rs) and the two presets built on them (slow_build,
very_slow_build: pre-splits + SortPrecision::U128, src/lib.rs:282-305), through the C ABI against the CPU oracle."""
import numpy as np
import pytest

import oracle_bind as ob
from obvhs_b200 import camera, test_util as tu
from test_gpu_parity import F32_MAX, api, assert_nodes_equal, oracle_refit_semantics, rays_for  # noqa: F401  (fixtures)
from test_oracle_splits import split_scenes

pytestmark = pytest.mark.gpu

LANES = [0, 1, 2, 4, 5, 6]  # the Vec3A padding lanes are unspecified


@pytest.mark.parametrize("name", ["splitty", "slivers", "terrain_plus", "kitchen", "cornell"])
def test_presplit_tris_bit_exact(api, scenes, name):
    tris = scenes[name] if name in scenes else split_scenes()[name]
    wa, widx, wavg, wlargest = ob.presplit_tris(tris)
    ga, gidx, gavg, glargest = api.presplit_tris(tris)
    assert gavg.view(np.uint32) == wavg.view(np.uint32), "sequential f32 sum of the half areas"
    assert glargest == wlargest
    assert ga.shape == wa.shape
    assert np.array_equal(gidx, widx)
    assert np.array_equal(ga[:, LANES].view(np.uint32), wa[:, LANES].view(np.uint32))


def test_split_aabbs_precise_and_preset_entry_points(api):
    tris = split_scenes()["splitty"]
    n = tris.shape[0]
    aabbs = ob.tri_aabbs(tris)
    ident = np.arange(n, dtype=np.uint32)
    _, _, avg, largest = ob.presplit_tris(tris)
    wa, widx, _, _ = ob.presplit_tris(tris)
    ga, gidx = api.split_aabbs_preset(aabbs, ident, tris, avg, largest)
    assert np.array_equal(gidx, widx) and np.array_equal(ga[:, LANES].view(np.uint32), wa[:, LANES].view(np.uint32))
    # the general form with other parameters, a permuted index map and a partial set of AABBs
    perm = np.random.default_rng(1).permutation(n).astype(np.uint32)
    for args in [(0.01, 0.5, 1.5, 1.2, 3, 8), (0.002, 0.05, 1.1, 1.05, 12, 5), (0.01, 0.5, 1.5, 1.2, 0, 8), (1e9, 1e9, 1.8, 1.6, 12, 12)]:
        wa, widx = ob.split_aabbs_precise(aabbs[perm], perm, tris, *args)
        ga, gidx = api.split_aabbs_precise(aabbs[perm], perm, tris, *args)
        assert ga.shape == wa.shape, args
        assert np.array_equal(gidx, widx) and np.array_equal(ga[:, LANES].view(np.uint32), wa[:, LANES].view(np.uint32)), args


def test_split_capacity_is_reported_not_overrun(api):
    import ctypes as C

    tris = split_scenes()["slivers"]
    n = tris.shape[0]
    ctx = api.default_context()
    a = np.zeros((n + 8, 8), np.float32)
    a[:n] = ob.tri_aabbs(tris)
    a[n:] = 777.0
    idx = np.zeros(n + 8, np.uint32)
    idx[:n] = np.arange(n)
    _, _, avg, largest = ob.presplit_tris(tris)
    count = C.c_size_t(0)
    rc = ctx.lib.obvhs_cuda_split_aabbs_preset(ctx.h, api._ptr(a), api._ptr(idx), n, n + 8, api._ptr(tris), n, float(avg), float(largest),
                                               C.byref(count))
    assert rc == api.ERR_CAPACITY and count.value == ob.presplit_tris(tris)[0].shape[0]
    assert np.all(a[n:] == 777.0) and np.array_equal(a[:n], ob.tri_aabbs(tris))  # nothing written


@pytest.mark.parametrize("preset", ["slow_build", "very_slow_build"])
@pytest.mark.parametrize("name", ["splitty", "slivers", "kitchen", "cornell", "terrain32"])
def test_slow_presets_cwbvh_bit_exact(api, scenes, name, preset):
    tris = scenes[name] if name in scenes else split_scenes()[name]
    want = ob.build_cwbvh_from_tris(tris, preset)
    wnodes, wprims, wtotal = want.get()
    got = api.build_cwbvh_from_tris(tris, api.BvhBuildParams.preset(preset))
    assert got.uses_spatial_splits
    gnodes, gprims, gtotal = got.download()
    assert np.array_equal(gprims, wprims)
    assert gnodes.tobytes() == wnodes.tobytes()
    chk = ob.cwbvh_from(gnodes, gprims, gtotal)
    ob.lib().orc_cwbvh_set_uses_spatial_splits(chk.h, 1)
    rc, msg = chk.validate(ob.tri_aabbs(tris))
    assert rc == 0, msg
    # the handle carries tris[primitive_indices] (longer than tris when something was split) and traces like the oracle
    rays = rays_for(tris, 48)
    wh = want.ray_traverse(want.bvh_tris(tris), rays)
    gh = got.ray_traverse(rays)
    assert np.array_equal(gh["primitive_id"], wh["primitive_id"]) and np.array_equal(gh["t"].view(np.uint32), wh["t"].view(np.uint32))
    got.set_triangles(tris)  # fewer triangles than primitives is fine for a split tree
    assert np.array_equal(got.ray_traverse(rays)["primitive_id"], wh["primitive_id"])


@pytest.mark.parametrize("preset", ["slow_build", "very_slow_build"])
@pytest.mark.parametrize("name", ["splitty", "slivers", "kitchen"])
def test_slow_presets_bvh2_bit_exact(api, scenes, name, preset):
    # bvh2/builder.rs:17-91 with pre_split: splits -> PLOC(U128) -> reinsertion -> collapse (duplicate-primitive rule,
    # leaf_collapser.rs:72-82) -> reinsertion
    tris = scenes[name] if name in scenes else split_scenes()[name]
    want = ob.build_bvh2_from_tris(tris, preset)
    wn, wp = want.get()
    got = api.build_bvh2_from_tris(tris, api.BvhBuildParams.preset(preset))
    assert got.uses_spatial_splits
    gn, gp = got.download()
    assert np.array_equal(gp, wp)
    assert_nodes_equal(gn, wn, f"{name} {preset}")
    rays = rays_for(tris, 48)
    wh = want.ray_traverse(want.bvh_tris(tris), rays)
    gh = got.ray_traverse(rays)
    assert np.array_equal(gh["primitive_id"], wh["primitive_id"]) and np.array_equal(gh["t"].view(np.uint32), wh["t"].view(np.uint32))


def test_kitchen_golden_hash_slow_presets_on_gpu(api, kitchen_tris):
    # examples/obj_cwbvh.rs:174-181 test_slow / test_very_slow through the GPU path only
    rays = camera.primary_rays(camera.kitchen_camera(32))
    for preset in ("slow_build", "very_slow_build"):
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


def test_large_presplit_sequential_sum(api):
    # 1.5M triangles: the one-thread ordered sum still matches the CPU's, and splitting scales past the first arena block
    tris = np.concatenate([tu.triangle_soup(1_500_000, 4), tu.soup_with_large_triangles(0, 2000, 3)], axis=0)
    wa, widx, wavg, wlargest = ob.presplit_tris(tris)
    ga, gidx, gavg, glargest = api.presplit_tris(tris)
    assert gavg.view(np.uint32) == wavg.view(np.uint32) and glargest == wlargest
    assert wa.shape[0] > tris.shape[0] + 5000
    assert np.array_equal(gidx, widx) and np.array_equal(ga[:, LANES].view(np.uint32), wa[:, LANES].view(np.uint32))
