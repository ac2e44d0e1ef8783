"""Spatial pre-splits (reference src/splits.rs) in the CPU oracle: properties every split set must have, and the two
presets that use them (slow_build / very_slow_build, src/lib.rs:282-305). The kitchen golden hash for those presets is in
test_oracle_golden.py. CPU only."""
import numpy as np
import pytest

import oracle_bind as ob
from obvhs_b200 import test_util as tu


def split_scenes():
    return {
        "splitty": tu.soup_with_large_triangles(3000, 40, 5),
        "slivers": tu.soup_with_large_triangles(500, 300, 9),
        "terrain_plus": np.concatenate([tu.demoscene(24, 0), tu.soup_with_large_triangles(0, 12, 2) * np.float32(8.0)], axis=0),
    }


@pytest.mark.parametrize("name", ["splitty", "slivers", "terrain_plus"])
def test_split_sets_partition_their_triangles(name):
    tris = split_scenes()[name]
    n = tris.shape[0]
    aabbs, idx, avg, largest = ob.presplit_tris(tris)
    base = ob.tri_aabbs(tris)
    assert aabbs.shape[0] > n, "the scene is meant to split"
    assert np.array_equal(idx[:n], np.arange(n, dtype=np.uint32))  # originals stay in place, right halves are appended
    assert idx.max() < n
    # the sequential f32 average and the max of the half areas (cwbvh/builder.rs:29-43)
    d = base[:, 4:7] - base[:, 0:3]
    half = (d[:, 0] + d[:, 1]) * d[:, 2] + d[:, 0] * d[:, 1]
    acc = np.float32(0.0)
    for h in half:
        acc = np.float32(acc + h)
    assert avg == np.float32(acc / np.float32(n)) and largest == half.max()
    # every piece lies inside its triangle's AABB, and the pieces of a triangle together cover it exactly
    assert np.all(aabbs[:, 0:3] >= base[idx, 0:3]) and np.all(aabbs[:, 4:7] <= base[idx, 4:7])
    lo = np.full((n, 3), np.inf, np.float32)
    hi = np.full((n, 3), -np.inf, np.float32)
    np.minimum.at(lo, idx, aabbs[:, 0:3])
    np.maximum.at(hi, idx, aabbs[:, 4:7])
    assert np.array_equal(lo, base[:, 0:3]) and np.array_equal(hi, base[:, 4:7])
    # points on a split triangle are inside at least one of its pieces
    rng = np.random.default_rng(3)
    split_ids = np.nonzero(np.bincount(idx, minlength=n) > 1)[0]
    for t in split_ids[:40]:
        pieces = aabbs[idx == t]
        w = rng.dirichlet((1.0, 1.0, 1.0), 64).astype(np.float32)
        pts = w[:, :1] * tris[t, 0:3] + w[:, 1:2] * tris[t, 4:7] + w[:, 2:3] * tris[t, 8:11]
        eps = np.float32(1e-5)
        inside = ((pts[:, None, :] >= pieces[None, :, 0:3] - eps) & (pts[:, None, :] <= pieces[None, :, 4:7] + eps)).all(axis=2).any(axis=1)
        assert inside.all(), t


def test_presets_without_large_triangles_do_not_split():
    for tris in (tu.cornell_box(), tu.demoscene(32, 0), tu.flat_plane(4)):
        aabbs, idx, _, _ = ob.presplit_tris(tris)
        assert aabbs.shape[0] == tris.shape[0]
        assert np.array_equal(aabbs[:, [0, 1, 2, 4, 5, 6]], ob.tri_aabbs(tris)[:, [0, 1, 2, 4, 5, 6]])


def test_precise_matches_preset_parameters():
    # split_aabbs_preset is split_aabbs_precise(avg*3, max(avg*4, avg*0.9 + largest*0.1), 1.8, 1.6, 12, 12) (splits.rs:23-33)
    tris = split_scenes()["splitty"]
    a, idx, avg, largest = ob.presplit_tris(tris)
    n = tris.shape[0]
    hi = max(np.float32(avg * np.float32(4.0)), np.float32(np.float32(avg * np.float32(0.9)) + np.float32(largest * np.float32(0.1))))
    b, idx2 = ob.split_aabbs_precise(ob.tri_aabbs(tris), np.arange(n, dtype=np.uint32), tris, np.float32(avg * np.float32(3.0)), hi,
                                     1.8, 1.6, 12, 12)
    assert np.array_equal(a, b) and np.array_equal(idx, idx2)
    # fewer iterations split less; zero iterations split nothing
    c, _ = ob.split_aabbs_precise(ob.tri_aabbs(tris), np.arange(n, dtype=np.uint32), tris, np.float32(avg * np.float32(3.0)), hi, 1.8, 1.6, 1, 12)
    assert n < c.shape[0] < a.shape[0]
    z, _ = ob.split_aabbs_precise(ob.tri_aabbs(tris), np.arange(n, dtype=np.uint32), tris, np.float32(avg * np.float32(3.0)), hi, 1.8, 1.6, 0, 12)
    assert z.shape[0] == n


@pytest.mark.parametrize("preset", ["slow_build", "very_slow_build"])
@pytest.mark.parametrize("name", ["splitty", "slivers"])
def test_split_builds_validate_and_trace_like_unsplit_builds(name, preset):
    tris = split_scenes()[name]
    aabbs = ob.tri_aabbs(tris)
    c = ob.build_cwbvh_from_tris(tris, preset)
    assert c.prim_count > tris.shape[0]
    rc, msg = c.validate(aabbs)
    assert rc == 0, msg
    b = ob.build_bvh2_from_tris(tris, preset)
    rc, msg = b.validate(aabbs, tight_fit=False)
    assert rc == 0, msg
    # closest hits do not depend on how the tree was built: same distances as the medium (unsplit) build
    from test_gpu_parity import rays_for

    rays = rays_for(tris, 48)
    ref = ob.build_cwbvh_from_tris(tris, "medium_build")
    want = ref.ray_traverse(ref.bvh_tris(tris), rays)
    got = c.ray_traverse(c.bvh_tris(tris), rays)
    assert np.array_equal(got["t"], want["t"])
    _, prims, _ = c.get()
    _, rprims, _ = ref.get()
    hit = np.isfinite(want["t"])
    same = prims[got["primitive_id"][hit]] == rprims[want["primitive_id"][hit]]
    assert same.mean() > 0.999  # exact-t ties between different triangles may resolve differently
    got2 = b.ray_traverse(b.bvh_tris(tris), rays)  # Bvh2::ray_traverse leaves hit.t = ray.tmax on a miss
    assert np.array_equal(got2["t"][hit], want["t"][hit]) and np.array_equal(got2["t"] < rays[:, 13], hit)
