"""GPU parity for PlocBuilder::full_rebuild / partial_rebuild / compute_rebuild_path_flags (reference src/ploc/rebuild.rs)
through the C ABI against the CPU oracle: node arrays bit for bit."""
import numpy as np
import pytest

import oracle_bind as ob
from obvhs_b200 import test_util as tu
from test_gpu_parity import api, assert_nodes_equal, oracle_refit_semantics  # noqa: F401  (fixtures)
from test_oracle_rebuild import leaf_ids, scene_list

pytestmark = pytest.mark.gpu


def both(api, tris, sd=1, thr=1, prec=64, reinsert=0.0, collapse=0):
    aabbs = ob.tri_aabbs(tris) if len(tris) else np.zeros((0, 8), np.float32)
    want = ob.ploc_build(aabbs, None, sd, prec, thr)
    got = api.PlocBuilder().build(sd, aabbs, None, api.SortPrecision(prec), thr)
    if reinsert > 0:
        want.reinsertion_run(reinsert)
        api.ReinsertionOptimizer().run(got, reinsert)
    if collapse:
        want.collapse(collapse, 1.0)
        got.collapse(collapse, 1.0)
    return want, got, aabbs


def same(got, want, what):
    gn, gp = got.download()
    wn, wp = want.get()
    assert np.array_equal(gp, wp), what
    assert_nodes_equal(gn, wn, what)
    assert got.max_depth == want.max_depth


@pytest.mark.parametrize("cfg", [(1, 1, 64), (6, 2, 64), (14, 0, 128)])
def test_full_rebuild_bit_exact(api, cfg):
    sd, thr, prec = cfg
    for tris in scene_list() + [tu.triangle_soup(5000, 2)]:
        want, got, aabbs = both(api, tris, 1, 1)
        want.full_rebuild(sd, prec, thr)
        api.PlocBuilder().full_rebuild(got, sd, api.SortPrecision(prec), thr)
        same(got, want, f"full_rebuild {len(tris)} {cfg}")
        assert got.children_are_ordered_after_parents or got.node_count < 2
        assert ob.bvh2_from(*got.download(), got.max_depth).validate(aabbs)[0] == 0 if len(tris) else True


def test_full_rebuild_of_collapsed_reinserted_tree_with_parents(api, scenes):
    # multi-primitive leaves stay as they are; parents are recomputed because the tree had them (rebuild.rs:147,177-179)
    tris = scenes["kitchen"]
    want, got, aabbs = both(api, tris, 6, 2, reinsert=0.05, collapse=4)
    want.compute_parents()
    got.compute_parents()
    want.full_rebuild(6, 64, 2)
    api.PlocBuilder().full_rebuild(got, 6, api.SortPrecision.U64, 2)
    same(got, want, "full_rebuild collapsed kitchen")
    assert np.array_equal(got.download(with_parents=True)[2], want.get(with_parents=True)[2])


@pytest.mark.parametrize("which", ["all", "one", "random"])
def test_partial_rebuild_reference_cases_bit_exact(api, which):
    # rebuild.rs:248-365
    for tris in (scene_list() if which != "one" else [tu.demoscene(8, 0)]):
        want, got, aabbs = both(api, tris, 1, 0 if which == "one" else 1)
        if want.node_count < 2:
            api.PlocBuilder().partial_rebuild(got, np.zeros(got.node_count, np.uint8), 1)
            same(got, want, "tiny")
            continue
        want.compute_parents()
        got.compute_parents()
        ids = leaf_ids(want)
        if which == "one":
            ids = ids[:1]
        elif which == "random":
            keep = tu.hash_noise(np.zeros(len(ids), np.uint32), np.zeros(len(ids), np.uint32), ids) > 0.5
            ids = ids[keep][:1]
        wflags = want.rebuild_path_flags(ids)
        gflags = api.compute_rebuild_path_flags(got, ids)
        assert np.array_equal(gflags, wflags)
        want.partial_rebuild(wflags, 1, 64, 0)
        api.PlocBuilder().partial_rebuild(got, lambda node_id: bool(gflags[node_id]), api.PlocSearchDistance.Minimum, api.SortPrecision.U64, 0)
        same(got, want, f"partial_rebuild {which} {len(tris)}")
        assert not got.children_are_ordered_after_parents
        assert np.array_equal(got.download(with_parents=True)[2], want.get(with_parents=True)[2])
        assert ob.bvh2_from(*got.download(), got.max_depth).validate(aabbs)[0] == 0


@pytest.mark.parametrize("scene,frac", [("terrain32", 0.03), ("kitchen", 0.01), ("kitchen", 0.4), ("soup4k", 1.0)])
@pytest.mark.parametrize("cfg", [(1, 0, 64), (6, 2, 64), (24, 1, 128)])
def test_partial_rebuild_after_moving_leaves(api, scenes, scene, frac, cfg):
    # examples/physics.rs:431-455: new boxes for some leaves, path flags, partial rebuild -- on a tree that went through
    # reinsertion (children no longer ordered after parents) and whose inner boxes are stale
    sd, thr, prec = cfg
    tris = scenes[scene]
    want, got, aabbs = both(api, tris, 6, 2, reinsert=0.05)
    want.compute_parents()
    got.compute_parents()
    nodes, prims = want.get()
    ids = leaf_ids(want)
    rng = np.random.default_rng(11)
    sel = np.sort(rng.choice(ids, max(1, int(len(ids) * frac)), replace=False)).astype(np.uint32)
    moved = aabbs.copy()
    prim_of = prims[nodes["first_index"][sel]]
    shift = (rng.random((len(sel), 3), dtype=np.float32) - np.float32(0.5)) * np.float32(0.3)
    moved[prim_of, 0:3] += shift
    moved[prim_of, 4:7] += shift
    want.set_node_aabbs(sel, moved[prim_of])
    got.set_node_aabbs(sel, moved[prim_of])
    wflags = want.rebuild_path_flags(sel)
    gflags = api.compute_rebuild_path_flags(got, sel)
    assert np.array_equal(gflags, wflags)
    want.partial_rebuild(wflags, sd, prec, thr)
    api.PlocBuilder().partial_rebuild(got, gflags, sd, api.SortPrecision(prec), thr)
    same(got, want, f"partial_rebuild {scene} {frac} {cfg}")
    # leaves are where the moved boxes say; the root box is whatever PLOC merged (it may be stale only above untouched subtrees)
    rc, msg = ob.bvh2_from(*got.download(), got.max_depth).validate(moved, tight_fit=False)
    assert rc == 0 or "not inside" in msg, msg


def test_partial_rebuild_with_arbitrary_flags_and_without_parents(api, scenes):
    # should_remove is any predicate in the reference, not only ancestor-closed path flags; parents stay None
    tris = scenes["terrain32"]
    want, got, aabbs = both(api, tris, 2, 0)
    rng = np.random.default_rng(5)
    flags = (rng.random(want.node_count) < 0.6).astype(np.uint8)
    want.partial_rebuild(flags, 2, 64, 0)
    api.PlocBuilder().partial_rebuild(got, flags, 2)
    same(got, want, "arbitrary flags")
    assert ob.bvh2_from(*got.download(), got.max_depth).validate(aabbs, tight_fit=False)[0] == 0
    with pytest.raises(api.ObvhsError):
        api.compute_rebuild_path_flags(got, np.array([3], np.uint32))  # parents not computed: the reference panics


def test_rebuild_large(api):
    tris = tu.triangle_soup(300_000, 8)
    want, got, aabbs = both(api, tris, 6, 2)
    want.compute_parents()
    got.compute_parents()
    ids = leaf_ids(want)
    sel = ids[::17]
    flags = want.rebuild_path_flags(sel)
    assert np.array_equal(api.compute_rebuild_path_flags(got, sel), flags)
    want.partial_rebuild(flags, 6, 64, 2, threads=8)
    api.PlocBuilder().partial_rebuild(got, flags, 6, api.SortPrecision.U64, 2)
    same(got, want, "partial large")
    want.full_rebuild(6, 64, 2, threads=8)
    api.PlocBuilder().full_rebuild(got, 6, api.SortPrecision.U64, 2)
    same(got, want, "full large")
