"""PlocBuilder::full_rebuild / partial_rebuild / compute_rebuild_path_flags (reference src/ploc/rebuild.rs) in the CPU
oracle, restating the reference's own tests (rebuild.rs:186-365: validate(tris, false, true) before and after). CPU only."""
import numpy as np
import pytest

import oracle_bind as ob
from obvhs_b200 import test_util as tu


def scene_list():
    sm = tu.demoscene(5, 0)
    return [tu.demoscene(31, 0), sm, sm[:1], sm[:2], sm[:3], sm[:0]]  # rebuild.rs:199


def build(tris, thr=1):
    aabbs = ob.tri_aabbs(tris) if len(tris) else np.zeros((0, 8), np.float32)
    return ob.ploc_build(aabbs, None, 1, 64, thr), aabbs


def leaf_ids(bvh):
    nodes, _ = bvh.get()
    return np.nonzero(nodes["prim_count"] != 0)[0].astype(np.uint32)


def test_full_rebuild():
    # rebuild.rs:197-217
    for tris in scene_list():
        bvh, aabbs = build(tris)
        assert bvh.validate(aabbs)[0] == 0
        before = bvh.get()
        bvh.full_rebuild(1, 64, 1)
        rc, msg = bvh.validate(aabbs)
        assert rc == 0, msg
        # rebuilding an untouched fresh PLOC tree from its leaves: same leaf set and root box; the leaves now enter in node
        # order instead of primitive order, so only the order among equal Morton codes may change
        after = bvh.get()
        assert np.array_equal(np.sort(after[1]), np.sort(before[1]))
        if len(tris):
            assert np.array_equal(after[0]["aabb"][0], before[0]["aabb"][0])


def test_full_rebuild_after_moving_leaves_equals_fresh_build():
    tris = tu.demoscene(24, 0)
    bvh, aabbs = build(tris, 0)
    ids = leaf_ids(bvh)
    nodes, prims = bvh.get()
    moved = aabbs.copy()
    moved[:, [1, 5]] += np.float32(0.25) * np.sin(np.arange(len(tris), dtype=np.float32))[:, None]
    # give every leaf the new box of its primitive, refit so that the root box is current, then rebuild
    bvh.set_node_aabbs(ids, moved[prims[nodes["first_index"][ids]]])
    bvh.refit_all()
    bvh.full_rebuild(6, 64, 2)
    rc, msg = bvh.validate(moved)
    assert rc == 0, msg
    fresh = ob.ploc_build(moved, None, 6, 64, 2)
    a, b = bvh.get()[0], fresh.get()[0]
    # same scene box, same leaves; topology can only differ through the order of equal Morton codes
    assert np.array_equal(a["aabb"][0], b["aabb"][0])
    assert bvh.node_count == fresh.node_count


@pytest.mark.parametrize("which", ["all", "one", "random"])
def test_partial_rebuild(which):
    # rebuild.rs:248-365
    for tris in scene_list() if which != "one" else [tu.demoscene(8, 0)]:
        bvh, aabbs = build(tris, 0 if which == "one" else 1)
        assert bvh.validate(aabbs)[0] == 0
        if bvh.node_count < 2:
            bvh.partial_rebuild(np.zeros(bvh.node_count, np.uint8), 1, 64, 0)  # returns at once (rebuild.rs:108-110)
            assert bvh.validate(aabbs)[0] == 0
            continue
        bvh.compute_parents()
        ids = leaf_ids(bvh)
        if which == "one":
            ids = ids[:1]
        elif which == "random":
            keep = tu.hash_noise(np.zeros(len(ids), np.uint32), np.zeros(len(ids), np.uint32), ids) > 0.5
            ids = ids[keep][:1]  # `.take(1)` in the reference
        flags = bvh.rebuild_path_flags(ids)
        assert flags[0] == (1 if len(ids) else 0)
        n_before = bvh.node_count
        bvh.partial_rebuild(flags, 1, 64, 0)
        rc, msg = bvh.validate(aabbs)
        assert rc == 0, (which, len(tris), msg)
        assert bvh.node_count == n_before and bvh.has_parents


def test_partial_rebuild_moves_only_flagged_paths():
    tris = tu.demoscene(24, 0)
    bvh, aabbs = build(tris, 0)
    bvh.reinsertion_run(0.1)  # a tree whose children are no longer ordered after their parents
    bvh.compute_parents()
    nodes, prims, parents = bvh.get(with_parents=True)
    ids = leaf_ids(bvh)
    sel = ids[:: max(1, len(ids) // 37)]
    moved = aabbs.copy()
    prim_of = prims[nodes["first_index"][sel]]
    moved[prim_of, 1] -= np.float32(0.5)
    moved[prim_of, 5] += np.float32(0.5)
    bvh.set_node_aabbs(sel, moved[prim_of])
    flags = bvh.rebuild_path_flags(sel)
    bvh.partial_rebuild(flags, 6, 64, 0)
    rc, msg = bvh.validate(moved)
    assert rc == 0, msg
    after = bvh.get()[0]
    # nodes that were not reached by the walk (no flagged parent chain) did not move
    reached = np.zeros(len(nodes), bool)
    reached[0] = True
    order = [0]
    while order:
        v = order.pop()
        if nodes["prim_count"][v] == 0 and (v == 0 or flags[v]):
            for c in (nodes["first_index"][v], nodes["first_index"][v] + 1):
                reached[c] = True
                order.append(int(c))
    assert np.array_equal(after[~reached], nodes[~reached])
    assert (~reached).sum() > 0 and reached.sum() > 2 * len(sel)
