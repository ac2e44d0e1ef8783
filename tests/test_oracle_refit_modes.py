"""Debug- vs release-build semantics of Bvh2::refit_from_fast (reference src/bvh2/mod.rs:722-751).

A release-built reference returns from the upward refit after two ancestors whose recomputed box EQUALS the stored one
(`#[cfg(not(debug_assertions))]`, :732-737) -- before writing that second box back. Equality is by value (-0.0 == +0.0), so the
only thing the early return can change is the SIGN BIT OF A ZERO LANE of some inner-node boxes (and through `CwBvhNode.p`, the raw
node minimum, of the matching CWBVH nodes: cwbvh/node.rs:14-54). The GPU path refits every dirty ancestor (= the debug-assertions
form, what `cargo test` runs). This test quantifies the deviation on the BASELINE configurations that fit the CPU suite and pins
its nature: same topology, same values, differing bits only in zero lanes."""
import numpy as np
import pytest

import oracle_bind as ob
from obvhs_b200 import test_util as tu

PRESETS = {  # (search distance, depth threshold, reinsertion ratio) of lib.rs:233-305
    "fast_build": (6, 2, 0.02),
    "medium_build": (14, 3, 0.05),
}


def _both_modes(tris, preset):
    r, thr, ratio = PRESETS[preset]
    aabbs = ob.tri_aabbs(tris)
    out = []
    for full in (0, 1):
        ob.lib().orc_set_refit_full(full)
        try:
            b = ob.ploc_build(aabbs, None, r, 64, thr)
            applied = b.reinsertion_run(ratio)
            nodes = b.get()[0].copy()
            cw = b.to_cwbvh(3, True).get()[0].copy()
        finally:
            ob.lib().orc_set_refit_full(0)
        out.append((applied, nodes, cw))
    return out


@pytest.mark.parametrize("scene,preset", [("cornell", "medium_build"), ("kitchen", "fast_build"), ("kitchen", "medium_build"),
                                          ("terrain", "fast_build"), ("terrain", "medium_build"), ("soup", "fast_build")])
def test_release_refit_deviation_is_zero_sign_only(scenes, scene, preset, record_property):
    tris = {"cornell": lambda: scenes["cornell"], "kitchen": lambda: scenes["kitchen"], "terrain": lambda: tu.demoscene(160, 0),
            "soup": lambda: tu.triangle_soup(60_000, 1)}[scene]()
    (a_rel, n_rel, c_rel), (a_dbg, n_dbg, c_dbg) = _both_modes(tris, preset)
    assert a_rel == a_dbg  # the same reinsertions are applied
    assert np.array_equal(n_rel["first_index"], n_dbg["first_index"]) and np.array_equal(n_rel["prim_count"], n_dbg["prim_count"])
    assert np.array_equal(n_rel["aabb"], n_dbg["aabb"])  # by value
    bits = n_rel["aabb"].view(np.uint32) != n_dbg["aabb"].view(np.uint32)
    assert np.all(n_rel["aabb"][bits] == 0.0)
    nodes_differ = int(bits.any(axis=1).sum())
    # CWBVH: same node count, and bytes differ only in the sign byte of a zero `p` lane (bytes 3, 7, 11 of the 80)
    assert c_rel.shape == c_dbg.shape
    raw_rel, raw_dbg = c_rel.view(np.uint8).reshape(-1, 80), c_dbg.view(np.uint8).reshape(-1, 80)
    diff = raw_rel != raw_dbg
    assert not diff[:, [i for i in range(80) if i not in (3, 7, 11)]].any()
    assert np.all(c_rel["p"][c_rel["p"].view(np.uint32) != c_dbg["p"].view(np.uint32)] == 0.0)
    cw_nodes_differ, cw_bytes_differ = int(diff.any(axis=1).sum()), int(diff.sum())
    record_property("bvh2_nodes_differ", nodes_differ)
    record_property("cwbvh_bytes_differ", cw_bytes_differ)
    print(f"{scene}/{preset}: {tris.shape[0]} tris, {a_rel} reinsertions: {nodes_differ} of {n_rel.shape[0]} Bvh2 nodes, "
          f"{cw_nodes_differ} of {c_rel.shape[0]} CwBvh nodes ({cw_bytes_differ} bytes) differ between release and debug refit")
    assert nodes_differ <= max(4, n_rel.shape[0] // 1000)
