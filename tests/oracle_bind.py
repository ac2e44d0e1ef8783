"""ctypes binding of the CPU oracle (oracle/libobvhs_oracle.so). TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from obvhs_b200.types import BVH2_NODE, CWBVH_NODE, RAY_HIT

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_ORACLE_DIR = os.path.join(_ROOT, "oracle")
_SO = os.path.join(_ORACLE_DIR, "libobvhs_oracle.so")


def build(force: bool = False) -> str:
    src = os.path.join(_ORACLE_DIR, "obvhs_oracle.cpp")
    hdr = os.path.join(_ORACLE_DIR, "obvhs_oracle.h")
    stale = (not os.path.exists(_SO)) or any(os.path.getmtime(p) > os.path.getmtime(_SO) for p in (src, hdr))
    if force or stale:
        subprocess.check_call(["make", "-C", _ORACLE_DIR, "-s"] + (["-B"] if force else []))
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        vp, sz, u32, i32, f32 = C.c_void_p, C.c_size_t, C.c_uint32, C.c_int, C.c_float
        sigs = {
            "orc_tri_aabbs": (None, [vp, sz, vp]),
            "orc_make_rays": (None, [vp, sz, f32, f32, vp]),
            "orc_morton_sort": (None, [vp, sz, i32, vp, vp, vp, vp]),
            "orc_ploc_build": (vp, [vp, vp, sz, u32, i32, sz, i32]),
            "orc_bvh2_free": (None, [vp]),
            "orc_bvh2_node_count": (sz, [vp]),
            "orc_bvh2_prim_count": (sz, [vp]),
            "orc_bvh2_max_depth": (sz, [vp]),
            "orc_bvh2_ploc_iterations": (sz, [vp]),
            "orc_bvh2_get": (None, [vp, vp, vp, vp]),
            "orc_bvh2_from": (vp, [vp, sz, vp, sz, sz]),
            "orc_bvh2_validate": (i32, [vp, vp, sz, i32, C.c_char_p]),
            "orc_bvh2_compute_parents": (None, [vp]),
            "orc_bvh2_refit_all": (None, [vp]),
            "orc_bvh2_reorder_in_stack_traversal_order": (None, [vp]),
            "orc_ploc_full_rebuild": (None, [vp, u32, i32, sz, i32]),
            "orc_ploc_partial_rebuild": (None, [vp, vp, u32, i32, sz, i32]),
            "orc_compute_rebuild_path_flags": (None, [vp, vp, sz, vp]),
            "orc_bvh2_set_node_aabbs": (None, [vp, vp, vp, sz]),
            "orc_bvh2_collapse": (None, [vp, u32, f32]),
            "orc_bvh2_has_parents": (i32, [vp]),
            "orc_bvh2_ray_traverse": (None, [vp, vp, vp, sz, vp, i32, vp]),
            "orc_bvh2_ray_traverse_miss": (None, [vp, vp, vp, sz, vp, i32, vp]),
            "orc_bvh2_ray_traverse_anyhit_count": (None, [vp, vp, vp, sz, vp, i32]),
            "orc_build_bvh2_from_tris": (vp, [vp, sz, u32, sz, f32, f32, i32, u32, f32, i32, i32, vp]),
            "orc_split_aabbs_precise": (sz, [vp, vp, sz, sz, vp, f32, f32, f32, f32, u32, u32]),
            "orc_presplit_tris": (sz, [vp, sz, vp, vp, sz, vp, vp]),
            "orc_bvh2_set_uses_spatial_splits": (None, [vp, i32]),
            "orc_cwbvh_set_uses_spatial_splits": (None, [vp, i32]),
            "orc_bvh2_set_leaf_aabbs": (None, [vp, vp]),
            "orc_reinsertion_run": (None, [vp, f32, vp, sz, i32]),
            "orc_reinsertion_run_with_candidates": (None, [vp, vp, sz, u32, i32]),
            "orc_reinsertion_last_applied": (sz, [vp]),
            "orc_set_refit_full": (None, [i32]),
            "orc_bvh2_to_cwbvh": (vp, [vp, u32, i32, i32]),
            "orc_cwbvh_from": (vp, [vp, sz, vp, sz, vp]),
            "orc_cwbvh_free": (None, [vp]),
            "orc_cwbvh_node_count": (sz, [vp]),
            "orc_cwbvh_prim_count": (sz, [vp]),
            "orc_cwbvh_get": (None, [vp, vp, vp, vp]),
            "orc_cwbvh_validate": (i32, [vp, vp, sz, C.c_char_p]),
            "orc_cwbvh_exact_node_aabbs": (sz, [vp, vp, sz]),
            "orc_build_cwbvh_from_tris": (vp, [vp, sz, u32, sz, f32, i32, u32, i32, i32, vp]),
            "orc_cwbvh_ray_traverse": (None, [vp, vp, vp, sz, vp, i32, i32, vp]),
            "orc_cwbvh_ray_traverse_miss": (None, [vp, vp, vp, sz, vp, i32, i32, vp]),
            "orc_cwbvh_ray_traverse_anyhit_count": (None, [vp, vp, vp, sz, vp, i32]),
            "orc_bvh2_aabb_traverse": (sz, [vp, vp, sz, vp, vp, sz]),
            "orc_bvh2_point_traverse": (sz, [vp, vp, sz, vp, vp, sz]),
            "orc_cwbvh_aabb_traverse": (sz, [vp, vp, sz, vp, vp, vp, sz]),
            "orc_cwbvh_point_traverse": (sz, [vp, vp, sz, vp, vp, vp, sz]),
            "orc_triangle_intersect": (f32, [vp, vp]),
            "orc_triangle_normal": (None, [vp, vp]),
            "orc_cwbvh_order_node_children": (None, [vp, vp, sz, i32]),
            "orc_cwbvh_order_children": (None, [vp, vp, i32]),
            "orc_aabb_half_area": (f32, [vp]),
            "orc_aabb_union": (None, [vp, vp, vp]),
            "orc_aabb_intersect_ray": (f32, [vp, vp]),
            "orc_aabb_intersect_aabb": (i32, [vp, vp]),
            "orc_aabb_contains_point": (i32, [vp, vp]),
            "orc_max_threads": (i32, []),
            "orc_test_split3_64": (C.c_uint64, [u32]),
            "orc_test_split3_128": (None, [C.c_uint64, vp, vp]),
        }
        for name, (res, args) in sigs.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def _f32c(a, cols):
    a = np.ascontiguousarray(a, dtype=np.float32)
    assert a.ndim == 2 and a.shape[1] == cols, a.shape
    return a


def tri_aabbs(tris) -> np.ndarray:
    tris = _f32c(tris, 12)
    out = np.zeros((tris.shape[0], 8), dtype=np.float32)
    lib().orc_tri_aabbs(_p(tris), tris.shape[0], _p(out))
    return out


def make_rays(origin, direction, tmin=0.0, tmax=np.inf) -> np.ndarray:
    """Ray::new (ray.rs:34-52) through the oracle's restatement: (n,16) float32 Ray array, one tmin / tmax for all."""
    od = np.ascontiguousarray(np.concatenate([np.asarray(origin, np.float32), np.asarray(direction, np.float32)], axis=1))
    out = np.zeros((od.shape[0], 16), dtype=np.float32)
    lib().orc_make_rays(_p(od), od.shape[0], float(tmin), float(tmax), _p(out))
    return out


def morton_sort(aabbs, precision=64):
    aabbs = _f32c(aabbs, 8)
    n = aabbs.shape[0]
    lo = np.zeros(n, dtype=np.uint64)
    hi = np.zeros(n, dtype=np.uint64)
    order = np.zeros(n, dtype=np.uint32)
    total = np.zeros(8, dtype=np.float32)
    lib().orc_morton_sort(_p(aabbs), n, precision, _p(lo), _p(hi), _p(order), _p(total))
    return lo, hi, order, total


class Bvh2:
    def __init__(self, handle):
        self.h = handle

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_bvh2_free(self.h)
            self.h = None

    @property
    def node_count(self):
        return lib().orc_bvh2_node_count(self.h)

    @property
    def prim_count(self):
        return lib().orc_bvh2_prim_count(self.h)

    @property
    def max_depth(self):
        return lib().orc_bvh2_max_depth(self.h)

    @property
    def ploc_iterations(self):
        return lib().orc_bvh2_ploc_iterations(self.h)

    def get(self, with_parents=False):
        nodes = np.zeros(self.node_count, dtype=BVH2_NODE)
        prims = np.zeros(self.prim_count, dtype=np.uint32)
        parents = np.zeros(self.node_count, dtype=np.uint32) if with_parents else None
        lib().orc_bvh2_get(self.h, _p(nodes), _p(prims), _p(parents) if with_parents else None)
        return (nodes, prims, parents) if with_parents else (nodes, prims)

    def validate(self, prim_aabbs, tight_fit=True):
        prim_aabbs = _f32c(prim_aabbs, 8)
        msg = C.create_string_buffer(256)
        rc = lib().orc_bvh2_validate(self.h, _p(prim_aabbs), prim_aabbs.shape[0], int(tight_fit), msg)
        return rc, msg.value.decode()

    def compute_parents(self):
        lib().orc_bvh2_compute_parents(self.h)

    def refit_all(self):
        lib().orc_bvh2_refit_all(self.h)

    def reorder_in_stack_traversal_order(self):
        """Bvh2::reorder_in_stack_traversal_order (bvh2/mod.rs:462-500)"""
        lib().orc_bvh2_reorder_in_stack_traversal_order(self.h)

    def set_leaf_aabbs(self, prim_aabbs):
        prim_aabbs = _f32c(prim_aabbs, 8)
        lib().orc_bvh2_set_leaf_aabbs(self.h, _p(prim_aabbs))

    def set_node_aabbs(self, node_ids, aabbs):
        ids = np.ascontiguousarray(node_ids, dtype=np.uint32)
        a = _f32c(aabbs, 8)
        lib().orc_bvh2_set_node_aabbs(self.h, _p(ids), _p(a), ids.shape[0])

    def aabb_traverse(self, queries):
        """Bvh2::aabb_traverse (bvh2/mod.rs:365-407) per query box: (counts, leaf node ids in call order)"""
        return _query(lib().orc_bvh2_aabb_traverse, self.h, queries, 8)

    def point_traverse(self, points):
        """Bvh2::point_traverse (bvh2/mod.rs:414-456)"""
        return _query(lib().orc_bvh2_point_traverse, self.h, points4(points), 4)

    def full_rebuild(self, sd, precision=64, thr=0, threads=1):
        """ploc/rebuild.rs:56-80"""
        lib().orc_ploc_full_rebuild(self.h, sd, precision, thr, threads)

    def partial_rebuild(self, should_remove, sd, precision=64, thr=0, threads=1):
        """ploc/rebuild.rs:101-135"""
        f = np.ascontiguousarray(should_remove, dtype=np.uint8)
        assert f.shape[0] == self.node_count
        lib().orc_ploc_partial_rebuild(self.h, _p(f), sd, precision, thr, threads)

    def rebuild_path_flags(self, leaves):
        """ploc/rebuild.rs:12-43 (parents must be computed)"""
        ids = np.ascontiguousarray(leaves, dtype=np.uint32)
        f = np.zeros(self.node_count, dtype=np.uint8)
        lib().orc_compute_rebuild_path_flags(self.h, _p(ids), ids.shape[0], _p(f))
        return f

    def reinsertion_run(self, ratio, seq=None, threads=1):
        if seq is None:
            lib().orc_reinsertion_run(self.h, ratio, None, 0, threads)
        else:
            s = np.ascontiguousarray(seq, dtype=np.float32)
            lib().orc_reinsertion_run(self.h, ratio, _p(s), s.shape[0], threads)
        return lib().orc_reinsertion_last_applied(self.h)

    def reinsertion_run_with_candidates(self, node_ids, iterations, threads=1):
        ids = np.ascontiguousarray(node_ids, dtype=np.uint32)
        lib().orc_reinsertion_run_with_candidates(self.h, _p(ids), ids.shape[0], int(iterations), threads)
        return lib().orc_reinsertion_last_applied(self.h)

    def collapse(self, max_prims, traversal_cost):
        """bvh2/leaf_collapser.rs:21-192"""
        lib().orc_bvh2_collapse(self.h, int(max_prims), float(traversal_cost))

    @property
    def has_parents(self):
        return bool(lib().orc_bvh2_has_parents(self.h))

    def bvh_tris(self, tris):
        """tris[primitive_indices] (examples/demoscene.rs: triangles reordered so that hit.primitive_id indexes them)"""
        _, prims = self.get()
        return np.ascontiguousarray(_f32c(tris, 12)[prims])

    def ray_traverse(self, bvh_tris, rays, threads=0, counters=None):
        bvh_tris, rays = _f32c(bvh_tris, 12), _f32c(rays, 16)
        hits = np.zeros(rays.shape[0], dtype=RAY_HIT)
        lib().orc_bvh2_ray_traverse(self.h, _p(bvh_tris), _p(rays), rays.shape[0], _p(hits), threads, None if counters is None else _p(counters))
        return hits

    def ray_traverse_miss(self, bvh_tris, rays, threads=0, counters=None):
        bvh_tris, rays = _f32c(bvh_tris, 12), _f32c(rays, 16)
        miss = np.zeros(rays.shape[0], dtype=np.uint8)
        lib().orc_bvh2_ray_traverse_miss(self.h, _p(bvh_tris), _p(rays), rays.shape[0], _p(miss), threads, None if counters is None else _p(counters))
        return miss

    def ray_traverse_anyhit_count(self, bvh_tris, rays, threads=0):
        bvh_tris, rays = _f32c(bvh_tris, 12), _f32c(rays, 16)
        counts = np.zeros(rays.shape[0], dtype=np.uint32)
        lib().orc_bvh2_ray_traverse_anyhit_count(self.h, _p(bvh_tris), _p(rays), rays.shape[0], _p(counts), threads)
        return counts

    def to_cwbvh(self, max_prims_per_leaf=3, order_children=True, include_exact=False):
        return CwBvh(lib().orc_bvh2_to_cwbvh(self.h, max_prims_per_leaf, int(order_children), int(include_exact)))


def _query(fn, handle, queries, cols, extra=()):
    """two calls: count, then fill. -> (counts, ids) with ids grouped by query in call order"""
    q = _f32c(queries, cols)
    n = q.shape[0]
    counts = np.zeros(n, dtype=np.uint32)
    total = fn(handle, _p(q), n, *extra, _p(counts), None, 0)
    ids = np.zeros(max(1, total), dtype=np.uint32)
    fn(handle, _p(q), n, *extra, _p(counts), _p(ids), total)
    return counts, ids[:total]


def points4(points):
    p = np.zeros((len(points), 4), np.float32)
    p[:, :3] = np.asarray(points, dtype=np.float32)[:, :3]
    return p


def bvh2_from(nodes, prims, max_depth=96) -> Bvh2:
    nodes = np.ascontiguousarray(nodes, dtype=BVH2_NODE)
    prims = np.ascontiguousarray(prims, dtype=np.uint32)
    return Bvh2(lib().orc_bvh2_from(_p(nodes), nodes.shape[0], _p(prims), prims.shape[0], max_depth))


def ploc_build(aabbs, indices=None, search_distance=14, precision=64, search_depth_threshold=0, threads=1) -> Bvh2:
    aabbs = _f32c(aabbs, 8)
    n = aabbs.shape[0]
    if indices is None:
        indices = np.arange(n, dtype=np.uint32)
    indices = np.ascontiguousarray(indices, dtype=np.uint32)
    return Bvh2(lib().orc_ploc_build(_p(aabbs), _p(indices), n, search_distance, precision, search_depth_threshold, threads))


class CwBvh:
    def __init__(self, handle):
        self.h = handle

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_cwbvh_free(self.h)
            self.h = None

    @property
    def node_count(self):
        return lib().orc_cwbvh_node_count(self.h)

    @property
    def prim_count(self):
        return lib().orc_cwbvh_prim_count(self.h)

    def get(self):
        nodes = np.zeros(self.node_count, dtype=CWBVH_NODE)
        prims = np.zeros(self.prim_count, dtype=np.uint32)
        total = np.zeros(8, dtype=np.float32)
        lib().orc_cwbvh_get(self.h, _p(nodes), _p(prims), _p(total))
        return nodes, prims, total

    def validate(self, prim_aabbs):
        prim_aabbs = _f32c(prim_aabbs, 8)
        msg = C.create_string_buffer(256)
        rc = lib().orc_cwbvh_validate(self.h, _p(prim_aabbs), prim_aabbs.shape[0], msg)
        return rc, msg.value.decode()

    def bvh_tris(self, tris):
        """examples/obj_cwbvh.rs:63-67: triangles permuted by primitive_indices."""
        _, prims, _ = self.get()
        return np.ascontiguousarray(np.asarray(tris, dtype=np.float32)[prims])

    def order_children(self, prim_aabbs, direct_layout=False):
        """CwBvh::order_children (cwbvh/mod.rs:520-524)"""
        lib().orc_cwbvh_order_children(self.h, _p(_f32c(prim_aabbs, 8)), int(direct_layout))

    def order_node_children(self, prim_aabbs, node_index, direct_layout=False):
        """CwBvh::order_node_children (cwbvh/mod.rs:538-735)"""
        lib().orc_cwbvh_order_node_children(self.h, _p(_f32c(prim_aabbs, 8)), int(node_index), int(direct_layout))

    def exact_node_aabbs(self):
        """CwBvh::exact_node_aabbs (None when the tree was converted without them)"""
        n = lib().orc_cwbvh_exact_node_aabbs(self.h, None, 0)
        if n == 0:
            return None
        out = np.zeros((n, 8), np.float32)
        lib().orc_cwbvh_exact_node_aabbs(self.h, _p(out), n)
        return out

    def aabb_traverse(self, queries, direction=(0.0, 0.0, 0.0)):
        """traverse!(.., node.intersect_aabb(&aabb, state.oct_inv4), ..) per query: (counts, primitive slots in call order)"""
        d = np.ascontiguousarray(direction, dtype=np.float32)
        return _query(lib().orc_cwbvh_aabb_traverse, self.h, queries, 8, (_p(d),))

    def point_traverse(self, points, direction=(0.0, 0.0, 0.0)):
        d = np.ascontiguousarray(direction, dtype=np.float32)
        return _query(lib().orc_cwbvh_point_traverse, self.h, points4(points), 4, (_p(d),))

    def ray_traverse(self, bvh_tris, rays, threads=0, use_simd=True, counters=None):
        bvh_tris = _f32c(bvh_tris, 12) if len(bvh_tris) else np.zeros((0, 12), np.float32)
        rays = _f32c(rays, 16)
        hits = np.zeros(rays.shape[0], dtype=RAY_HIT)
        lib().orc_cwbvh_ray_traverse(self.h, _p(bvh_tris), _p(rays), rays.shape[0], _p(hits), threads, int(use_simd),
                                     _p(counters) if counters is not None else None)
        return hits

    def ray_traverse_miss(self, bvh_tris, rays, threads=0, use_simd=True, counters=None):
        bvh_tris = _f32c(bvh_tris, 12) if len(bvh_tris) else np.zeros((0, 12), np.float32)
        rays = _f32c(rays, 16)
        miss = np.zeros(rays.shape[0], dtype=np.uint8)
        lib().orc_cwbvh_ray_traverse_miss(self.h, _p(bvh_tris), _p(rays), rays.shape[0], _p(miss), threads, int(use_simd),
                                          _p(counters) if counters is not None else None)
        return miss

    def ray_traverse_anyhit_count(self, bvh_tris, rays, threads=0):
        bvh_tris = _f32c(bvh_tris, 12)
        rays = _f32c(rays, 16)
        counts = np.zeros(rays.shape[0], dtype=np.uint32)
        lib().orc_cwbvh_ray_traverse_anyhit_count(self.h, _p(bvh_tris), _p(rays), rays.shape[0], _p(counts), threads)
        return counts


def cwbvh_from(nodes, prims, total_aabb=None) -> CwBvh:
    nodes = np.ascontiguousarray(nodes, dtype=CWBVH_NODE)
    prims = np.ascontiguousarray(prims, dtype=np.uint32)
    total = np.ascontiguousarray(total_aabb if total_aabb is not None else np.zeros(8), dtype=np.float32)
    return CwBvh(lib().orc_cwbvh_from(_p(nodes), nodes.shape[0], _p(prims), prims.shape[0], _p(total)))


# BvhBuildParams presets (reference src/lib.rs:233-305) as (search_distance, depth_threshold, reinsertion_ratio,
# precision, max_prims_per_leaf[, pre_split])
PRESETS = {
    "fastest_build": (1, 0, 0.0, 64, 1),
    "very_fast_build": (1, 0, 0.01, 64, 8),
    "fast_build": (6, 2, 0.02, 64, 8),
    "medium_build": (14, 3, 0.05, 64, 8),
    "slow_build": (24, 2, 0.2, 128, 8, 1),
    "very_slow_build": (14, 1, 1.0, 128, 8, 1),
}


def split_aabbs_precise(aabbs, indices, tris, lo, hi, f_lo, f_hi, max_iterations, split_tests):
    """splits.rs:49-125; returns the grown (aabbs, indices)"""
    aabbs = _f32c(aabbs, 8)
    tris = _f32c(tris, 12)
    n = aabbs.shape[0]
    cap = max(2 * n, 1024)
    while True:
        a = np.zeros((cap, 8), np.float32)
        idx = np.zeros(cap, np.uint32)
        a[:n] = aabbs
        idx[:n] = np.asarray(indices, dtype=np.uint32)
        m = lib().orc_split_aabbs_precise(_p(a), _p(idx), n, cap, _p(tris), lo, hi, f_lo, f_hi, max_iterations, split_tests)
        if m <= cap:
            return a[:m].copy(), idx[:m].copy()
        cap = m


def presplit_tris(tris):
    """cwbvh/builder.rs:27-54 prologue + split_aabbs_preset: returns (aabbs, indices, avg_half_area, largest_half_area)"""
    tris = _f32c(tris, 12)
    n = tris.shape[0]
    cap = max(2 * n, 1024)
    avg, largest = C.c_float(0), C.c_float(0)
    while True:
        a = np.zeros((cap, 8), np.float32)
        idx = np.zeros(cap, np.uint32)
        m = lib().orc_presplit_tris(_p(tris), n, _p(a), _p(idx), cap, C.byref(avg), C.byref(largest))
        if m <= cap:
            return a[:m].copy(), idx[:m].copy(), np.float32(avg.value), np.float32(largest.value)
        cap = m


def build_cwbvh_from_tris(tris, preset="medium_build", threads=1):
    tris = _f32c(tris, 12) if len(tris) else np.zeros((0, 12), np.float32)
    cfg = PRESETS[preset] if isinstance(preset, str) else preset
    sd, thr, ratio, prec, mp = cfg[:5]
    pre_split = int(cfg[5]) if len(cfg) > 5 else 0
    secs = C.c_double(0.0)
    h = lib().orc_build_cwbvh_from_tris(_p(tris), tris.shape[0], sd, thr, ratio, prec, mp, pre_split, threads, C.byref(secs))
    c = CwBvh(h)
    c.core_build_seconds = secs.value
    return c


# (search distance, depth threshold, reinsertion ratio, post-collapse multiplier, precision, max prims/leaf, collapse cost)
BVH2_PRESETS = {
    "fastest_build": (1, 0, 0.0, 0.0, 64, 1, 1.0),
    "very_fast_build": (1, 0, 0.01, 0.0, 64, 8, 3.0),
    "fast_build": (6, 2, 0.02, 0.0, 64, 8, 3.0),
    "medium_build": (14, 3, 0.05, 2.0, 64, 8, 3.0),
    "slow_build": (24, 2, 0.2, 2.0, 128, 8, 3.0, 1),
    "very_slow_build": (14, 1, 1.0, 1.0, 128, 8, 3.0, 1),
}


def build_bvh2_from_tris(tris, preset="medium_build", threads=1) -> Bvh2:
    """bvh2/builder.rs:17-91"""
    tris = _f32c(tris, 12) if len(tris) else np.zeros((0, 12), np.float32)
    cfg = BVH2_PRESETS[preset] if isinstance(preset, str) else preset
    sd, thr, ratio, mult, prec, mp, cost = cfg[:7]
    pre_split = int(cfg[7]) if len(cfg) > 7 else 0
    secs = C.c_double(0.0)
    b = Bvh2(lib().orc_build_bvh2_from_tris(_p(tris), tris.shape[0], sd, thr, ratio, mult, prec, mp, cost, pre_split, threads, C.byref(secs)))
    b.core_build_seconds = secs.value
    return b


def triangle_normals(tris) -> np.ndarray:
    """`Triangle::compute_normal` (reference src/triangle.rs:20-24), numpy float32 restatement."""
    t = np.asarray(tris, dtype=np.float32)
    e1 = t[:, 4:7] - t[:, 0:3]
    e2 = t[:, 8:11] - t[:, 0:3]
    c = np.stack(
        [
            e1[:, 1] * e2[:, 2] - e2[:, 1] * e1[:, 2],
            e1[:, 2] * e2[:, 0] - e2[:, 2] * e1[:, 0],
            e1[:, 0] * e2[:, 1] - e2[:, 0] * e1[:, 1],
        ],
        axis=1,
    ).astype(np.float32)
    d = (c[:, 0] * c[:, 0] + c[:, 1] * c[:, 1]) + c[:, 2] * c[:, 2]
    with np.errstate(divide="ignore", invalid="ignore"):
        rcp = np.float32(1.0) / np.sqrt(d)
    ok = np.isfinite(rcp) & (rcp > 0)
    return np.where(ok[:, None], c * rcp[:, None], np.float32(0.0)).astype(np.float32)
