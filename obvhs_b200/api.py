"""Host-side mirror of the reference's API for the hot path, over the C ABI of libobvhs_cuda (include/obvhs_cuda.h).

Names, argument meaning and error behaviour follow obvhs 0.3.1:

* ``PlocBuilder.build(search_distance, aabbs, indices, sort_precision, search_depth_threshold) -> Bvh2``
  (src/ploc/mod.rs:95-102), ``PlocSearchDistance`` (:534-562), ``SortPrecision`` (:658-661)
* ``ReinsertionOptimizer().run(bvh, batch_size_ratio, ratio_sequence)`` (src/bvh2/reinsertion.rs:40-57)
* ``bvh2_to_cwbvh(bvh2, max_prims_per_leaf, order_children, include_exact_node_aabbs) -> CwBvh``
  (src/cwbvh/bvh2_to_cwbvh.rs:490-510)
* ``build_cwbvh_from_tris(triangles, config, core_build_time) -> CwBvh`` (src/cwbvh/builder.rs:20-85),
  ``BvhBuildParams`` + presets (src/lib.rs:208-305)
* ``CwBvh.ray_traverse / ray_traverse_miss`` batched over an array of rays (src/cwbvh/mod.rs:169-225)

Arrays may be numpy arrays (host) or torch CUDA tensors (device, used in place); results come back as the same
kind as the `out=` argument, numpy by default. There is NO CPU fallback: importing works anywhere, but any call
needs the compiled library and a CUDA device and raises otherwise.
"""
from __future__ import annotations

import ctypes as C
import enum
import math
import os
from dataclasses import dataclass

import numpy as np

from .types import BVH2_NODE, CWBVH_NODE, RAY_HIT, RAY_HIT8

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libobvhs_cuda.so")


class ObvhsError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libobvhs_cuda error {code}: {msg}")
        self.code = code


class BuildParamsC(C.Structure):
    _fields_ = [
        ("pre_split", C.c_uint32),
        ("ploc_search_distance", C.c_uint32),
        ("search_depth_threshold", C.c_uint64),
        ("reinsertion_batch_ratio", C.c_float),
        ("post_collapse_reinsertion_batch_ratio_multiplier", C.c_float),
        ("sort_precision", C.c_uint32),
        ("max_prims_per_leaf", C.c_uint32),
        ("collapse_traversal_cost", C.c_float),
    ]


_lib = None

# every symbol include/obvhs_cuda.h declares: name -> (restype, argtypes)
_vp, _sz, _u32, _i32, _f32, _u64 = C.c_void_p, C.c_size_t, C.c_uint32, C.c_int, C.c_float, C.c_uint64
_PP = C.POINTER(C.c_void_p)
SIGNATURES = {
    "obvhs_cuda_create": (_i32, [_i32, _vp, _PP]),
    "obvhs_cuda_destroy": (None, [_vp]),
    "obvhs_cuda_last_error": (C.c_char_p, [_vp]),
    "obvhs_cuda_synchronize": (_i32, [_vp]),
    "obvhs_cuda_launch_count": (_u64, [_vp]),
    "obvhs_cuda_set_option": (_i32, [_vp, C.c_char_p, C.c_char_p]),
    "obvhs_cuda_build_params_preset": (_i32, [C.c_char_p, C.POINTER(BuildParamsC)]),
    "obvhs_cuda_split_aabbs_precise": (_i32, [_vp, _vp, _vp, _sz, _sz, _vp, _sz, _f32, _f32, _f32, _f32, _u32, _u32, C.POINTER(_sz)]),
    "obvhs_cuda_split_aabbs_preset": (_i32, [_vp, _vp, _vp, _sz, _sz, _vp, _sz, _f32, _f32, C.POINTER(_sz)]),
    "obvhs_cuda_presplit_tris": (_i32, [_vp, _vp, _sz, _vp, _vp, _sz, C.POINTER(_sz), C.POINTER(_f32), C.POINTER(_f32)]),
    "obvhs_cuda_bvh2_uses_spatial_splits": (_i32, [_vp]),
    "obvhs_cuda_bvh2_set_uses_spatial_splits": (None, [_vp, _i32]),
    "obvhs_cuda_cwbvh_uses_spatial_splits": (_i32, [_vp]),
    "obvhs_cuda_cwbvh_set_uses_spatial_splits": (None, [_vp, _i32]),
    "obvhs_cuda_morton_sort": (_i32, [_vp, _vp, _sz, _u32, _vp, _vp, _vp, _vp]),
    "obvhs_cuda_ploc_build": (_i32, [_vp, _vp, _vp, _sz, _u32, _u32, _sz, _PP]),
    "obvhs_cuda_ploc_build_tris": (_i32, [_vp, _vp, _sz, _u32, _u32, _sz, _PP]),
    "obvhs_cuda_ploc_full_rebuild": (_i32, [_vp, _vp, _u32, _u32, _sz]),
    "obvhs_cuda_ploc_partial_rebuild": (_i32, [_vp, _vp, _vp, _u32, _u32, _sz]),
    "obvhs_cuda_compute_rebuild_path_flags": (_i32, [_vp, _vp, _vp, _sz, _vp]),
    "obvhs_cuda_bvh2_set_node_aabbs": (_i32, [_vp, _vp, _vp, _vp, _sz]),
    "obvhs_cuda_bvh2_free": (None, [_vp]),
    "obvhs_cuda_bvh2_node_count": (_sz, [_vp]),
    "obvhs_cuda_bvh2_prim_count": (_sz, [_vp]),
    "obvhs_cuda_bvh2_max_depth": (_sz, [_vp]),
    "obvhs_cuda_bvh2_ploc_iterations": (_sz, [_vp]),
    "obvhs_cuda_bvh2_children_ordered_after_parents": (_i32, [_vp]),
    "obvhs_cuda_bvh2_download": (_i32, [_vp, _vp, _vp, _vp, _vp]),
    "obvhs_cuda_bvh2_upload": (_i32, [_vp, _vp, _sz, _vp, _sz, _sz, _i32, _PP]),
    "obvhs_cuda_bvh2_compute_parents": (_i32, [_vp, _vp]),
    "obvhs_cuda_bvh2_refit_all": (_i32, [_vp, _vp]),
    "obvhs_cuda_bvh2_reorder_in_stack_traversal_order": (_i32, [_vp, _vp]),
    "obvhs_cuda_bvh2_set_leaf_aabbs": (_i32, [_vp, _vp, _vp, _sz]),
    "obvhs_cuda_reinsertion_run": (_i32, [_vp, _vp, _f32, _vp, _sz, C.POINTER(_u64)]),
    "obvhs_cuda_reinsertion_run_with_candidates": (_i32, [_vp, _vp, _vp, _sz, _u32, C.POINTER(_u64)]),
    "obvhs_cuda_bvh2_collapse": (_i32, [_vp, _vp, _u32, _f32]),
    "obvhs_cuda_build_bvh2_from_tris": (_i32, [_vp, _vp, _sz, C.POINTER(BuildParamsC), C.POINTER(C.c_double), _PP]),
    "obvhs_cuda_bvh2_set_triangles": (_i32, [_vp, _vp, _vp, _sz]),
    "obvhs_cuda_bvh2_ray_traverse_batch": (_i32, [_vp, _vp, _vp, _sz, _vp]),
    "obvhs_cuda_bvh2_ray_traverse_miss_batch": (_i32, [_vp, _vp, _vp, _sz, _vp]),
    "obvhs_cuda_bvh2_ray_traverse_anyhit_count_batch": (_i32, [_vp, _vp, _vp, _sz, _vp]),
    "obvhs_cuda_bvh2_ray_traverse_batch_counted": (_i32, [_vp, _vp, _vp, _sz, _vp, _vp]),
    "obvhs_cuda_bvh2_to_cwbvh": (_i32, [_vp, _vp, _u32, _i32, _i32, _PP]),
    "obvhs_cuda_build_cwbvh": (_i32, [_vp, _vp, _sz, C.POINTER(BuildParamsC), C.POINTER(C.c_double), _PP]),
    "obvhs_cuda_build_bvh2": (_i32, [_vp, _vp, _sz, C.POINTER(BuildParamsC), C.POINTER(C.c_double), _PP]),
    "obvhs_cuda_build_cwbvh_from_tris": (_i32, [_vp, _vp, _sz, C.POINTER(BuildParamsC), C.POINTER(C.c_double), _PP]),
    "obvhs_cuda_cwbvh_free": (None, [_vp]),
    "obvhs_cuda_cwbvh_node_count": (_sz, [_vp]),
    "obvhs_cuda_cwbvh_prim_count": (_sz, [_vp]),
    "obvhs_cuda_cwbvh_exact_node_aabbs": (_i32, [_vp, _vp, _vp, _sz, C.POINTER(_sz)]),
    "obvhs_cuda_cwbvh_compute_parents": (_i32, [_vp, _vp, _vp]),
    "obvhs_cuda_cwbvh_order_children": (_i32, [_vp, _vp, _vp, _sz, _i32]),
    "obvhs_cuda_cwbvh_download": (_i32, [_vp, _vp, _vp, _vp, _vp]),
    "obvhs_cuda_cwbvh_upload": (_i32, [_vp, _vp, _sz, _vp, _sz, _vp, _PP]),
    "obvhs_cuda_cwbvh_set_triangles": (_i32, [_vp, _vp, _vp, _sz]),
    "obvhs_cuda_cwbvh_triangle_bytes": (_sz, []),
    "obvhs_cuda_cwbvh_device_ptrs": (_i32, [_vp, _PP, _PP, _PP]),
    "obvhs_cuda_cwbvh_alloc": (_i32, [_vp, _sz, _sz, _i32, _vp, _PP]),
    "obvhs_cuda_cwbvh_ray_traverse_batch": (_i32, [_vp, _vp, _vp, _sz, _vp]),
    "obvhs_cuda_cwbvh_ray_traverse_miss_batch": (_i32, [_vp, _vp, _vp, _sz, _vp]),
    "obvhs_cuda_cwbvh_ray_traverse_anyhit_count_batch": (_i32, [_vp, _vp, _vp, _sz, _vp]),
    "obvhs_cuda_cwbvh_ray_traverse_batch_counted": (_i32, [_vp, _vp, _vp, _sz, _vp, _vp]),
    "obvhs_cuda_cwbvh_ray_new_traverse_batch_counted": (_i32, [_vp, _vp, _vp, _sz, _vp, _vp]),
    "obvhs_cuda_nccl_unique_id": (_i32, [_vp]),
    "obvhs_cuda_comm_init": (_i32, [_vp, _vp, _i32, _i32]),
    "obvhs_cuda_cwbvh_broadcast": (_i32, [_vp, _PP, _i32]),
    "obvhs_cuda_ray_new_batch": (_i32, [_vp, _vp, _sz, _vp]),
    "obvhs_cuda_cwbvh_ray_new_traverse_batch": (_i32, [_vp, _vp, _vp, _sz, _vp]),
    "obvhs_cuda_cwbvh_ray_new_traverse_miss_batch": (_i32, [_vp, _vp, _vp, _sz, _vp]),
    "obvhs_cuda_cwbvh_ray_new_traverse_anyhit_count_batch": (_i32, [_vp, _vp, _vp, _sz, _vp]),
    "obvhs_cuda_bvh2_ray_new_traverse_batch": (_i32, [_vp, _vp, _vp, _sz, _vp]),
    "obvhs_cuda_bvh2_ray_new_traverse_miss_batch": (_i32, [_vp, _vp, _vp, _sz, _vp]),
    "obvhs_cuda_cwbvh_ray_od_traverse_batch": (_i32, [_vp, _vp, _vp, _sz, _f32, _f32, _vp]),
    "obvhs_cuda_cwbvh_ray_od_traverse_miss_batch": (_i32, [_vp, _vp, _vp, _sz, _f32, _f32, _vp]),
    "obvhs_cuda_cwbvh_ray_od_traverse_hit8_batch": (_i32, [_vp, _vp, _vp, _sz, _f32, _f32, _vp]),
    "obvhs_cuda_bvh2_ray_od_traverse_batch": (_i32, [_vp, _vp, _vp, _sz, _f32, _f32, _vp]),
    "obvhs_cuda_bvh2_aabb_traverse_batch": (_i32, [_vp, _vp, _vp, _sz, _vp, _vp, _sz, C.POINTER(_sz)]),
    "obvhs_cuda_bvh2_point_traverse_batch": (_i32, [_vp, _vp, _vp, _sz, _vp, _vp, _sz, C.POINTER(_sz)]),
    "obvhs_cuda_cwbvh_aabb_traverse_batch": (_i32, [_vp, _vp, _vp, _sz, _vp, _vp, _vp, _sz, C.POINTER(_sz)]),
    "obvhs_cuda_cwbvh_point_traverse_batch": (_i32, [_vp, _vp, _vp, _sz, _vp, _vp, _vp, _sz, C.POINTER(_sz)]),
    "obvhs_cuda_make_rays": (_i32, [_vp, _vp, _sz, _f32, _f32, _vp]),
}


def load_library(path: str = LIB_PATH):
    """dlopen the C-ABI library and bind every declared symbol. Raises if it is missing: there is no fallback.
    OBVHS_LIB_PATH selects another build of the same library (tuning variants, obvhs_b200/build.py: build_variant)."""
    global _lib
    if _lib is None:
        path = os.environ.get("OBVHS_LIB_PATH", path)
        if not os.path.exists(path):
            raise ObvhsError(-2, f"{path} not built (run `python -m obvhs_b200.build`); there is no CPU fallback")
        lib = C.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def _is_torch(x) -> bool:
    return hasattr(x, "data_ptr") and hasattr(x, "is_cuda")


def _ptr(x):
    if x is None:
        return None
    if _is_torch(x):
        assert x.is_contiguous(), "tensors passed to libobvhs_cuda must be contiguous"
        return C.c_void_p(x.data_ptr())
    return x.ctypes.data_as(C.c_void_p)


def _as_f32(x, cols):
    """numpy -> contiguous float32 (n, cols); torch tensors are checked and passed through."""
    if _is_torch(x):
        assert x.dim() == 2 and x.shape[1] == cols and str(x.dtype) == "torch.float32", (x.shape, x.dtype)
        return x.contiguous()
    a = np.ascontiguousarray(x, dtype=np.float32)
    if a.size == 0:
        a = a.reshape(0, cols)
    assert a.ndim == 2 and a.shape[1] == cols, a.shape
    return a


def _as_rays(x):
    """(n,16) float32 = Ray structs, (n,8) float32 = Ray::new arguments (types.make_ray_args) -> (array, is_args)."""
    cols = int(x.shape[1]) if getattr(x, "ndim", 0) == 2 or (_is_torch(x) and x.dim() == 2) else 16
    if cols == 8:
        return _as_f32(x, 8), True
    return _as_f32(x, 16), False


class Context:
    """One device + stream + reusable builder scratch (obvhs_cuda_create)."""

    def __init__(self, device: int = 0, stream=None, traverse: str | None = None):
        self.lib = load_library()
        h = C.c_void_p()
        rc = self.lib.obvhs_cuda_create(device, C.c_void_p(stream) if stream else None, C.byref(h))
        if rc != 0:
            raise ObvhsError(rc, "obvhs_cuda_create failed (no CUDA device? there is no CPU fallback)")
        self.h = h
        self.device = device
        if traverse is not None:
            self.set_option("traverse", traverse)

    def set_option(self, key: str, value: str):
        """obvhs_cuda_set_option: tuning knobs that never change a result (traversal kernel choice, stage tracing)."""
        self.check(self.lib.obvhs_cuda_set_option(self.h, key.encode(), str(value).encode()))

    def check(self, rc: int):
        if rc != 0:
            raise ObvhsError(rc, (self.lib.obvhs_cuda_last_error(self.h) or b"").decode())

    def synchronize(self):
        self.check(self.lib.obvhs_cuda_synchronize(self.h))

    def comm_init(self, unique_id: bytes, rank: int, world: int):
        """obvhs_cuda_comm_init: bind an NCCL communicator over `world` ranks to this context (collective). `unique_id` is the
        128-byte id `nccl_unique_id()` returned on one rank, shipped to the others by the caller."""
        buf = (C.c_uint8 * 128).from_buffer_copy(bytes(unique_id))
        self.check(self.lib.obvhs_cuda_comm_init(self.h, buf, int(rank), int(world)))
        self.comm_rank, self.comm_world = int(rank), int(world)

    @property
    def launch_count(self) -> int:
        return int(self.lib.obvhs_cuda_launch_count(self.h))

    def close(self):
        if getattr(self, "h", None):
            self.lib.obvhs_cuda_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def nccl_unique_id() -> bytes:
    """obvhs_cuda_nccl_unique_id (call on ONE rank, then ship the bytes to the others)."""
    buf = (C.c_uint8 * 128)()
    rc = load_library().obvhs_cuda_nccl_unique_id(buf)
    if rc != 0:
        raise ObvhsError(rc, "NCCL is not available (libnccl.so.2 not found)")
    return bytes(buf)


_default_ctx = {}


def default_context(device: int = 0) -> Context:
    if device not in _default_ctx:
        _default_ctx[device] = Context(device)
    return _default_ctx[device]


class PlocSearchDistance(enum.IntEnum):
    """src/ploc/mod.rs:534-548"""

    Minimum = 1
    VeryLow = 2
    Low = 6
    Medium = 14
    High = 24
    VeryHigh = 32


class SortPrecision(enum.IntEnum):
    """src/ploc/mod.rs:658-661"""

    U128 = 128
    U64 = 64


@dataclass
class BvhBuildParams:
    """src/lib.rs:208-231"""

    pre_split: bool
    ploc_search_distance: PlocSearchDistance
    search_depth_threshold: int
    reinsertion_batch_ratio: float
    post_collapse_reinsertion_batch_ratio_multiplier: float
    sort_precision: SortPrecision
    max_prims_per_leaf: int
    collapse_traversal_cost: float

    @classmethod
    def preset(cls, name: str) -> "BvhBuildParams":
        c = BuildParamsC()
        if load_library().obvhs_cuda_build_params_preset(name.encode(), C.byref(c)) != 0:
            raise ValueError(f"unknown preset {name}")
        return cls(bool(c.pre_split), PlocSearchDistance(c.ploc_search_distance), int(c.search_depth_threshold),
                   float(c.reinsertion_batch_ratio), float(c.post_collapse_reinsertion_batch_ratio_multiplier),
                   SortPrecision(c.sort_precision), int(c.max_prims_per_leaf), float(c.collapse_traversal_cost))

    # src/lib.rs:233-305
    @classmethod
    def fastest_build(cls): return cls.preset("fastest_build")
    @classmethod
    def very_fast_build(cls): return cls.preset("very_fast_build")
    @classmethod
    def fast_build(cls): return cls.preset("fast_build")
    @classmethod
    def medium_build(cls): return cls.preset("medium_build")
    @classmethod
    def slow_build(cls): return cls.preset("slow_build")
    @classmethod
    def very_slow_build(cls): return cls.preset("very_slow_build")

    def to_c(self) -> BuildParamsC:
        return BuildParamsC(int(self.pre_split), int(self.ploc_search_distance), int(self.search_depth_threshold),
                            float(self.reinsertion_batch_ratio), float(self.post_collapse_reinsertion_batch_ratio_multiplier),
                            int(self.sort_precision), int(self.max_prims_per_leaf), float(self.collapse_traversal_cost))


def _query_batch(ctx, fn, handle, q, extra=()):
    """count pass, then fill: -> (counts[n], ids[total]) with every query's ids in the reference's call order"""
    n = q.shape[0]
    counts = np.zeros(n, dtype=np.uint32)
    total = _sz(0)
    ctx.check(fn(ctx.h, handle, _ptr(q), n, *extra, _ptr(counts), None, 0, C.byref(total)))
    ids = np.zeros(max(1, total.value), dtype=np.uint32)
    if total.value:
        ctx.check(fn(ctx.h, handle, _ptr(q), n, *extra, _ptr(counts), _ptr(ids), total.value, C.byref(total)))
    return counts, ids[: total.value]


def _points4(points):
    if _is_torch(points):
        return _as_f32(points, 4)
    p = np.asarray(points, dtype=np.float32)
    if p.ndim == 2 and p.shape[1] == 4:
        return np.ascontiguousarray(p)
    out = np.zeros((p.shape[0], 4), np.float32)
    out[:, :3] = p[:, :3]
    return out


class Bvh2:
    """Device-resident Bvh2 (src/bvh2/mod.rs:31-85)."""

    def __init__(self, ctx: Context, handle):
        self.ctx, self.h = ctx, handle

    def __del__(self):
        if getattr(self, "h", None):  # the native handle keeps its context alive, so the order of destruction is free
            self.ctx.lib.obvhs_cuda_bvh2_free(self.h)
        self.h = None

    node_count = property(lambda s: int(s.ctx.lib.obvhs_cuda_bvh2_node_count(s.h)))
    prim_count = property(lambda s: int(s.ctx.lib.obvhs_cuda_bvh2_prim_count(s.h)))
    max_depth = property(lambda s: int(s.ctx.lib.obvhs_cuda_bvh2_max_depth(s.h)))
    ploc_iterations = property(lambda s: int(s.ctx.lib.obvhs_cuda_bvh2_ploc_iterations(s.h)))
    children_are_ordered_after_parents = property(lambda s: bool(s.ctx.lib.obvhs_cuda_bvh2_children_ordered_after_parents(s.h)))
    uses_spatial_splits = property(lambda s: bool(s.ctx.lib.obvhs_cuda_bvh2_uses_spatial_splits(s.h)),
                                   lambda s, v: s.ctx.lib.obvhs_cuda_bvh2_set_uses_spatial_splits(s.h, int(v)))

    @classmethod
    def upload(cls, nodes, primitive_indices, max_depth=96, children_ordered_after_parents=False, ctx: Context | None = None):
        ctx = ctx or default_context()
        nodes = np.ascontiguousarray(nodes, dtype=BVH2_NODE)
        prims = np.ascontiguousarray(primitive_indices, dtype=np.uint32)
        h = C.c_void_p()
        ctx.check(ctx.lib.obvhs_cuda_bvh2_upload(ctx.h, _ptr(nodes), nodes.shape[0], _ptr(prims), prims.shape[0], max_depth,
                                                  int(children_ordered_after_parents), C.byref(h)))
        return cls(ctx, h)

    def download(self, with_parents=False):
        """-> (nodes[BVH2_NODE], primitive_indices[u32][, parents[u32]]) as numpy arrays"""
        nodes = np.zeros(self.node_count, dtype=BVH2_NODE)
        prims = np.zeros(self.prim_count, dtype=np.uint32)
        parents = np.zeros(self.node_count, dtype=np.uint32) if with_parents else None
        self.ctx.check(self.ctx.lib.obvhs_cuda_bvh2_download(self.ctx.h, self.h, _ptr(nodes), _ptr(prims), _ptr(parents)))
        return (nodes, prims, parents) if with_parents else (nodes, prims)

    def compute_parents(self):
        self.ctx.check(self.ctx.lib.obvhs_cuda_bvh2_compute_parents(self.ctx.h, self.h))

    def reorder_in_stack_traversal_order(self):
        """Bvh2::reorder_in_stack_traversal_order (src/bvh2/mod.rs:462-500)."""
        self.ctx.check(self.ctx.lib.obvhs_cuda_bvh2_reorder_in_stack_traversal_order(self.ctx.h, self.h))

    def compute_primitives_to_nodes(self):
        """Bvh2::compute_primitives_to_nodes (src/bvh2/mod.rs:647-665) from the downloaded tree (host-side bookkeeping)."""
        from .types import compute_primitives_to_nodes

        nodes, prims = self.download()
        return compute_primitives_to_nodes(nodes, prims)

    def refit_all(self):
        self.ctx.check(self.ctx.lib.obvhs_cuda_bvh2_refit_all(self.ctx.h, self.h))

    def set_node_aabbs(self, node_ids, aabbs):
        """bvh.nodes[id].set_aabb(aabb) for a list of nodes (examples/physics.rs:446)"""
        ids = np.ascontiguousarray(node_ids, dtype=np.uint32)
        a = _as_f32(aabbs, 8)
        assert a.shape[0] == ids.shape[0]
        self.ctx.check(self.ctx.lib.obvhs_cuda_bvh2_set_node_aabbs(self.ctx.h, self.h, _ptr(ids), _ptr(a), ids.shape[0]))

    def set_leaf_aabbs(self, prim_aabbs):
        a = _as_f32(prim_aabbs, 8)
        self.ctx.check(self.ctx.lib.obvhs_cuda_bvh2_set_leaf_aabbs(self.ctx.h, self.h, _ptr(a), a.shape[0]))

    def collapse(self, max_prims: int, traversal_cost: float):
        """src/bvh2/leaf_collapser.rs:21 `collapse(bvh, max_prims, traversal_cost)`"""
        self.ctx.check(self.ctx.lib.obvhs_cuda_bvh2_collapse(self.ctx.h, self.h, int(max_prims), float(traversal_cost)))

    def set_triangles(self, tris):
        """Attach tris[primitive_indices] for ray traversal (examples/demoscene.rs:66-70)."""
        t = _as_f32(tris, 12)
        self.ctx.check(self.ctx.lib.obvhs_cuda_bvh2_set_triangles(self.ctx.h, self.h, _ptr(t), t.shape[0]))

    def aabb_traverse(self, aabbs):
        """Batched Bvh2::aabb_traverse (src/bvh2/mod.rs:365-407) with eval -> true: (counts per query, leaf node ids)."""
        return _query_batch(self.ctx, self.ctx.lib.obvhs_cuda_bvh2_aabb_traverse_batch, self.h, _as_f32(aabbs, 8))

    def point_traverse(self, points):
        """Batched Bvh2::point_traverse (src/bvh2/mod.rs:414-456): (counts per query, leaf node ids)."""
        return _query_batch(self.ctx, self.ctx.lib.obvhs_cuda_bvh2_point_traverse_batch, self.h, _points4(points))

    def ray_traverse(self, rays, out=None, counters=None):
        """Batched Bvh2::ray_traverse (src/bvh2/mod.rs:148-172); returns / fills a RAY_HIT array (or an (n,4) int32 device tensor)."""
        r, is_args = _as_rays(rays)
        n = r.shape[0]
        hits = out if out is not None else np.zeros(n, dtype=RAY_HIT)
        if is_args:
            assert counters is None
            self.ctx.check(self.ctx.lib.obvhs_cuda_bvh2_ray_new_traverse_batch(self.ctx.h, self.h, _ptr(r), n, _ptr(hits)))
        elif counters is None:
            self.ctx.check(self.ctx.lib.obvhs_cuda_bvh2_ray_traverse_batch(self.ctx.h, self.h, _ptr(r), n, _ptr(hits)))
        else:
            self.ctx.check(self.ctx.lib.obvhs_cuda_bvh2_ray_traverse_batch_counted(self.ctx.h, self.h, _ptr(r), n, _ptr(hits), _ptr(counters)))
        return hits

    def ray_od_traverse(self, origin_dir, tmin=0.0, tmax=math.inf, out=None):
        """ray_traverse of Ray::new(o, d, tmin, tmax) for (n,6) float32 [origin | direction] records: 24 bytes per ray cross PCIe,
        one pair of bounds per batch (defaults = Ray::new_inf, src/ray.rs:55-57)."""
        od = _as_f32(origin_dir, 6)
        n = od.shape[0]
        hits = out if out is not None else np.zeros(n, dtype=RAY_HIT)
        self.ctx.check(self.ctx.lib.obvhs_cuda_bvh2_ray_od_traverse_batch(self.ctx.h, self.h, _ptr(od), n, float(tmin), float(tmax), _ptr(hits)))
        return hits

    def ray_traverse_miss(self, rays, out=None):
        """src/bvh2/mod.rs:185-213"""
        r, is_args = _as_rays(rays)
        n = r.shape[0]
        miss = out if out is not None else np.zeros(n, dtype=np.uint8)
        fn = self.ctx.lib.obvhs_cuda_bvh2_ray_new_traverse_miss_batch if is_args else self.ctx.lib.obvhs_cuda_bvh2_ray_traverse_miss_batch
        self.ctx.check(fn(self.ctx.h, self.h, _ptr(r), n, _ptr(miss)))
        return miss

    def ray_traverse_anyhit_count(self, rays, out=None):
        """src/bvh2/mod.rs:225-238 with a counting closure"""
        r = _as_f32(rays, 16)
        n = r.shape[0]
        counts = out if out is not None else np.zeros(n, dtype=np.uint32)
        self.ctx.check(self.ctx.lib.obvhs_cuda_bvh2_ray_traverse_anyhit_count_batch(self.ctx.h, self.h, _ptr(r), n, _ptr(counts)))
        return counts


ERR_CAPACITY = -6


def _split_retry(n, run):
    """Vec growth over the C ABI: run(aabbs, indices, capacity, count) until the grown arrays fit."""
    cap = n + n // 8 + 1024
    while True:
        a = np.zeros((cap, 8), np.float32)
        idx = np.zeros(cap, np.uint32)
        count = _sz(0)
        try:
            run(a, idx, cap, count)
            return a[: count.value].copy(), idx[: count.value].copy()
        except ObvhsError as e:
            if e.code != ERR_CAPACITY:
                raise
            cap = int(count.value)


def split_aabbs_precise(aabbs, indices, triangles, area_thresh_low, area_thresh_high, split_factor_low, split_factor_high,
                        max_iterations, split_tests, ctx: Context | None = None):
    """src/splits.rs:49-125. Returns the grown (aabbs, indices) (the reference mutates its two Vecs)."""
    ctx = ctx or default_context()
    src = _as_f32(aabbs, 8)
    src_idx = np.ascontiguousarray(indices, dtype=np.uint32)
    t = _as_f32(triangles, 12)
    n = src.shape[0]

    def run(a, idx, cap, count):
        a[:n] = src
        idx[:n] = src_idx
        ctx.check(ctx.lib.obvhs_cuda_split_aabbs_precise(ctx.h, _ptr(a), _ptr(idx), n, cap, _ptr(t), t.shape[0], area_thresh_low,
                                                         area_thresh_high, split_factor_low, split_factor_high, max_iterations,
                                                         split_tests, C.byref(count)))

    return _split_retry(n, run)


def split_aabbs_preset(aabbs, indices, triangles, avg_half_area, largest_half_area, ctx: Context | None = None):
    """src/splits.rs:16-34"""
    ctx = ctx or default_context()
    src = _as_f32(aabbs, 8)
    src_idx = np.ascontiguousarray(indices, dtype=np.uint32)
    t = _as_f32(triangles, 12)
    n = src.shape[0]

    def run(a, idx, cap, count):
        a[:n] = src
        idx[:n] = src_idx
        ctx.check(ctx.lib.obvhs_cuda_split_aabbs_preset(ctx.h, _ptr(a), _ptr(idx), n, cap, _ptr(t), t.shape[0], float(avg_half_area),
                                                        float(largest_half_area), C.byref(count)))

    return _split_retry(n, run)


def presplit_tris(triangles, ctx: Context | None = None):
    """The builders' pre-split prologue (src/cwbvh/builder.rs:27-54): -> (aabbs, indices, avg_half_area, largest_half_area)."""
    ctx = ctx or default_context()
    t = _as_f32(triangles, 12)
    avg, largest = _f32(0), _f32(0)

    def run(a, idx, cap, count):
        ctx.check(ctx.lib.obvhs_cuda_presplit_tris(ctx.h, _ptr(t), t.shape[0], _ptr(a), _ptr(idx), cap, C.byref(count), C.byref(avg),
                                                   C.byref(largest)))

    a, idx = _split_retry(t.shape[0], run)
    return a, idx, np.float32(avg.value), np.float32(largest.value)


def compute_rebuild_path_flags(bvh: Bvh2, leaves, flags=None):
    """src/ploc/rebuild.rs:12-43: one byte per node, 1 for the given leaf nodes and all their ancestors."""
    ids = np.ascontiguousarray(leaves, dtype=np.uint32)
    out = flags if flags is not None else np.zeros(bvh.node_count, dtype=np.uint8)
    bvh.ctx.check(bvh.ctx.lib.obvhs_cuda_compute_rebuild_path_flags(bvh.ctx.h, bvh.h, _ptr(ids), ids.shape[0], _ptr(out)))
    return out


class PlocBuilder:
    """src/ploc/mod.rs:35-159. The context keeps the scratch the reference's builder keeps for reuse."""

    def __init__(self, ctx: Context | None = None):
        self.ctx = ctx or default_context()

    @classmethod
    def with_capacity(cls, _prim_count: int, ctx: Context | None = None):
        return cls(ctx)

    def build(self, search_distance, aabbs, indices=None, sort_precision=SortPrecision.U64, search_depth_threshold: int = 0) -> Bvh2:
        a = _as_f32(aabbs, 8)
        n = a.shape[0]
        idx = None
        if indices is not None:
            idx = indices if _is_torch(indices) else np.ascontiguousarray(indices, dtype=np.uint32)
            assert idx.shape[0] == n
        h = C.c_void_p()
        self.ctx.check(self.ctx.lib.obvhs_cuda_ploc_build(self.ctx.h, _ptr(a), _ptr(idx), n, int(search_distance), int(sort_precision),
                                                         int(search_depth_threshold), C.byref(h)))
        return Bvh2(self.ctx, h)

    def build_tris(self, search_distance, tris, sort_precision=SortPrecision.U64, search_depth_threshold: int = 0) -> Bvh2:
        """PlocBuilder::build over &[Triangle] (Boundable, src/triangle.rs:28-30)."""
        t = _as_f32(tris, 12)
        h = C.c_void_p()
        self.ctx.check(self.ctx.lib.obvhs_cuda_ploc_build_tris(self.ctx.h, _ptr(t), t.shape[0], int(search_distance), int(sort_precision),
                                                              int(search_depth_threshold), C.byref(h)))
        return Bvh2(self.ctx, h)

    def full_rebuild(self, bvh: Bvh2, search_distance, sort_precision=SortPrecision.U64, search_depth_threshold: int = 0):
        """src/ploc/rebuild.rs:56-80"""
        self.ctx.check(self.ctx.lib.obvhs_cuda_ploc_full_rebuild(self.ctx.h, bvh.h, int(search_distance), int(sort_precision),
                                                                 int(search_depth_threshold)))

    def partial_rebuild(self, bvh: Bvh2, should_remove, search_distance, sort_precision=SortPrecision.U64, search_depth_threshold: int = 0):
        """src/ploc/rebuild.rs:101-135. should_remove: one flag per node (array / device tensor), or a callable node_id -> bool
        like the reference's closure (evaluated on the host for every node)."""
        if callable(should_remove):
            should_remove = np.fromiter((bool(should_remove(i)) for i in range(bvh.node_count)), dtype=np.uint8, count=bvh.node_count)
        flags = should_remove if _is_torch(should_remove) else np.ascontiguousarray(should_remove, dtype=np.uint8)
        assert flags.shape[0] == bvh.node_count
        self.ctx.check(self.ctx.lib.obvhs_cuda_ploc_partial_rebuild(self.ctx.h, bvh.h, _ptr(flags), int(search_distance), int(sort_precision),
                                                                    int(search_depth_threshold)))

    def morton_sort(self, aabbs, sort_precision=SortPrecision.U64):
        """Stage probe: (codes_lo, codes_hi, sorted order, scene AABB) -- see obvhs_cuda_morton_sort."""
        a = _as_f32(aabbs, 8)
        n = a.shape[0]
        lo, hi = np.zeros(n, np.uint64), np.zeros(n, np.uint64)
        order, total = np.zeros(n, np.uint32), np.zeros(8, np.float32)
        self.ctx.check(self.ctx.lib.obvhs_cuda_morton_sort(self.ctx.h, _ptr(a), n, int(sort_precision), _ptr(lo), _ptr(hi), _ptr(order),
                                                          _ptr(total)))
        return lo, hi, order, total


class ReinsertionOptimizer:
    """src/bvh2/reinsertion.rs:22-57"""

    def __init__(self):
        self.applied = 0

    def run(self, bvh: Bvh2, batch_size_ratio: float, ratio_sequence=None):
        seq = None if ratio_sequence is None else np.ascontiguousarray(ratio_sequence, dtype=np.float32)
        applied = C.c_uint64(0)
        bvh.ctx.check(bvh.ctx.lib.obvhs_cuda_reinsertion_run(bvh.ctx.h, bvh.h, float(batch_size_ratio), _ptr(seq),
                                                            0 if seq is None else seq.shape[0], C.byref(applied)))
        self.applied = int(applied.value)
        return self.applied

    def run_with_candidates(self, bvh: Bvh2, candidates, iterations: int):
        """src/bvh2/reinsertion.rs:66-90"""
        ids = candidates if _is_torch(candidates) else np.ascontiguousarray(candidates, dtype=np.uint32)
        applied = C.c_uint64(0)
        bvh.ctx.check(bvh.ctx.lib.obvhs_cuda_reinsertion_run_with_candidates(bvh.ctx.h, bvh.h, _ptr(ids), int(ids.shape[0]), int(iterations),
                                                                            C.byref(applied)))
        self.applied = int(applied.value)
        return self.applied


class CwBvh:
    """Device-resident CwBvh (src/cwbvh/mod.rs:43-55) with, optionally, the triangles permuted by primitive_indices."""

    def __init__(self, ctx: Context, handle):
        self.ctx, self.h = ctx, handle
        self.core_build_seconds = 0.0

    def __del__(self):
        if getattr(self, "h", None):
            self.ctx.lib.obvhs_cuda_cwbvh_free(self.h)
        self.h = None

    node_count = property(lambda s: int(s.ctx.lib.obvhs_cuda_cwbvh_node_count(s.h)))
    prim_count = property(lambda s: int(s.ctx.lib.obvhs_cuda_cwbvh_prim_count(s.h)))
    uses_spatial_splits = property(lambda s: bool(s.ctx.lib.obvhs_cuda_cwbvh_uses_spatial_splits(s.h)),
                                   lambda s, v: s.ctx.lib.obvhs_cuda_cwbvh_set_uses_spatial_splits(s.h, int(v)))

    @classmethod
    def upload(cls, nodes, primitive_indices, total_aabb=None, ctx: Context | None = None):
        ctx = ctx or default_context()
        nodes = np.ascontiguousarray(nodes, dtype=CWBVH_NODE)
        prims = np.ascontiguousarray(primitive_indices, dtype=np.uint32)
        total = np.ascontiguousarray(np.zeros(8) if total_aabb is None else total_aabb, dtype=np.float32)
        h = C.c_void_p()
        ctx.check(ctx.lib.obvhs_cuda_cwbvh_upload(ctx.h, _ptr(nodes), nodes.shape[0], _ptr(prims), prims.shape[0], _ptr(total), C.byref(h)))
        return cls(ctx, h)

    @classmethod
    def alloc(cls, node_count, prim_count, with_triangles=True, total_aabb=None, ctx: Context | None = None):
        ctx = ctx or default_context()
        total = np.ascontiguousarray(np.zeros(8) if total_aabb is None else total_aabb, dtype=np.float32)
        h = C.c_void_p()
        ctx.check(ctx.lib.obvhs_cuda_cwbvh_alloc(ctx.h, node_count, prim_count, int(with_triangles), _ptr(total), C.byref(h)))
        return cls(ctx, h)

    @staticmethod
    def broadcast(bvh: "CwBvh | None", ctx: Context, root: int = 0) -> "CwBvh":
        """obvhs_cuda_cwbvh_broadcast (collective over ctx's communicator): `bvh` is the finished tree on `root`; elsewhere None or a
        replica from an earlier broadcast (reused when the sizes match). Returns this rank's handle; the transfer is ordered on
        the context's stream."""
        h = C.c_void_p(bvh.h.value if bvh is not None else None)
        ctx.check(ctx.lib.obvhs_cuda_cwbvh_broadcast(ctx.h, C.byref(h), int(root)))
        if bvh is not None:
            if bvh.h.value == h.value:
                return bvh
            bvh.h = None  # the library freed (replaced) the old handle
        return CwBvh(ctx, h)

    def download(self):
        nodes = np.zeros(self.node_count, dtype=CWBVH_NODE)
        prims = np.zeros(self.prim_count, dtype=np.uint32)
        total = np.zeros(8, dtype=np.float32)
        self.ctx.check(self.ctx.lib.obvhs_cuda_cwbvh_download(self.ctx.h, self.h, _ptr(nodes), _ptr(prims), _ptr(total)))
        return nodes, prims, total

    def compute_parents(self, out=None):
        """CwBvh::compute_parents (src/cwbvh/mod.rs:494-509): parent node index per node (parents[0] = 0)."""
        parents = out if out is not None else np.zeros(self.node_count, dtype=np.uint32)
        self.ctx.check(self.ctx.lib.obvhs_cuda_cwbvh_compute_parents(self.ctx.h, self.h, _ptr(parents)))
        return parents

    def order_children(self, prim_aabbs, direct_layout: bool = False):
        """CwBvh::order_children(primitives, direct_layout) (src/cwbvh/mod.rs:520-524); the primitives as (n, 8) AABBs."""
        a = _as_f32(prim_aabbs, 8)
        self.ctx.check(self.ctx.lib.obvhs_cuda_cwbvh_order_children(self.ctx.h, self.h, _ptr(a), a.shape[0], int(direct_layout)))

    def exact_node_aabbs(self):
        """CwBvh::exact_node_aabbs (src/cwbvh/mod.rs:47) as an (n, 8) float32 array, or None when absent."""
        count = _sz(0)
        self.ctx.check(self.ctx.lib.obvhs_cuda_cwbvh_exact_node_aabbs(self.ctx.h, self.h, None, 0, C.byref(count)))
        if count.value == 0:
            return None
        out = np.zeros((count.value, 8), dtype=np.float32)
        self.ctx.check(self.ctx.lib.obvhs_cuda_cwbvh_exact_node_aabbs(self.ctx.h, self.h, _ptr(out), count.value, C.byref(count)))
        return out

    def set_triangles(self, tris):
        t = _as_f32(tris, 12)
        self.ctx.check(self.ctx.lib.obvhs_cuda_cwbvh_set_triangles(self.ctx.h, self.h, _ptr(t), t.shape[0]))

    def device_ptrs(self):
        """(nodes, primitive_indices, bvh_tris) device addresses as ints (0 = absent)."""
        a, b, c = C.c_void_p(), C.c_void_p(), C.c_void_p()
        self.ctx.check(self.ctx.lib.obvhs_cuda_cwbvh_device_ptrs(self.h, C.byref(a), C.byref(b), C.byref(c)))
        return a.value or 0, b.value or 0, c.value or 0

    def total_aabb(self):
        """CwBvh::total_aabb as 8 floats (Aabb layout) without downloading the tree."""
        total = np.zeros(8, dtype=np.float32)
        self.ctx.check(self.ctx.lib.obvhs_cuda_cwbvh_download(self.ctx.h, self.h, None, None, _ptr(total)))
        return total

    def aabb_traverse(self, aabbs, traversal_direction=(0.0, 0.0, 0.0)):
        """traverse!(.., node.intersect_aabb(&aabb, state.oct_inv4), ..) per query box (src/cwbvh/node.rs:157-177):
        (counts per query, primitive slots in call order)."""
        d = np.ascontiguousarray(traversal_direction, dtype=np.float32)
        return _query_batch(self.ctx, self.ctx.lib.obvhs_cuda_cwbvh_aabb_traverse_batch, self.h, _as_f32(aabbs, 8), (_ptr(d),))

    def point_traverse(self, points, traversal_direction=(0.0, 0.0, 0.0)):
        """traverse!(.., node.contains_point(&point, state.oct_inv4), ..) per point (src/cwbvh/node.rs:180-200)."""
        d = np.ascontiguousarray(traversal_direction, dtype=np.float32)
        return _query_batch(self.ctx, self.ctx.lib.obvhs_cuda_cwbvh_point_traverse_batch, self.h, _points4(points), (_ptr(d),))

    def ray_traverse(self, rays, out=None, counters=None):
        """Batched CwBvh::ray_traverse with the triangle closure: -> RayHit per ray (numpy RAY_HIT array, or `out`).

        `out` may be a torch CUDA tensor of shape (n, 4) int32/float32-viewable (16 bytes per ray) to keep hits on device.
        counters: optional np.uint64[2] (or device tensor) accumulating nodes visited / triangles tested."""
        r, is_args = _as_rays(rays)
        n = r.shape[0]
        hits = out if out is not None else np.zeros(n, dtype=RAY_HIT)
        if is_args and counters is not None:  # (n,8) Ray::new arguments: the constructor runs on the device
            self.ctx.check(self.ctx.lib.obvhs_cuda_cwbvh_ray_new_traverse_batch_counted(self.ctx.h, self.h, _ptr(r), n, _ptr(hits), _ptr(counters)))
        elif is_args:
            self.ctx.check(self.ctx.lib.obvhs_cuda_cwbvh_ray_new_traverse_batch(self.ctx.h, self.h, _ptr(r), n, _ptr(hits)))
        elif counters is None:
            self.ctx.check(self.ctx.lib.obvhs_cuda_cwbvh_ray_traverse_batch(self.ctx.h, self.h, _ptr(r), n, _ptr(hits)))
        else:
            self.ctx.check(self.ctx.lib.obvhs_cuda_cwbvh_ray_traverse_batch_counted(self.ctx.h, self.h, _ptr(r), n, _ptr(hits), _ptr(counters)))
        return hits

    def ray_od_traverse(self, origin_dir, tmin=0.0, tmax=math.inf, out=None, hit8=False):
        """ray_traverse of Ray::new(o, d, tmin, tmax) for (n,6) float32 [origin | direction] records (ObvhsRayOd): 24 bytes per ray
        cross PCIe, one pair of bounds per batch (defaults = Ray::new_inf, src/ray.rs:55-57, as examples/demoscene.rs builds its rays).
        hit8: results as 8-byte RAY_HIT8 {primitive_id, t} records (all this path writes into a RayHit) instead of RAY_HIT."""
        od = _as_f32(origin_dir, 6)
        n = od.shape[0]
        hits = out if out is not None else np.zeros(n, dtype=RAY_HIT8 if hit8 else RAY_HIT)
        fn = self.ctx.lib.obvhs_cuda_cwbvh_ray_od_traverse_hit8_batch if hit8 else self.ctx.lib.obvhs_cuda_cwbvh_ray_od_traverse_batch
        self.ctx.check(fn(self.ctx.h, self.h, _ptr(od), n, float(tmin), float(tmax), _ptr(hits)))
        return hits

    def ray_od_traverse_miss(self, origin_dir, tmin=0.0, tmax=math.inf, out=None):
        od = _as_f32(origin_dir, 6)
        n = od.shape[0]
        miss = out if out is not None else np.zeros(n, dtype=np.uint8)
        self.ctx.check(self.ctx.lib.obvhs_cuda_cwbvh_ray_od_traverse_miss_batch(self.ctx.h, self.h, _ptr(od), n, float(tmin), float(tmax), _ptr(miss)))
        return miss

    def ray_traverse_miss(self, rays, out=None):
        r, is_args = _as_rays(rays)
        n = r.shape[0]
        miss = out if out is not None else np.zeros(n, dtype=np.uint8)
        fn = self.ctx.lib.obvhs_cuda_cwbvh_ray_new_traverse_miss_batch if is_args else self.ctx.lib.obvhs_cuda_cwbvh_ray_traverse_miss_batch
        self.ctx.check(fn(self.ctx.h, self.h, _ptr(r), n, _ptr(miss)))
        return miss

    def ray_traverse_anyhit_count(self, rays, out=None):
        r, is_args = _as_rays(rays)
        n = r.shape[0]
        counts = out if out is not None else np.zeros(n, dtype=np.uint32)
        fn = self.ctx.lib.obvhs_cuda_cwbvh_ray_new_traverse_anyhit_count_batch if is_args else self.ctx.lib.obvhs_cuda_cwbvh_ray_traverse_anyhit_count_batch
        self.ctx.check(fn(self.ctx.h, self.h, _ptr(r), n, _ptr(counts)))
        return counts


def bvh2_to_cwbvh(bvh2: Bvh2, max_prims_per_leaf: int = 3, order_children: bool = True, include_exact_node_aabbs: bool = False) -> CwBvh:
    """src/cwbvh/bvh2_to_cwbvh.rs:490-510"""
    h = C.c_void_p()
    bvh2.ctx.check(bvh2.ctx.lib.obvhs_cuda_bvh2_to_cwbvh(bvh2.ctx.h, bvh2.h, int(max_prims_per_leaf), int(order_children),
                                                        int(include_exact_node_aabbs), C.byref(h)))
    return CwBvh(bvh2.ctx, h)


def build_cwbvh_from_tris(triangles, config: BvhBuildParams, core_build_time: list | None = None, ctx: Context | None = None) -> CwBvh:
    """src/cwbvh/builder.rs:20-85. `core_build_time` is a one-element list of seconds that is incremented (the
    reference's `&mut Duration`). The result carries the permuted triangles and is ready to traverse."""
    ctx = ctx or default_context()
    t = _as_f32(triangles, 12)
    secs = C.c_double(0.0)
    params = config.to_c()
    h = C.c_void_p()
    ctx.check(ctx.lib.obvhs_cuda_build_cwbvh_from_tris(ctx.h, _ptr(t), t.shape[0], C.byref(params), C.byref(secs), C.byref(h)))
    bvh = CwBvh(ctx, h)
    bvh.core_build_seconds = secs.value
    if core_build_time is not None:
        core_build_time[0] += secs.value
    return bvh


def build_bvh2_from_tris(triangles, config: BvhBuildParams, core_build_time: list | None = None, ctx: Context | None = None) -> Bvh2:
    """src/bvh2/builder.rs:17-91. The result carries the permuted triangles and is ready to traverse."""
    ctx = ctx or default_context()
    t = _as_f32(triangles, 12)
    secs = C.c_double(0.0)
    params = config.to_c()
    h = C.c_void_p()
    ctx.check(ctx.lib.obvhs_cuda_build_bvh2_from_tris(ctx.h, _ptr(t), t.shape[0], C.byref(params), C.byref(secs), C.byref(h)))
    bvh = Bvh2(ctx, h)
    bvh.core_build_seconds = secs.value
    if core_build_time is not None:
        core_build_time[0] += secs.value
    return bvh


def _build_from_aabbs(fn_name, cls, primitives, config, core_build_time, ctx):
    ctx = ctx or default_context()
    a = _as_f32(primitives, 8)
    secs = C.c_double(0.0)
    params = config.to_c()
    h = C.c_void_p()
    ctx.check(getattr(ctx.lib, fn_name)(ctx.h, _ptr(a), a.shape[0], C.byref(params), C.byref(secs), C.byref(h)))
    bvh = cls(ctx, h)
    bvh.core_build_seconds = secs.value
    if core_build_time is not None:
        core_build_time[0] += secs.value
    return bvh


def build_cwbvh(primitives, config: BvhBuildParams, core_build_time: list | None = None, ctx: Context | None = None) -> CwBvh:
    """src/cwbvh/builder.rs:98-123 `build_cwbvh<T: Boundable>`; primitives are AABBs (n, 8). pre_split is ignored."""
    return _build_from_aabbs("obvhs_cuda_build_cwbvh", CwBvh, primitives, config, core_build_time, ctx)


def build_bvh2(primitives, config: BvhBuildParams, core_build_time: list | None = None, ctx: Context | None = None) -> Bvh2:
    """src/bvh2/builder.rs:103-140 `build_bvh2<T: Boundable>`; primitives are AABBs (n, 8). pre_split is ignored."""
    return _build_from_aabbs("obvhs_cuda_build_bvh2", Bvh2, primitives, config, core_build_time, ctx)


def make_rays(origin_dir, tmin=0.0, tmax=3.4028234663852886e38, out=None, ctx: Context | None = None):
    """Ray::new for n rays (src/ray.rs:34-52). origin_dir: (n,6) float32."""
    ctx = ctx or default_context()
    od = _as_f32(origin_dir, 6)
    n = od.shape[0]
    rays = out if out is not None else np.zeros((n, 16), dtype=np.float32)
    ctx.check(ctx.lib.obvhs_cuda_make_rays(ctx.h, _ptr(od), n, float(tmin), float(tmax), _ptr(rays)))
    return rays


def ray_new(args, out=None, ctx: Context | None = None):
    """Ray::new (src/ray.rs:34-52) for n argument records ((n,8) float32, types.make_ray_args) -> (n,16) Ray array."""
    ctx = ctx or default_context()
    a = _as_f32(args, 8)
    n = a.shape[0]
    rays = out if out is not None else np.zeros((n, 16), dtype=np.float32)
    ctx.check(ctx.lib.obvhs_cuda_ray_new_batch(ctx.h, _ptr(a), n, _ptr(rays)))
    return rays
