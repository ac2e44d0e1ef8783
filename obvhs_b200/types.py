"""POD layouts shared by the host API, the C ABI (include/obvhs_cuda.h) and the tests.

Every layout is byte-identical to the reference's `#[repr(C)]` Pod type it mirrors:

* ``Aabb``      32 B  two 16-byte Vec3A lanes            (reference src/aabb.rs:11-16)
* ``Triangle``  48 B  three Vec3A                         (src/triangle.rs:8-13)
* ``Bvh2Node``  48 B  aabb, prim_count, first_index, meta (src/bvh2/node.rs:40-66)
* ``CwBvhNode`` 80 B  compressed wide node                (src/cwbvh/node.rs:12-54)
* ``Ray``       64 B  origin, direction, inv_direction, tmin, tmax (src/ray.rs:15-30)
* ``RayHit``    16 B  primitive/geometry/instance id, t  (src/ray.rs:63-70)

AABBs, triangles and rays are handled as plain float32 matrices (n x 8, n x 12, n x 16) because that is what
the kernels read; the padding lane of each Vec3A is unspecified in the reference and is never compared.
"""
from __future__ import annotations

import numpy as np

AABB_F32 = 8
TRI_F32 = 12
RAY_F32 = 16

BVH2_NODE = np.dtype(
    [("aabb", "<f4", (8,)), ("prim_count", "<u4"), ("first_index", "<u4"), ("meta1", "<u4"), ("meta2", "<u4")]
)
CWBVH_NODE = np.dtype(
    [
        ("p", "<f4", (3,)),
        ("e", "u1", (3,)),
        ("imask", "u1"),
        ("child_base_idx", "<u4"),
        ("primitive_base_idx", "<u4"),
        ("child_meta", "u1", (8,)),
        ("child_min_x", "u1", (8,)),
        ("child_max_x", "u1", (8,)),
        ("child_min_y", "u1", (8,)),
        ("child_max_y", "u1", (8,)),
        ("child_min_z", "u1", (8,)),
        ("child_max_z", "u1", (8,)),
    ]
)
RAY_HIT = np.dtype([("primitive_id", "<u4"), ("geometry_id", "<u4"), ("instance_id", "<u4"), ("t", "<f4")])

assert BVH2_NODE.itemsize == 48
assert CWBVH_NODE.itemsize == 80
RAY_HIT8 = np.dtype([("primitive_id", "<u4"), ("t", "<f4")])  # ObvhsRayHit8
assert RAY_HIT.itemsize == 16 and RAY_HIT8.itemsize == 8

INVALID_ID = 0xFFFFFFFF
F32_EPSILON = np.float32(1.1920929e-07)


def triangles_from_vertices(v0, v1, v2) -> np.ndarray:
    """Pack three (n,3) float32 vertex arrays into the 48-byte Triangle layout (n,12)."""
    v0 = np.asarray(v0, dtype=np.float32)
    n = v0.shape[0]
    out = np.zeros((n, TRI_F32), dtype=np.float32)
    out[:, 0:3] = v0
    out[:, 4:7] = np.asarray(v1, dtype=np.float32)
    out[:, 8:11] = np.asarray(v2, dtype=np.float32)
    return out


def aabbs_from_min_max(mn, mx) -> np.ndarray:
    mn = np.asarray(mn, dtype=np.float32)
    out = np.zeros((mn.shape[0], AABB_F32), dtype=np.float32)
    out[:, 0:3] = mn
    out[:, 4:7] = np.asarray(mx, dtype=np.float32)
    return out


def safe_inverse(x: np.ndarray) -> np.ndarray:
    """reference src/ray.rs:6-12: |x| <= f32::EPSILON -> signum(x)/EPSILON (signum(+-0) = +-1), else 1/x."""
    x = np.asarray(x, dtype=np.float32)
    small = np.abs(x) <= F32_EPSILON
    sign = np.where(np.signbit(x), np.float32(-1.0), np.float32(1.0))
    with np.errstate(divide="ignore"):
        inv = np.float32(1.0) / x
    return np.where(small, sign / F32_EPSILON, inv).astype(np.float32)


def make_rays(origin, direction, tmin=0.0, tmax=np.float32(3.4028235e38)) -> np.ndarray:
    """`Ray::new` (reference src/ray.rs:34-52) for n rays -> (n,16) float32 in the 64-byte Ray layout."""
    direction = np.asarray(direction, dtype=np.float32)
    n = direction.shape[0]
    rays = np.zeros((n, RAY_F32), dtype=np.float32)
    rays[:, 0:3] = np.asarray(origin, dtype=np.float32)
    rays[:, 4:7] = direction
    rays[:, 8:11] = safe_inverse(direction)
    rays[:, 12] = np.asarray(tmin, dtype=np.float32)
    rays[:, 13] = np.asarray(tmax, dtype=np.float32)
    return rays


RAY_NEW_F32 = 8


def make_ray_args(origin, direction, tmin=0.0, tmax=np.float32(3.4028235e38)) -> np.ndarray:
    """The arguments of `Ray::new(origin, direction, min, max)` (reference src/ray.rs:34-52) for n rays -> (n,8) float32
    [ox oy oz tmin | dx dy dz tmax] (ObvhsRayNew, 32 bytes). The traversal calls accept it in place of the (n,16) Ray array
    and run the constructor (safe_inverse) on the device; `ray_new` returns exactly what `make_rays` builds on the host."""
    direction = np.asarray(direction, dtype=np.float32)
    n = direction.shape[0]
    a = np.zeros((n, RAY_NEW_F32), dtype=np.float32)
    a[:, 0:3] = np.asarray(origin, dtype=np.float32)
    a[:, 3] = np.asarray(tmin, dtype=np.float32)
    a[:, 4:7] = direction
    a[:, 7] = np.asarray(tmax, dtype=np.float32)
    return a


def ray_args_of(rays: np.ndarray) -> np.ndarray:
    """(n,16) Ray structs -> their (n,8) constructor arguments (drops inv_direction)."""
    rays = np.asarray(rays, dtype=np.float32)
    a = np.empty((rays.shape[0], RAY_NEW_F32), dtype=np.float32)
    a[:, 0:3] = rays[:, 0:3]
    a[:, 3] = rays[:, 12]
    a[:, 4:7] = rays[:, 4:7]
    a[:, 7] = rays[:, 13]
    return a


def compute_primitives_to_nodes(nodes: np.ndarray, primitive_indices: np.ndarray) -> np.ndarray:
    """`Bvh2::compute_primitives_to_nodes` (reference src/bvh2/mod.rs:647-665) over downloaded arrays: for every primitive id the
    leaf NODE that holds it (0xFFFFFFFF when none; with spatial splits the last leaf in node order wins, as in the reference's
    loop). What examples/physics.rs indexes before `resize_node` / `reinsert_node`; host-side bookkeeping, not a kernel."""
    prims = np.asarray(primitive_indices, dtype=np.uint32)
    out = np.full(prims.shape[0], 0xFFFFFFFF, dtype=np.uint32)
    leaf = np.nonzero(nodes["prim_count"] != 0)[0]
    counts = nodes["prim_count"][leaf].astype(np.int64)
    if leaf.size == 0 or counts.sum() == 0:
        return out
    first = np.repeat(nodes["first_index"][leaf].astype(np.int64), counts)
    within = np.arange(counts.sum(), dtype=np.int64) - np.repeat(np.cumsum(counts) - counts, counts)
    out[prims[first + within]] = np.repeat(leaf.astype(np.uint32), counts)  # repeated ids: the last assignment wins
    return out
