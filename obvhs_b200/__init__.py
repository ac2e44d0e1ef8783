"""B200-native hot path of obvhs (PLOC BVH2 build, reinsertion, CWBVH collapse, ray traversal) behind a C ABI.

The reference's names, re-exported from the ctypes host side (`obvhs_b200.api`); nothing here computes on the CPU: creating a
`Context` (explicitly or through the first call) raises when the CUDA library or device is missing.
"""
from .api import (  # noqa: F401
    Bvh2,
    BvhBuildParams,
    Context,
    CwBvh,
    ObvhsError,
    PlocBuilder,
    PlocSearchDistance,
    ReinsertionOptimizer,
    SortPrecision,
    build_bvh2,
    build_bvh2_from_tris,
    build_cwbvh,
    build_cwbvh_from_tris,
    bvh2_to_cwbvh,
    compute_rebuild_path_flags,
    nccl_unique_id,
    presplit_tris,
    ray_new,
    split_aabbs_precise,
    split_aabbs_preset,
)
from .types import RAY_HIT, RAY_HIT8, make_ray_args, make_rays, ray_args_of, safe_inverse  # noqa: F401
