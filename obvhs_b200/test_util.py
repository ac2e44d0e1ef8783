"""Scene generators, hashing and sampling helpers -- host-side mirror of the reference's `test_util` module.

Mirrors reference src/test_util.rs (sampling: lines 9-46, 63-95; geometry: 184-304) plus the scene builders of
examples/cornell_box_cwbvh.rs:22-61 and the OBJ flattening of examples/helpers/load_obj.rs:7-45. These are INPUT
generators for tests and benchmarks (vectorised numpy, float32 arithmetic with the reference's operation order).
Only `uhash` / `hash_vec3a_vec` are result-bearing: they define the kitchen golden hash
(examples/obj_cwbvh.rs:142-159).
"""
from __future__ import annotations

import ctypes
import ctypes.util
import math
import os

import numpy as np

from .types import triangles_from_vertices

_U32 = np.uint32
TAU = np.float32(6.2831855)

_ASSET_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "assets")


# ----------------------------------------------------------------------------------------------------------
# sampling (src/test_util.rs:9-46)
# ----------------------------------------------------------------------------------------------------------
def uhash(x):
    """src/test_util.rs:9-17 (lowbias32)."""
    x = np.asarray(x, dtype=np.uint32).copy()
    with np.errstate(over="ignore"):
        x ^= x >> _U32(16)
        x *= _U32(0x7FEB352D)
        x ^= x >> _U32(15)
        x *= _U32(0x846CA68B)
        x ^= x >> _U32(16)
    return x


def hash_vec3a_vec(v) -> int:
    """src/test_util.rs:25-30: XOR of uhash(bits) over x,y,z of every vector."""
    bits = np.ascontiguousarray(np.asarray(v, dtype=np.float32)[:, 0:3]).view(np.uint32)
    return int(np.bitwise_xor.reduce(uhash(bits).ravel())) if bits.size else 0


def uhash2(a, b):
    """src/test_util.rs:33-35."""
    a = np.asarray(a, dtype=np.uint32)
    b = np.asarray(b, dtype=np.uint32)
    with np.errstate(over="ignore"):
        return uhash((a * _U32(1597334673)) ^ (b * _U32(3812015801)))


def unormf(n):
    """src/test_util.rs:38-40: n as f32 * (1.0 / 0xffffffff as f32)."""
    return np.asarray(n, dtype=np.uint32).astype(np.float32) * (np.float32(1.0) / np.float32(0xFFFFFFFF))


def hash_noise(x, y, frame):
    """src/test_util.rs:43-46 for coord=(x,y)."""
    x = np.asarray(x, dtype=np.uint32)
    y = np.asarray(y, dtype=np.uint32)
    with np.errstate(over="ignore"):
        return unormf(uhash2(x, (y << _U32(11)) + np.asarray(frame, dtype=np.uint32)))


def build_orthonormal_basis(n):
    """src/test_util.rs:50-61; n is (k,3) float32, returns the three column vectors (each (k,3))."""
    n = np.asarray(n, dtype=np.float32)
    one = np.float32(1.0)
    sign = np.where(np.signbit(n[:, 2]), -one, one).astype(np.float32)
    a = -one / (sign + n[:, 2])
    b = n[:, 0] * n[:, 1] * a
    c0 = np.stack([one + sign * n[:, 0] * n[:, 0] * a, sign * b, -sign * n[:, 0]], axis=1)
    c1 = np.stack([b, sign + n[:, 1] * n[:, 1] * a, -n[:, 1]], axis=1)
    return c0.astype(np.float32), c1.astype(np.float32), n


def cosine_sample_hemisphere(ux, uy):
    """src/test_util.rs:63-72."""
    ux = np.asarray(ux, dtype=np.float32)
    uy = np.asarray(uy, dtype=np.float32)
    r = np.sqrt(ux)
    theta = uy * TAU
    z = np.sqrt(np.maximum(np.float32(0.0), np.float32(1.0) - ux))
    return np.stack([r * np.cos(theta), r * np.sin(theta), z], axis=1).astype(np.float32)


def uniform_sample_sphere(ux, uy):
    """src/test_util.rs:75-80."""
    ux = np.asarray(ux, dtype=np.float32)
    uy = np.asarray(uy, dtype=np.float32)
    z = np.float32(1.0) - np.float32(2.0) * ux
    r = np.sqrt(np.float32(1.0) - z * z)
    theta = uy * TAU
    return np.stack([r * np.cos(theta), r * np.sin(theta), z], axis=1).astype(np.float32)


def _cubic(v0, v1, v2, v3, x):
    """src/test_util.rs:97-104."""
    p = (v3 - v2) - (v0 - v1)
    q = (v0 - v1) - p
    r = v2 - v0
    return p * (x * x * x) + q * (x * x) + r * x + v1


def bicubic_noise(cx, cy, seed):
    """src/test_util.rs:107-131."""
    cx = np.asarray(cx, dtype=np.float32)
    cy = np.asarray(cy, dtype=np.float32)
    ix = np.floor(cx).astype(np.uint32)
    iy = np.floor(cy).astype(np.uint32)
    fx = cx - ix.astype(np.float32)
    fy = cy - iy.astype(np.float32)
    cols = []
    for j in range(4):
        cols.append(
            _cubic(
                hash_noise(ix, iy + _U32(j), seed),
                hash_noise(ix + _U32(1), iy + _U32(j), seed),
                hash_noise(ix + _U32(2), iy + _U32(j), seed),
                hash_noise(ix + _U32(3), iy + _U32(j), seed),
                fx,
            )
        )
    return _cubic(cols[0], cols[1], cols[2], cols[3], fy).astype(np.float32)


# ----------------------------------------------------------------------------------------------------------
# geometry (src/test_util.rs:184-304)
# ----------------------------------------------------------------------------------------------------------
def _tris(rows) -> np.ndarray:
    a = np.asarray(rows, dtype=np.float32).reshape(-1, 3, 3)
    return triangles_from_vertices(a[:, 0], a[:, 1], a[:, 2])


def cube() -> np.ndarray:
    """src/test_util.rs:199-212 CUBE."""
    return _tris(
        [
            [(-1, 1, -1), (1, 1, 1), (1, 1, -1)],
            [(1, 1, 1), (-1, -1, 1), (1, -1, 1)],
            [(-1, 1, 1), (-1, -1, -1), (-1, -1, 1)],
            [(1, -1, -1), (-1, -1, 1), (-1, -1, -1)],
            [(1, 1, -1), (1, -1, 1), (1, -1, -1)],
            [(-1, 1, -1), (1, -1, -1), (-1, -1, -1)],
            [(-1, 1, -1), (-1, 1, 1), (1, 1, 1)],
            [(1, 1, 1), (-1, 1, 1), (-1, -1, 1)],
            [(-1, 1, 1), (-1, 1, -1), (-1, -1, -1)],
            [(1, -1, -1), (1, -1, 1), (-1, -1, 1)],
            [(1, 1, -1), (1, 1, 1), (1, -1, 1)],
            [(-1, 1, -1), (1, 1, -1), (1, -1, -1)],
        ]
    )


def plane() -> np.ndarray:
    """src/test_util.rs:215-218 PLANE."""
    return _tris([[(1, 0, 1), (-1, 0, -1), (-1, 0, 1)], [(1, 0, 1), (1, 0, -1), (-1, 0, -1)]])


def _normalize(v):
    v = np.asarray(v, dtype=np.float32)
    d = (v[..., 0] * v[..., 0] + v[..., 1] * v[..., 1]) + v[..., 2] * v[..., 2]
    return (v / np.sqrt(d)[..., None]).astype(np.float32)


def icosphere(subdivisions: int) -> np.ndarray:
    """src/test_util.rs:221-267."""
    phi = (np.float32(1.0) + np.sqrt(np.float32(5.0))) / np.float32(2.0)
    a, b, c, d, e = np.float32(1.0), np.float32(-1.0), np.float32(0.0), phi, -phi
    p = np.array(
        [(b, d, c), (a, d, c), (b, e, c), (a, e, c), (c, b, d), (c, a, d), (c, b, e), (c, a, e), (d, c, b), (d, c, a), (e, c, b), (e, c, a)],
        dtype=np.float32,
    )
    p = _normalize(p)
    idx = [
        (0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11), (1, 5, 9), (5, 11, 4), (11, 10, 2), (10, 7, 6), (7, 1, 8),
        (3, 9, 4), (3, 4, 2), (3, 2, 6), (3, 6, 8), (3, 8, 9), (4, 9, 5), (2, 4, 11), (6, 2, 10), (8, 6, 7), (9, 8, 1),
    ]  # fmt: skip
    t = np.array([[p[i], p[j], p[k]] for i, j, k in idx], dtype=np.float32)  # (n,3,3)
    half = np.float32(0.5)
    for _ in range(subdivisions):
        v0, v1, v2 = t[:, 0], t[:, 1], t[:, 2]
        m01 = _normalize((v0 + v1) * half)
        m12 = _normalize((v1 + v2) * half)
        m20 = _normalize((v2 + v0) * half)
        new = np.stack(
            [
                np.stack([v0, m01, m20], axis=1),
                np.stack([v1, m12, m01], axis=1),
                np.stack([v2, m20, m12], axis=1),
                np.stack([m01, m12, m20], axis=1),
            ],
            axis=1,
        )  # (n,4,3,3)
        t = new.reshape(-1, 3, 3)
    return triangles_from_vertices(t[:, 0], t[:, 1], t[:, 2])


def height_to_triangles(height, x_resolution: int, z_resolution: int) -> np.ndarray:
    """src/test_util.rs:270-304. `height` is a ((z_resolution+1), (x_resolution+1)) float32 grid h[z, x]."""
    height = np.asarray(height, dtype=np.float32)
    two, one = np.float32(2.0), np.float32(1.0)
    xs = np.arange(x_resolution + 1, dtype=np.float32) / np.float32(x_resolution) * two - one
    zs = np.arange(z_resolution + 1, dtype=np.float32) / np.float32(z_resolution) * two - one
    X, Z = np.meshgrid(np.arange(x_resolution), np.arange(z_resolution))  # row-major: z outer, x inner
    X = X.ravel()
    Z = Z.ravel()

    def vert(xi, zi):
        return np.stack([xs[xi], height[zi, xi], zs[zi]], axis=1)

    v00 = vert(X, Z)
    v10 = vert(X + 1, Z)
    v01 = vert(X, Z + 1)
    v11 = vert(X + 1, Z + 1)
    n = X.shape[0]
    out = np.zeros((2 * n, 12), dtype=np.float32)
    out[0::2] = triangles_from_vertices(v00, v01, v10)
    out[1::2] = triangles_from_vertices(v10, v01, v11)
    return out


def flat_plane(res: int = 4) -> np.ndarray:
    """tests/mod.rs:92,106: height_to_triangles(|_,_| 0.0, 4, 4)."""
    return height_to_triangles(np.zeros((res + 1, res + 1), dtype=np.float32), res, res)


def demoscene(terrain_res: int, seed: int) -> np.ndarray:
    """src/test_util.rs:287-304: 16-octave bicubic hash-noise terrain, 2*terrain_res^2 triangles."""
    r = terrain_res
    gx, gy = np.meshgrid(np.arange(r + 1, dtype=np.float32), np.arange(r + 1, dtype=np.float32))
    cx = gx / np.float32(r)
    cy = gy / np.float32(r)
    cs, ns = np.float32(1.579), np.float32(0.579)
    acc = np.zeros_like(cx)
    for i in range(1, 17):
        cs = np.float32(cs * np.float32(1.579))
        ns = np.float32(ns * np.float32(-0.579))
        acc = acc + bicubic_noise(cx * cs, cy * cs, np.uint32(seed + i)) * ns
    one_m = np.float32(1.0) - cy
    h = acc * np.power(one_m, np.float32(0.579)) + np.power(one_m, np.float32(1.579)) * np.float32(0.579)
    return height_to_triangles(h.astype(np.float32), r, r)  # h[y(z), x]


def triangle_soup(n: int, seed: int = 0) -> np.ndarray:
    """n hashed small triangles in the unit cube (incoherent stress scene)."""
    i = np.arange(n, dtype=np.uint32) + np.uint32((seed * 0x9E3779B9) & 0xFFFFFFFF)
    c = np.stack([unormf(uhash2(i, k)) for k in (1, 2, 3)], axis=1)
    s = np.float32(0.5 * float(n) ** (-1.0 / 3.0))
    vs = []
    for k in range(3):
        off = np.stack([unormf(uhash2(i, 4 + 3 * k + a)) for a in range(3)], axis=1)
        vs.append(c + s * (np.float32(2.0) * off - np.float32(1.0)))
    return triangles_from_vertices(vs[0], vs[1], vs[2])


def soup_with_large_triangles(n: int, n_large: int, seed: int = 0) -> np.ndarray:
    """A triangle soup plus a few long, thin, diagonal triangles spanning the unit cube: the case spatial pre-splits
    (src/splits.rs) exist for; the large ones are split through several iterations."""
    small = triangle_soup(n, seed) if n else np.zeros((0, 12), np.float32)
    i = np.arange(n_large, dtype=np.uint32) + np.uint32(977) + np.uint32(seed)
    vs = [np.stack([unormf(uhash2(i, 31 + 3 * k + a)) for a in range(3)], axis=1) for k in range(3)]
    # slivers along a random diagonal: the third vertex sits near the middle of the long edge
    v2 = (vs[0] + vs[1]) * np.float32(0.5) + (vs[2] - np.float32(0.5)) * np.float32(0.04)
    large = triangles_from_vertices(vs[0], vs[1], v2.astype(np.float32))
    out = np.concatenate([small, large], axis=0)
    # interleave so the large ones are not all at the end of the index range
    perm = np.argsort(uhash2(np.arange(out.shape[0], dtype=np.uint32), np.uint32(99) + np.uint32(seed)), kind="stable")
    return np.ascontiguousarray(out[perm])


# ----------------------------------------------------------------------------------------------------------
# glam-style transforms used by the Cornell box (examples/cornell_box_cwbvh.rs:22-61)
# ----------------------------------------------------------------------------------------------------------
def _quat_axis(axis: int, angle: float):
    s, c = np.float32(math.sin(np.float32(angle) * np.float32(0.5))), np.float32(math.cos(np.float32(angle) * np.float32(0.5)))
    q = [np.float32(0.0)] * 4
    q[axis] = s
    q[3] = c
    return q


def _mat_srt(scale, quat, translation):
    x, y, z, w = quat
    x2, y2, z2 = x + x, y + y, z + z
    xx, xy, xz = x * x2, x * y2, x * z2
    yy, yz, zz = y * y2, y * z2, z * z2
    wx, wy, wz = w * x2, w * y2, w * z2
    one = np.float32(1.0)
    ax = np.array([one - (yy + zz), xy + wz, xz - wy], dtype=np.float32) * np.float32(scale[0])
    ay = np.array([xy - wz, one - (xx + zz), yz + wx], dtype=np.float32) * np.float32(scale[1])
    az = np.array([xz + wy, yz - wx, one - (xx + yy)], dtype=np.float32) * np.float32(scale[2])
    return ax, ay, az, np.asarray(translation, dtype=np.float32)


def _transform(tris: np.ndarray, m) -> np.ndarray:
    ax, ay, az, t = m
    out = tris.copy()
    for k in (0, 4, 8):
        p = tris[:, k : k + 3]
        out[:, k : k + 3] = ((ax[None, :] * p[:, 0:1] + ay[None, :] * p[:, 1:2]) + az[None, :] * p[:, 2:3]) + t[None, :]
    return out


def cornell_box() -> np.ndarray:
    """examples/cornell_box_cwbvh.rs:22-61 -> 34 triangles."""
    ident = [np.float32(0.0)] * 3 + [np.float32(1.0)]
    floor = plane()
    rad = lambda d: np.float32(d) * (np.float32(math.pi) / np.float32(180.0))  # noqa: E731
    box1 = _transform(cube(), _mat_srt((0.3, 0.3, 0.3), _quat_axis(1, rad(-17.5)), (0.33, 0.3, 0.37)))
    box2 = _transform(cube(), _mat_srt((0.3, 0.6, 0.3), _quat_axis(1, rad(17.5)), (-0.33, 0.6, -0.29)))
    ceiling = _transform(floor, _mat_srt((1, 1, 1), ident, (0.0, 2.0, 0.0)))
    half_pi = np.float32(math.pi) * np.float32(0.5)
    wall1 = _transform(floor, _mat_srt((1, 1, 1), _quat_axis(0, half_pi), (0.0, 1.0, -1.0)))
    wall2 = _transform(floor, _mat_srt((1, 1, 1), _quat_axis(2, -half_pi), (-1.0, 1.0, 0.0)))
    wall3 = _transform(floor, _mat_srt((1, 1, 1), _quat_axis(2, -half_pi), (1.0, 1.0, 0.0)))
    return np.concatenate([floor, box1, box2, ceiling, wall1, wall2, wall3], axis=0)


# ----------------------------------------------------------------------------------------------------------
# OBJ loading (examples/helpers/load_obj.rs:7-45)
# ----------------------------------------------------------------------------------------------------------
def zstd_decompress(blob: bytes, max_size: int = 1 << 28) -> bytes:
    """Decompress a zstd frame with the system libzstd (no python zstd module in the image)."""
    name = ctypes.util.find_library("zstd") or "libzstd.so.1"
    lib = ctypes.CDLL(name)
    lib.ZSTD_decompress.restype = ctypes.c_size_t
    lib.ZSTD_decompress.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t]
    lib.ZSTD_isError.restype = ctypes.c_uint
    lib.ZSTD_isError.argtypes = [ctypes.c_size_t]
    lib.ZSTD_getFrameContentSize.restype = ctypes.c_ulonglong
    lib.ZSTD_getFrameContentSize.argtypes = [ctypes.c_void_p, ctypes.c_size_t]
    size = lib.ZSTD_getFrameContentSize(blob, len(blob))
    cap = int(size) if 0 < size < max_size else 64 << 20
    buf = ctypes.create_string_buffer(cap)
    got = lib.ZSTD_decompress(buf, cap, blob, len(blob))
    if lib.ZSTD_isError(got):
        raise RuntimeError("zstd decompression failed")
    return buf.raw[:got]


def parse_obj_triangles(text: str) -> np.ndarray:
    """Positions only; polygons are emitted in file order, quads as (a,b,c),(a,c,d) like load_obj.rs:24-40."""
    verts = []
    faces = []
    for line in text.splitlines():
        if line.startswith("v "):
            _, x, y, z = line.split()[:4]
            verts.append((float(x), float(y), float(z)))
        elif line.startswith("f "):
            idx = []
            for tok in line.split()[1:]:
                i = int(tok.split("/")[0])
                idx.append(i - 1 if i > 0 else len(verts) + i)
            faces.append((idx[0], idx[1], idx[2]))
            if len(idx) == 4:
                faces.append((idx[0], idx[2], idx[3]))
    v = np.asarray(verts, dtype=np.float32)
    f = np.asarray(faces, dtype=np.int64)
    return triangles_from_vertices(v[f[:, 0]], v[f[:, 1]], v[f[:, 2]])


def load_obj(path: str) -> np.ndarray:
    with open(path, "rb") as fh:
        blob = fh.read()
    if path.endswith(".zst"):
        blob = zstd_decompress(blob)
    return parse_obj_triangles(blob.decode("utf-8", errors="replace"))


def kitchen() -> np.ndarray:
    """The 56,939-triangle kitchen scene (BASELINE config 1) from the repo's own binary fixture.

    obvhs_b200/assets/kitchen_tris.npz holds the vertex positions and triangle indices extracted from the
    reference's assets/kitchen.obj.zst by tests/golden/make_kitchen_fixture.py (file order preserved).
    """
    with np.load(os.path.join(_ASSET_DIR, "kitchen_tris.npz")) as z:
        v = z["positions"].astype(np.float32)
        f = z["faces"].astype(np.int64)
    return triangles_from_vertices(v[f[:, 0]], v[f[:, 1]], v[f[:, 2]])
