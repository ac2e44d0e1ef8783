"""Cameras and ray recipes of the reference's example drivers, vectorised (float32, glam operation order).

* `primary_rays`       : the per-pixel recipe shared by examples/obj_cwbvh.rs:94-104,
                          examples/cornell_box_cwbvh.rs:104-115 and tests/mod.rs:154-176.
* `demoscene_primary`  : AA-jittered, depth-of-field primary rays of examples/demoscene.rs:126-152.
* `diffuse_bounce_rays`: cosine-hemisphere bounce rays of examples/demoscene.rs:163-178 (the incoherent set).

glam itself is not vendored in the reference; Mat4 routines follow its published scalar implementation
(perspective_infinite_reverse_rh, look_at_rh, the cofactor inverse, column-major mat*vec). These only generate
INPUT rays: ulp differences against the SSE2 backend cannot change a parity result, because the oracle and the
GPU consume the very same ray arrays.
"""
from __future__ import annotations

import math

import numpy as np

from . import test_util as tu
from .types import make_rays

f32 = np.float32


def _norm3(v):
    v = np.asarray(v, dtype=np.float32)
    d = (v[..., 0] * v[..., 0] + v[..., 1] * v[..., 1]) + v[..., 2] * v[..., 2]
    return (v / np.sqrt(d)[..., None]).astype(np.float32)


def _cross(a, b):
    a = np.asarray(a, dtype=np.float32)
    b = np.asarray(b, dtype=np.float32)
    return np.array(
        [a[1] * b[2] - b[1] * a[2], a[2] * b[0] - b[2] * a[0], a[0] * b[1] - b[0] * a[1]], dtype=np.float32
    )


def _dot(a, b):
    return f32(f32(a[0] * b[0] + a[1] * b[1]) + a[2] * b[2])


def perspective_infinite_reverse_rh(fov_y: float, aspect: float, z_near: float) -> np.ndarray:
    """Columns of glam Mat4::perspective_infinite_reverse_rh."""
    f = f32(1.0) / f32(math.tan(f32(0.5) * f32(fov_y)))
    m = np.zeros((4, 4), dtype=np.float32)  # m[col][row]
    m[0, 0] = f / f32(aspect)
    m[1, 1] = f
    m[2, 3] = f32(-1.0)
    m[3, 2] = f32(z_near)
    return m


def look_at_rh(eye, center, up) -> np.ndarray:
    eye = np.asarray(eye, dtype=np.float32)
    fwd = _norm3(np.asarray(center, dtype=np.float32) - eye)
    s = _norm3(_cross(fwd, np.asarray(up, dtype=np.float32)))
    u = _cross(s, fwd)
    m = np.zeros((4, 4), dtype=np.float32)
    m[0] = [s[0], u[0], -fwd[0], 0.0]
    m[1] = [s[1], u[1], -fwd[1], 0.0]
    m[2] = [s[2], u[2], -fwd[2], 0.0]
    m[3] = [-_dot(eye, s), -_dot(eye, u), _dot(eye, fwd), 1.0]
    return m


def mat4_inverse(m: np.ndarray) -> np.ndarray:
    """glam Mat4::inverse (cofactor expansion), float32. m[col][row]."""
    m = np.asarray(m, dtype=np.float32)
    (m00, m01, m02, m03), (m10, m11, m12, m13), (m20, m21, m22, m23), (m30, m31, m32, m33) = (
        [f32(x) for x in m[c]] for c in range(4)
    )
    coef00 = m22 * m33 - m32 * m23
    coef02 = m12 * m33 - m32 * m13
    coef03 = m12 * m23 - m22 * m13
    coef04 = m21 * m33 - m31 * m23
    coef06 = m11 * m33 - m31 * m13
    coef07 = m11 * m23 - m21 * m13
    coef08 = m21 * m32 - m31 * m22
    coef10 = m11 * m32 - m31 * m12
    coef11 = m11 * m22 - m21 * m12
    coef12 = m20 * m33 - m30 * m23
    coef14 = m10 * m33 - m30 * m13
    coef15 = m10 * m23 - m20 * m13
    coef16 = m20 * m32 - m30 * m22
    coef18 = m10 * m32 - m30 * m12
    coef19 = m10 * m22 - m20 * m12
    coef20 = m20 * m31 - m30 * m21
    coef22 = m10 * m31 - m30 * m11
    coef23 = m10 * m21 - m20 * m11
    v = lambda *a: np.array(a, dtype=np.float32)  # noqa: E731
    fac0, fac1, fac2 = v(coef00, coef00, coef02, coef03), v(coef04, coef04, coef06, coef07), v(coef08, coef08, coef10, coef11)
    fac3, fac4, fac5 = v(coef12, coef12, coef14, coef15), v(coef16, coef16, coef18, coef19), v(coef20, coef20, coef22, coef23)
    vec0, vec1, vec2, vec3 = v(m10, m00, m00, m00), v(m11, m01, m01, m01), v(m12, m02, m02, m02), v(m13, m03, m03, m03)
    inv0 = vec1 * fac0 - vec2 * fac1 + vec3 * fac2
    inv1 = vec0 * fac0 - vec2 * fac3 + vec3 * fac4
    inv2 = vec0 * fac1 - vec1 * fac3 + vec3 * fac5
    inv3 = vec0 * fac2 - vec1 * fac4 + vec2 * fac5
    sign_a, sign_b = v(1, -1, 1, -1), v(-1, 1, -1, 1)
    inv = np.stack([inv0 * sign_a, inv1 * sign_b, inv2 * sign_a, inv3 * sign_b]).astype(np.float32)
    col0 = v(inv[0, 0], inv[1, 0], inv[2, 0], inv[3, 0])
    d = m[0] * col0
    det = f32(f32(f32(d[0] + d[1]) + d[2]) + d[3])
    return (inv * (f32(1.0) / det)).astype(np.float32)


def _mat_vec(m, x, y, z, w):
    """glam Mat4 * Vec4 for arrays of components: ((c0*x + c1*y) + c2*z) + c3*w."""
    return [((m[0, r] * x + m[1, r] * y) + m[2, r] * z) + m[3, r] * w for r in range(4)]


class Camera:
    def __init__(self, width: int, height: int, fov_deg: float, eye, look_at, up=(0.0, 1.0, 0.0), fov_is_radians=False):
        self.width, self.height = int(width), int(height)
        self.eye = np.asarray(eye, dtype=np.float32)
        fov = f32(fov_deg) if fov_is_radians else f32(fov_deg) * (f32(math.pi) / f32(180.0))
        aspect = f32(width) / f32(height)
        self.proj_inv = mat4_inverse(perspective_infinite_reverse_rh(fov, aspect, 0.01))
        self.view_inv = mat4_inverse(look_at_rh(self.eye, look_at, up))

    def directions(self, px, py):
        """px, py: float32 pixel coordinates (already jittered if wanted) -> unit directions (n,3)."""
        one, two = f32(1.0), f32(2.0)
        u = px / f32(self.width)
        v = one - py / f32(self.height)
        nx, ny = u * two - one, v * two - one
        w1 = np.ones_like(nx)
        vs = _mat_vec(self.proj_inv, nx, ny, w1, w1)
        vs = [c / vs[3] for c in vs]
        ws = _mat_vec(self.view_inv, *vs)
        d = np.stack([ws[0] - self.eye[0], ws[1] - self.eye[1], ws[2] - self.eye[2]], axis=1).astype(np.float32)
        return _norm3(d)


def primary_rays(cam: Camera, tmax=f32(3.4028235e38), column_major=False):
    """One ray per pixel, `Ray::new(eye, direction, 0.0, f32::MAX)` (examples/obj_cwbvh.rs:94-104)."""
    i = np.arange(cam.width * cam.height)
    if column_major:  # tests/mod.rs:154-156 iterates x outer, y inner
        px, py = (i // cam.height), (i % cam.height)
    else:
        px, py = (i % cam.width), (i // cam.width)
    d = cam.directions(px.astype(np.float32), py.astype(np.float32))
    return make_rays(np.broadcast_to(cam.eye, d.shape), d, 0.0, tmax)


def kitchen_camera(width: int) -> Camera:
    """examples/obj_cwbvh.rs:70-81."""
    height = int(f32(width) * f32(0.5625))
    return Camera(width, height, 90.0, (3.0, 1.5, 1.4), (-3.9, 1.5, -1.7))


def cornell_camera(width=1280, height=720) -> Camera:
    """examples/cornell_box_cwbvh.rs:84-95."""
    return Camera(width, height, 90.0, (0.0, 1.0, 2.1), (0.0, 1.0, 0.0))


def demoscene_camera(width=1280) -> Camera:
    """examples/demoscene.rs:79-104."""
    height = int(f32(width) * f32(0.3711))
    eye = np.array([0.0, 0.0, 1.35], dtype=np.float32)
    return Camera(width, height, 17.0, eye, eye + np.array([0.0, 0.16, -1.0], dtype=np.float32))


def demoscene_primary(cam: Camera, aa_sample: int):
    """examples/demoscene.rs:126-152: AA jitter + depth-of-field fuzz; `Ray::new_inf`."""
    i = np.arange(cam.width * cam.height, dtype=np.uint32)
    fx, fy = i % np.uint32(cam.width), i // np.uint32(cam.width)
    s = np.uint32(aa_sample)
    n0, n512, n1024 = tu.hash_noise(fx, fy, s), tu.hash_noise(fx, fy, s + np.uint32(512)), tu.hash_noise(fx, fy, s + np.uint32(1024))
    ax = n0 * f32(0.5) - f32(0.25)
    ay = n512 * f32(0.5) - f32(0.25)
    d = cam.directions(fx.astype(np.float32) + ax, fy.astype(np.float32) + ay)
    fuzz = np.stack([n0, n512, n1024], axis=1)
    sensor = cam.eye[None, :] + (fuzz * f32(2.0) - f32(1.0)) * f32(0.002)
    focal = cam.eye[None, :] + d * f32(2.4)
    cam_dir = _norm3(focal - sensor)
    return make_rays(sensor.astype(np.float32), cam_dir, 0.0, np.inf)


def diffuse_bounce_rays(rays, hit_t, normals, cam: Camera, aa_sample: int):
    """examples/demoscene.rs:163-178 for the rays that hit (hit_t < f32::MAX). normals: double-sided hit normals.

    Returns (bounce_rays, index of the source ray)."""
    hit = np.nonzero(hit_t < f32(3.4028235e38))[0]
    o = rays[hit, 0:3]
    d = rays[hit, 4:7]
    t = hit_t[hit].astype(np.float32)[:, None]
    hit_p = o + d * t - d * f32(0.01)
    i = hit.astype(np.uint32)
    fx, fy = i % np.uint32(cam.width), i // np.uint32(cam.width)
    s = np.uint32(aa_sample)
    local = tu.cosine_sample_hemisphere(tu.hash_noise(fx, fy, s), tu.hash_noise(fx, fy, s + np.uint32(1024)))
    c0, c1, c2 = tu.build_orthonormal_basis(normals[hit])
    world = (c0 * local[:, 0:1] + c1 * local[:, 1:2]) + c2 * local[:, 2:3]
    return make_rays(hit_p.astype(np.float32), _norm3(world), 0.0, np.inf), hit


def shading_normals(bvh_tris, prim_ids, directions):
    """triangle.rs:20-24 `compute_normal` ((v1-v0)x(v2-v0), normalize_or_zero) of the hit triangle, flipped towards the
    ray (`normal *= normal.dot(-ray.direction).signum()`, examples/demoscene.rs:166-167). prim_ids index bvh_tris
    (CwBvh primitive order); entries >= len(bvh_tris) (misses) give (0,0,0)."""
    t = np.asarray(bvh_tris, dtype=np.float32).reshape(-1, 12)
    ok = np.asarray(prim_ids) < t.shape[0]
    tt = t[np.where(ok, prim_ids, 0)]
    v0, v1, v2 = tt[:, 0:3], tt[:, 4:7], tt[:, 8:11]
    a, b = v1 - v0, v2 - v0
    n = np.stack([a[:, 1] * b[:, 2] - b[:, 1] * a[:, 2], a[:, 2] * b[:, 0] - b[:, 2] * a[:, 0], a[:, 0] * b[:, 1] - b[:, 0] * a[:, 1]], axis=1)
    d = (n[:, 0] * n[:, 0] + n[:, 1] * n[:, 1]) + n[:, 2] * n[:, 2]
    with np.errstate(divide="ignore", invalid="ignore"):
        r = f32(1.0) / np.sqrt(d)
    r = np.where(np.isfinite(r) & (r > 0) & ok, r, f32(0.0)).astype(np.float32)
    n = n * r[:, None]
    dirs = np.asarray(directions, dtype=np.float32)
    s = -((n[:, 0] * dirs[:, 0] + n[:, 1] * dirs[:, 1]) + n[:, 2] * dirs[:, 2])
    return (n * np.where(np.signbit(s), f32(-1.0), f32(1.0))[:, None]).astype(np.float32)


def demoscene_bounce_set(cam: Camera, samples, bvh_tris, trace):
    """The incoherent ray set of examples/demoscene.rs:126-178: for each AA sample the jittered primary rays are traced
    with `trace(rays) -> RayHit array` (closest hit, primitive ids in CwBvh order), and one cosine-hemisphere bounce ray
    leaves every hit point. Returns (bounce rays (m,16) f32, number of primary rays traced)."""
    out, n_primary = [], 0
    for s in samples:
        prim = demoscene_primary(cam, int(s))
        hits = trace(prim)
        n_primary += prim.shape[0]
        nrm = shading_normals(bvh_tris, hits["primitive_id"], prim[:, 4:7])
        b, _ = diffuse_bounce_rays(prim, hits["t"], nrm, cam, int(s))
        out.append(b)
    return (np.concatenate(out, axis=0) if out else np.zeros((0, 16), np.float32)), n_primary
