"""Ray recipes of examples/demoscene.rs generated ON THE DEVICE (torch tensors), for workloads too large for the numpy generators
of `camera.py` (100 M rays): jittered depth-of-field primary rays (demoscene.rs:126-152) and the cosine-hemisphere bounce ray that
leaves every primary hit (demoscene.rs:163-178). Same formulas and operation order as `camera.demoscene_primary` /
`camera.diffuse_bounce_rays`; they only produce INPUT rays (a consumer that needs bit-identical sets on the CPU downloads these).

Rays are emitted as the 32-byte arguments of `Ray::new` -- (n, 8) float32 [ox oy oz tmin | dx dy dz tmax] (ObvhsRayNew) -- which the
traversal entry points accept directly.
"""
from __future__ import annotations

import math

import numpy as np
import torch

_M = 0xFFFFFFFF
_INV_U32 = float(np.float32(1.0) / np.float32(0xFFFFFFFF))
_TAU = float(np.float32(6.2831855))
_F32_MAX = 3.4028234663852886e38


def _uhash(x):  # src/test_util.rs:9-17 on int64 lanes holding u32 values
    x = x ^ (x >> 16)
    x = (x * 0x7FEB352D) & _M
    x = x ^ (x >> 15)
    x = (x * 0x846CA68B) & _M
    return x ^ (x >> 16)


def hash_noise(x, y, frame: int):
    """src/test_util.rs:33-46: unormf(uhash2(x, (y << 11) + frame)); x, y int64 tensors of pixel coordinates."""
    b = ((y << 11) + int(frame)) & _M
    h = _uhash(((x * 1597334673) & _M) ^ ((b * 3812015801) & _M))
    return h.to(torch.float32) * _INV_U32


def _norm3(v):
    d = (v[:, 0] * v[:, 0] + v[:, 1] * v[:, 1]) + v[:, 2] * v[:, 2]
    return v / torch.sqrt(d)[:, None]


class DeviceCamera:
    """`camera.Camera` with its two inverse matrices on the device."""

    def __init__(self, cam, device):
        self.width, self.height = cam.width, cam.height
        self.eye = torch.from_numpy(cam.eye).to(device)
        self.proj_inv = [[float(x) for x in col] for col in cam.proj_inv]  # m[col][row], python floats that are exact f32
        self.view_inv = [[float(x) for x in col] for col in cam.view_inv]
        self.device = device

    @staticmethod
    def _mat_vec(m, x, y, z, w):
        return [((m[0][r] * x + m[1][r] * y) + m[2][r] * z) + m[3][r] * w for r in range(4)]

    def directions(self, px, py):
        u = px / float(self.width)
        v = 1.0 - py / float(self.height)
        nx, ny = u * 2.0 - 1.0, v * 2.0 - 1.0
        w1 = torch.ones_like(nx)
        vs = self._mat_vec(self.proj_inv, nx, ny, w1, w1)
        vs = [c / vs[3] for c in vs]
        ws = self._mat_vec(self.view_inv, *vs)
        d = torch.stack([ws[0] - self.eye[0], ws[1] - self.eye[1], ws[2] - self.eye[2]], dim=1)
        return _norm3(d)


def _ray_args(origin, direction, tmin=0.0, tmax=math.inf):
    n = origin.shape[0]
    a = torch.empty((n, 8), dtype=torch.float32, device=origin.device)
    a[:, 0:3] = origin
    a[:, 3] = tmin
    a[:, 4:7] = direction
    a[:, 7] = tmax
    return a


def demoscene_primary(cam: DeviceCamera, aa_sample: int):
    """examples/demoscene.rs:126-152 for every pixel of `cam`: (n, 8) Ray::new records, pixel-major."""
    i = torch.arange(cam.width * cam.height, dtype=torch.int64, device=cam.device)
    fx, fy = i % cam.width, i // cam.width
    n0, n512, n1024 = hash_noise(fx, fy, aa_sample), hash_noise(fx, fy, aa_sample + 512), hash_noise(fx, fy, aa_sample + 1024)
    ax = n0 * 0.5 - 0.25
    ay = n512 * 0.5 - 0.25
    d = cam.directions(fx.to(torch.float32) + ax, fy.to(torch.float32) + ay)
    fuzz = torch.stack([n0, n512, n1024], dim=1)
    sensor = cam.eye[None, :] + (fuzz * 2.0 - 1.0) * 0.002
    focal = cam.eye[None, :] + d * 2.4
    return _ray_args(sensor, _norm3(focal - sensor))


def _cosine_sample_hemisphere(ux, uy):  # src/test_util.rs:63-72
    r = torch.sqrt(ux)
    theta = uy * _TAU
    z = torch.sqrt(torch.clamp_min(1.0 - ux, 0.0))
    return torch.stack([r * torch.cos(theta), r * torch.sin(theta), z], dim=1)


def _orthonormal_basis(n):  # src/test_util.rs:50-61
    sign = torch.where(torch.signbit(n[:, 2]), -1.0, 1.0).to(torch.float32)
    a = -1.0 / (sign + n[:, 2])
    b = n[:, 0] * n[:, 1] * a
    c0 = torch.stack([1.0 + sign * n[:, 0] * n[:, 0] * a, sign * b, -sign * n[:, 0]], dim=1)
    c1 = torch.stack([b, sign + n[:, 1] * n[:, 1] * a, -n[:, 1]], dim=1)
    return c0, c1, n


def shading_normals(rt_tris, prim_ids, directions):
    """Hit normal flipped towards the ray (triangle.rs:20-24, demoscene.rs:166-167). rt_tris: the tree's (n, 16) float32
    RtTriangle records {v0, e1 = v0 - v1, e2 = v2 - v0, ng = e1 x e2}; (v1 - v0) x (v2 - v0) = -ng and the flip makes the sign moot."""
    ng = -rt_tris[prim_ids.to(torch.int64), 12:15]
    d = (ng[:, 0] * ng[:, 0] + ng[:, 1] * ng[:, 1]) + ng[:, 2] * ng[:, 2]
    r = 1.0 / torch.sqrt(d)
    r = torch.where(torch.isfinite(r) & (r > 0), r, torch.zeros_like(r))
    n = ng * r[:, None]
    s = -((n[:, 0] * directions[:, 0] + n[:, 1] * directions[:, 1]) + n[:, 2] * directions[:, 2])
    return n * torch.where(torch.signbit(s), -1.0, 1.0).to(torch.float32)[:, None]


def diffuse_bounce(cam: DeviceCamera, aa_sample: int, primary, hits, rt_tris):
    """examples/demoscene.rs:163-178: one cosine-hemisphere ray from every primary hit. primary (n, 8) records, hits (n, 4) int32
    RayHit rows (primitive_id, -, -, t bits) -> (m, 8) records of the m rays that hit, in pixel order."""
    t_all = hits[:, 3].view(torch.float32)
    hit = torch.nonzero(t_all < _F32_MAX, as_tuple=False)[:, 0]
    o, d = primary[hit, 0:3], primary[hit, 4:7]
    t = t_all[hit][:, None]
    hit_p = o + d * t - d * 0.01
    fx, fy = hit % cam.width, hit // cam.width
    local = _cosine_sample_hemisphere(hash_noise(fx, fy, aa_sample), hash_noise(fx, fy, aa_sample + 1024))
    nrm = shading_normals(rt_tris, hits[hit, 0], d)
    c0, c1, c2 = _orthonormal_basis(nrm)
    world = (c0 * local[:, 0:1] + c1 * local[:, 1:2]) + c2 * local[:, 2:3]
    return _ray_args(hit_p, _norm3(world))


def bounce_set(cam: DeviceCamera, bvh, rt_tris, lo: int, hi: int, total: int, max_samples: int = 4096):
    """Rays [lo, hi) of the global incoherent set: bounce rays of AA samples 0, 1, 2, ... concatenated in (sample, pixel) order and
    cut at `total` rays. Every rank walks the same samples (the set is a pure function of the tree) and keeps its own range.
    Returns ((hi - lo, 8) float32 records, AA samples walked, primary rays traced)."""
    out = torch.empty((hi - lo, 8), dtype=torch.float32, device=cam.device)
    hits = torch.empty((cam.width * cam.height, 4), dtype=torch.int32, device=cam.device)
    pos, s, n_primary = 0, 0, 0
    while pos < min(hi, total) and s < max_samples:
        prim = demoscene_primary(cam, s)
        bvh.ray_traverse(prim, out=hits)
        n_primary += prim.shape[0]
        b = diffuse_bounce(cam, s, prim, hits, rt_tris)
        a, e = max(lo, pos), min(hi, pos + b.shape[0])
        if e > a:
            out[a - lo:e - lo] = b[a - pos:e - pos]
        pos += b.shape[0]
        s += 1
    if pos < hi:
        raise RuntimeError(f"bounce_set: only {pos} rays after {s} AA samples, {hi} wanted")
    return out, s, n_primary
