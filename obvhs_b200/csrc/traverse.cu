// traverse.cu -- batched CWBVH ray traversal over triangles (closest hit / miss / all-hit count).
//
// Replaces, for a batch of rays, the reference's per-ray
//   CwBvh::ray_traverse / ray_traverse_miss / ray_traverse_anyhit      src/cwbvh/mod.rs:169-245
//   traverse! state machine                                            src/cwbvh/traverse_macro.rs:59-126
//   CwBvhNode::intersect_ray (SSE2 4-wide x2 slab test)                src/cwbvh/node.rs:86-231, src/cwbvh/simd.rs:17-100
//   Triangle::intersect (Moller-Trumbore)                              src/triangle.rs:35-76
//
// One ray per thread. Every ray performs the reference's exact visit order (highest set bit first for node and
// primitive groups, remainder pushed only when non-empty, primitives of a node drained before the next node test,
// strict t < tmax), so hit primitive ids are bit-exact including exact-t ties. The 80-byte node is
// fetched as five 16-byte ld.global.nc vectors; the 8 child slab tests are unrolled per thread; the octant order
// comes from the bit index (slot ^ oct_inv) packed into the hit mask; the group stack is 32 x uint2 like the
// reference's StackStack<UVec2, 32> (saturating push).
#include "common.cuh"

#ifndef OBVHS_STATIC_ONE_TRI
#define OBVHS_STATIC_ONE_TRI false
#endif
#ifndef OBVHS_PERSISTENT_ONE_TRI
#define OBVHS_PERSISTENT_ONE_TRI true
#endif

namespace {

constexpr float NODE_EPSILON = 0.0001f;  // cwbvh/node.rs:82 / simd.rs:83

struct RayRegs {
    float ox, oy, oz, dx, dy, dz, ix, iy, iz, tmin, tmax;
};

// triangle.rs:35-76; glam sse2 cross/dot operand order, no FMA. The handle keeps its triangles in the
// reference's OWN precomputed form, RtTriangle {v0, e1 = v0 - v1, e2 = v2 - v0, ng = e1 x e2} (rt_triangle.rs:171-183, 64 bytes):
// e1, e2 and ng are the first values Triangle::intersect computes, rounded identically when the handle is filled
// (rt_triangle_of), so RtTriangle::intersect (rt_triangle.rs:199-239) returns bit for bit what Triangle::intersect does while
// the per-test work drops by the 15 operations of the edge / normal set-up.
constexpr int RT_TRI_VEC4 = 4;  // float4 per stored triangle
__device__ __forceinline__ void rt_triangle_of(const float4 a, const float4 b, const float4 c4, float4 (&out)[4]) {
    const float e1x = a.x - b.x, e1y = a.y - b.y, e1z = a.z - b.z;        // v0 - v1
    const float e2x = c4.x - a.x, e2y = c4.y - a.y, e2z = c4.z - a.z;     // v2 - v0
    out[0] = make_float4(a.x, a.y, a.z, 0.f);
    out[1] = make_float4(e1x, e1y, e1z, 0.f);
    out[2] = make_float4(e2x, e2y, e2z, 0.f);
    out[3] = make_float4(e1y * e2z - e2y * e1z, e1z * e2x - e2z * e1x, e1x * e2y - e2x * e1y, 0.f);  // e1 x e2
}
__device__ __forceinline__ float tri_intersect_vals(const float4 a, const float4 e1, const float4 e2, const float4 n, const RayRegs& r) {
    float cx = a.x - r.ox, cy = a.y - r.oy, cz = a.z - r.oz;        // v0 - origin
    float rx = r.dy * cz - cy * r.dz, ry = r.dz * cx - cz * r.dx, rz = r.dx * cy - cx * r.dy;  // d x c
    float inv_det = 1.0f / ((n.x * r.dx + n.y * r.dy) + n.z * r.dz);
    float u = ((rx * e2.x + ry * e2.y) + rz * e2.z) * inv_det;
    float v = ((rx * e1.x + ry * e1.y) + rz * e1.z) * inv_det;
    float w = 1.0f - u - v;
    u32 sign = __float_as_uint(u) | __float_as_uint(v) | __float_as_uint(w);
    bool valid = (inv_det != 0.0f) && ((sign & 0x80000000u) == 0);
    if (valid) {
        float tt = ((n.x * cx + n.y * cy) + n.z * cz) * inv_det;
        if (tt >= r.tmin && tt <= r.tmax) return tt;
    }
    return __int_as_float(0x7f800000);
}
__device__ __forceinline__ float tri_intersect(const float4* __restrict__ tris, u32 id, const RayRegs& r) {
    const float4* t = tris + (size_t)id * RT_TRI_VEC4;
    return tri_intersect_vals(__ldg(t), __ldg(t + 1), __ldg(t + 2), __ldg(t + 3), r);
}
__device__ __forceinline__ float4 as_f4(const uint4 q) {
    return make_float4(__uint_as_float(q.x), __uint_as_float(q.y), __uint_as_float(q.z), __uint_as_float(q.w));
}
__device__ __forceinline__ float tri_intersect_regs(const uint4 q0, const uint4 q1, const uint4 q2, const uint4 q3, const RayRegs& r) {
    return tri_intersect_vals(as_f4(q0), as_f4(q1), as_f4(q2), as_f4(q3), r);
}

// u8 -> f32 without the (quarter-rate) I2F pipe: 0x4B000000 | byte is the float 8388608 + byte, and subtracting 8388608
// is exact, so the result is bit-identical to (float)byte. One PRMT + one FADD, both full-rate.
// `magic` is 0x4B000000 passed as a kernel argument: kept in a register so the byte selector can be the immediate
// operand of PRMT (with a literal magic the compiler rebuilds the selector register for every byte).
template <int J>
__device__ __forceinline__ float byte_f(u32 w, u32 magic) {
    return __uint_as_float(__byte_perm(w, magic, 0x7550 + J)) - 8388608.0f;
}

// The reference's `max(a,b) = a > b ? a : b` (SSE) and fmaxf only differ when an operand is NaN (the sign of a zero cannot
// change the `tmin <= tmax` decision). EXACT = false is used for nodes whose six affine coefficients are finite and rays
// whose tmax is not NaN: then every slab value is finite or +-inf, never NaN, and FMNMX gives the same hit mask.
template <bool EXACT>
__device__ __forceinline__ float tmax2(float a, float b) { return EXACT ? smax(a, b) : fmaxf(a, b); }
template <bool EXACT>
__device__ __forceinline__ float tmin2(float a, float b) { return EXACT ? smin(a, b) : fminf(a, b); }

// q(byte) * ad, rounded once, exactly as the reference's `q * ad` (simd.rs:56-80):
//  EXACT : (u - 2^23) * ad with u = 2^23 + byte               (PRMT, FADD, FMUL)
//  !EXACT: fma(u, ad, c) with c = -(2^23 * ad)                 (PRMT, FFMA)
//          2^23 * ad is exact (a power-of-two scaling; the caller sends overflowing c to the EXACT path), so the fma rounds
//          the exact product byte * ad once: the same float. Only the sign of a zero product can differ (+0 instead of -0
//          when ad < 0 and byte == 0), which no later min/max/compare can observe.
template <bool EXACT, int J>
__device__ __forceinline__ float byte_mul(u32 w, u32 magic, float ad, float c) {
    if (EXACT) return byte_f<J>(w, magic) * ad;
    return __fmaf_rn(__uint_as_float(__byte_perm(w, magic, 0x7550 + J)), ad, c);
}

template <bool EXACT, int H, int J>
__device__ __forceinline__ u32 child_test(const u32 (&xlo)[2], const u32 (&xhi)[2], const u32 (&ylo)[2], const u32 (&yhi)[2], const u32 (&zlo)[2],
                                          const u32 (&zhi)[2], float adx, float ady, float adz, float cx, float cy, float cz, float aox, float aoy,
                                          float aoz, float ray_tmax, u32 child_bits, u32 bit_index, u32 magic) {
    float tminx = byte_mul<EXACT, J>(xlo[H], magic, adx, cx) + aox, tmaxx = byte_mul<EXACT, J>(xhi[H], magic, adx, cx) + aox;
    float tminy = byte_mul<EXACT, J>(ylo[H], magic, ady, cy) + aoy, tmaxy = byte_mul<EXACT, J>(yhi[H], magic, ady, cy) + aoy;
    float tminz = byte_mul<EXACT, J>(zlo[H], magic, adz, cz) + aoz, tmaxz = byte_mul<EXACT, J>(zhi[H], magic, adz, cz) + aoz;
    float tmn = tmax2<EXACT>(tminx, tmax2<EXACT>(tminy, tminz));  // simd.rs:81-84 nesting
    float tmx = tmin2<EXACT>(tmaxx, tmin2<EXACT>(tmaxy, tmaxz));
    tmn = tmax2<EXACT>(tmn, NODE_EPSILON);
    tmx = tmin2<EXACT>(tmx, ray_tmax);
    return (tmn <= tmx) ? ((child_bits >> (8 * J)) & 0xffu) << ((bit_index >> (8 * J)) & 0xffu) : 0u;
}

// The fast path on packed FP32 pairs (sm_100a: fma.rn.f32x2 / add.rn.f32x2 = FFMA2 / FADD2, two independent IEEE operations per
// instruction, so every value is the one the scalar path computes): the near and far plane of one axis of one child share their
// three coefficients, which halves the FMA-pipe instructions of the node test (96 -> 48 of ~250; ptxas emits the scalar coefficients
// as broadcast operands, `FFMA2 R28, R28.F32x2.HI_LO, R41.F32, R52.F32`, so no registers are added: 56). Measured on B200: the
// one-ray-per-thread kernel gains 5 % on the kitchen (7 900 -> 8 330 Mrays/s); the persistent kernel does not move (soup 852, bounce
// 4 790, terrain 2 280 Mrays/s either way): a packed instruction seems to hold its issue slot for two cycles.
#ifndef OBVHS_NODE_F32X2
#define OBVHS_NODE_F32X2 1
#endif
__device__ __forceinline__ unsigned long long pack2(float lo, float hi) {
    unsigned long long r;
    asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
template <int J>
__device__ __forceinline__ void slab2(u32 wlo, u32 whi, u32 magic, unsigned long long ad2, unsigned long long c2, unsigned long long ao2, float& tnear,
                                      float& tfar) {
    const unsigned long long u = pack2(__uint_as_float(__byte_perm(wlo, magic, 0x7550 + J)), __uint_as_float(__byte_perm(whi, magic, 0x7550 + J)));
    unsigned long long r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(u), "l"(ad2), "l"(c2));
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(r), "l"(ao2));
    asm("mov.b64 {%0,%1}, %2;" : "=f"(tnear), "=f"(tfar) : "l"(r));
}
template <int H, int J>
__device__ __forceinline__ u32 child_test2(const u32 (&xlo)[2], const u32 (&xhi)[2], const u32 (&ylo)[2], const u32 (&yhi)[2], const u32 (&zlo)[2],
                                           const u32 (&zhi)[2], unsigned long long adx2, unsigned long long ady2, unsigned long long adz2,
                                           unsigned long long cx2, unsigned long long cy2, unsigned long long cz2, unsigned long long aox2,
                                           unsigned long long aoy2, unsigned long long aoz2, float ray_tmax, u32 child_bits, u32 bit_index, u32 magic) {
    float tminx, tmaxx, tminy, tmaxy, tminz, tmaxz;
    slab2<J>(xlo[H], xhi[H], magic, adx2, cx2, aox2, tminx, tmaxx);
    slab2<J>(ylo[H], yhi[H], magic, ady2, cy2, aoy2, tminy, tmaxy);
    slab2<J>(zlo[H], zhi[H], magic, adz2, cz2, aoz2, tminz, tmaxz);
    float tmn = fmaxf(tminx, fmaxf(tminy, tminz));  // simd.rs:81-84 nesting
    float tmx = fminf(tmaxx, fminf(tmaxy, tmaxz));
    tmn = fmaxf(tmn, NODE_EPSILON);
    tmx = fminf(tmx, ray_tmax);
    return (tmn <= tmx) ? ((child_bits >> (8 * J)) & 0xffu) << ((bit_index >> (8 * J)) & 0xffu) : 0u;
}

template <bool EXACT>
__device__ __forceinline__ u32 node_children(const uint4 q1, const uint4 q2, const uint4 q3, const uint4 q4, bool rdx, bool rdy, bool rdz, float adx,
                                             float ady, float adz, float cx, float cy, float cz, float aox, float aoy, float aoz, float ray_tmax,
                                             u32 oct_inv4, u32 magic) {
    // q2 = {min_x[0..3], min_x[4..7], max_x[0..3], max_x[4..7]}
    const u32 xlo[2] = {rdx ? q2.z : q2.x, rdx ? q2.w : q2.y}, xhi[2] = {rdx ? q2.x : q2.z, rdx ? q2.y : q2.w};
    const u32 ylo[2] = {rdy ? q3.z : q3.x, rdy ? q3.w : q3.y}, yhi[2] = {rdy ? q3.x : q3.z, rdy ? q3.y : q3.w};
    const u32 zlo[2] = {rdz ? q4.z : q4.x, rdz ? q4.w : q4.y}, zhi[2] = {rdz ? q4.x : q4.z, rdz ? q4.y : q4.w};
    u32 hit_mask = 0;
    if (!EXACT && OBVHS_NODE_F32X2) {
        const unsigned long long adx2 = pack2(adx, adx), ady2 = pack2(ady, ady), adz2 = pack2(adz, adz);
        const unsigned long long cx2 = pack2(cx, cx), cy2 = pack2(cy, cy), cz2 = pack2(cz, cz);
        const unsigned long long aox2 = pack2(aox, aox), aoy2 = pack2(aoy, aoy), aoz2 = pack2(aoz, aoz);
#define OBVHS_HALF2(H, M)                                                                                                         \
    {                                                                                                                             \
        const u32 m = (M);                                                                                                        \
        const u32 is_inner = (m & (m << 1)) & 0x10101010u;                                                                        \
        const u32 inner_mask = (is_inner >> 4) * 0xffu;                                                                           \
        const u32 bit_index = (m ^ (oct_inv4 & inner_mask)) & 0x1f1f1f1fu;                                                        \
        const u32 child_bits = (m >> 5) & 0x07070707u;                                                                            \
        hit_mask |= child_test2<H, 0>(xlo, xhi, ylo, yhi, zlo, zhi, adx2, ady2, adz2, cx2, cy2, cz2, aox2, aoy2, aoz2, ray_tmax, child_bits, bit_index, magic); \
        hit_mask |= child_test2<H, 1>(xlo, xhi, ylo, yhi, zlo, zhi, adx2, ady2, adz2, cx2, cy2, cz2, aox2, aoy2, aoz2, ray_tmax, child_bits, bit_index, magic); \
        hit_mask |= child_test2<H, 2>(xlo, xhi, ylo, yhi, zlo, zhi, adx2, ady2, adz2, cx2, cy2, cz2, aox2, aoy2, aoz2, ray_tmax, child_bits, bit_index, magic); \
        hit_mask |= child_test2<H, 3>(xlo, xhi, ylo, yhi, zlo, zhi, adx2, ady2, adz2, cx2, cy2, cz2, aox2, aoy2, aoz2, ray_tmax, child_bits, bit_index, magic); \
    }
        OBVHS_HALF2(0, q1.z)
        OBVHS_HALF2(1, q1.w)
#undef OBVHS_HALF2
        return hit_mask;
    }
#define OBVHS_HALF(H, M)                                                                                                          \
    {                                                                                                                             \
        /* node.rs:207-231 get_child_and_index_bits on 4 bytes at a time */                                                       \
        const u32 m = (M);                                                                                                        \
        const u32 is_inner = (m & (m << 1)) & 0x10101010u;                                                                        \
        const u32 inner_mask = (is_inner >> 4) * 0xffu;                                                                           \
        const u32 bit_index = (m ^ (oct_inv4 & inner_mask)) & 0x1f1f1f1fu;                                                        \
        const u32 child_bits = (m >> 5) & 0x07070707u;                                                                            \
        hit_mask |= child_test<EXACT, H, 0>(xlo, xhi, ylo, yhi, zlo, zhi, adx, ady, adz, cx, cy, cz, aox, aoy, aoz, ray_tmax, child_bits, bit_index, magic); \
        hit_mask |= child_test<EXACT, H, 1>(xlo, xhi, ylo, yhi, zlo, zhi, adx, ady, adz, cx, cy, cz, aox, aoy, aoz, ray_tmax, child_bits, bit_index, magic); \
        hit_mask |= child_test<EXACT, H, 2>(xlo, xhi, ylo, yhi, zlo, zhi, adx, ady, adz, cx, cy, cz, aox, aoy, aoz, ray_tmax, child_bits, bit_index, magic); \
        hit_mask |= child_test<EXACT, H, 3>(xlo, xhi, ylo, yhi, zlo, zhi, adx, ady, adz, cx, cy, cz, aox, aoy, aoz, ray_tmax, child_bits, bit_index, magic); \
    }
    OBVHS_HALF(0, q1.z)
    OBVHS_HALF(1, q1.w)
#undef OBVHS_HALF
    return hit_mask;
}

// node.rs:86-231 + simd.rs:17-100: returns the 32-bit hit mask (hi 8 bits inner children by octant priority,
// lo 24 bits primitive bits). ray_nan: the ray's tmax is NaN (forces the exact-select path).
__device__ __forceinline__ u32 node_intersect(const uint4 q0, const uint4 q1, const uint4 q2, const uint4 q3, const uint4 q4,
                                              const RayRegs& r, u32 oct_inv4, u32 magic) {
    float px = __uint_as_float(q0.x), py = __uint_as_float(q0.y), pz = __uint_as_float(q0.z);
    float ex = __uint_as_float((q0.w & 0xffu) << 23), ey = __uint_as_float(((q0.w >> 8) & 0xffu) << 23),
          ez = __uint_as_float(((q0.w >> 16) & 0xffu) << 23);  // node.rs:269-275
    float adx = ex * r.ix, ady = ey * r.iy, adz = ez * r.iz;
    float aox = (px - r.ox) * r.ix, aoy = (py - r.oy) * r.iy, aoz = (pz - r.oz) * r.iz;
    bool rdx = r.dx < 0.0f, rdy = r.dy < 0.0f, rdz = r.dz < 0.0f;
    // fast path: all coefficients finite, 2^23 * ad included (a huge finite sum that overflows only sends us to the exact
    // path), and tmax not NaN
    float cx = adx * -8388608.0f, cy = ady * -8388608.0f, cz = adz * -8388608.0f;
    float mag = fabsf(cx) + fabsf(cy) + fabsf(cz) + fabsf(aox) + fabsf(aoy) + fabsf(aoz);
    if (mag < __int_as_float(0x7f800000) && r.tmax == r.tmax)
        return node_children<false>(q1, q2, q3, q4, rdx, rdy, rdz, adx, ady, adz, cx, cy, cz, aox, aoy, aoz, r.tmax, oct_inv4, magic);
    return node_children<true>(q1, q2, q3, q4, rdx, rdy, rdz, adx, ady, adz, cx, cy, cz, aox, aoy, aoz, r.tmax, oct_inv4, magic);
}

// ---- coherence vote (auto mode) --------------------------------------------------------------------------------
// Which kernel is faster depends on whether the 32 rays of a warp do similar work. Measured on B200: primary camera rays
// (kitchen) run 25-40 % faster one-ray-per-thread; bounce / random rays run 2-3x faster on the persistent refill kernel.
// The one-ray-per-thread kernel is launched over the whole batch and every CTA votes on its own 128 rays: a warp is coherent
// when every direction is within ~25 degrees of lane 0's and every origin within 2 % of the scene diagonal of lane 0's.
// A CTA with an incoherent warp appends its index to the deferred list and exits; the persistent kernel, launched right
// behind, takes the rays of the deferred CTAs (and exits at once when there are none). No separate probe pass, and a batch
// that mixes camera and bounce rays gets each part on the kernel that suits it.
struct DeferList {
    u32* count;   // number of deferred 128-ray blocks
    u32* blocks;  // their indices, in arrival order
    float max_dist2;
};
__device__ __forceinline__ bool warp_is_coherent(const RayRegs& r, bool valid, float max_dist2) {
    const float ox = __shfl_sync(0xffffffffu, r.ox, 0), oy = __shfl_sync(0xffffffffu, r.oy, 0), oz = __shfl_sync(0xffffffffu, r.oz, 0);
    const float dx = __shfl_sync(0xffffffffu, r.dx, 0), dy = __shfl_sync(0xffffffffu, r.dy, 0), dz = __shfl_sync(0xffffffffu, r.dz, 0);
    const float dot = r.dx * dx + r.dy * dy + r.dz * dz, l2 = r.dx * r.dx + r.dy * r.dy + r.dz * r.dz, l02 = dx * dx + dy * dy + dz * dz;
    const float ex = r.ox - ox, ey = r.oy - oy, ez = r.oz - oz;
    const bool ok = !valid || (dot > 0.0f && dot * dot >= 0.82f * l2 * l02 && (ex * ex + ey * ey + ez * ez) <= max_dist2);
    return __all_sync(0xffffffffu, ok);
}

// ---- per-ray state machines ---------------------------------------------------------------------------------------
// Both kernels below are generic over a TREE policy (CwTree, Bvh2Tree): State, stack element, begin(), step(), store().
// MODE 0 closest hit -> ObvhsRayHit; 1 miss -> u8; 2 all-hit count -> u32
struct RayResult {
    u32 hit_id;
    float hit_t;
    u32 count;
    bool is_miss;
};
// ray.rs:6-12, 34-52
__device__ __forceinline__ float safe_inverse(float x) {
    const float EPS = 1.1920929e-07f;
    if (fabsf(x) <= EPS) return copysignf(1.0f, x) / EPS;
    return 1.0f / x;
}
// Ray `i` of the batch (RayFormat, common.cuh). kind 0: the reference's 64-byte Ray {origin, direction, inv_direction, tmin, tmax}
// (ray.rs:15-30). kind 1: the 32-byte arguments of Ray::new {origin, tmin, direction, tmax}; the constructor (ray.rs:34-52:
// inv_direction = safe_inverse(direction), IEEE division) runs here, where its result is consumed -- no separate pass, half the
// bytes per ray. kind 2: 24 bytes {origin, direction} with one (tmin, tmax) for the whole batch, as every camera / bounce loop of the
// reference's examples constructs its rays (Ray::new(o, d, 0.0, f32::MAX), Ray::new_inf): the bounds come from the constant bank.
__device__ __forceinline__ void ray_load(RayRegs& r, const float4* __restrict__ rays, size_t i, const RayFormat fmt) {
    if (fmt.kind == 2) {
        const float2* p = reinterpret_cast<const float2*>(rays) + 3 * i;
        const float2 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2);
        r = RayRegs{a.x, a.y, b.x, b.y, c.x, c.y, safe_inverse(b.y), safe_inverse(c.x), safe_inverse(c.y), fmt.tmin, fmt.tmax};
    } else if (fmt.kind == 1) {
        const float4 o = __ldg(rays + 2 * i), d = __ldg(rays + 2 * i + 1);
        r = RayRegs{o.x, o.y, o.z, d.x, d.y, d.z, safe_inverse(d.x), safe_inverse(d.y), safe_inverse(d.z), o.w, d.w};
    } else {
        const float4* rp = rays + 4 * i;
        float4 ro = __ldg(rp), rd = __ldg(rp + 1), ri = __ldg(rp + 2), rt = __ldg(rp + 3);
        r = RayRegs{ro.x, ro.y, ro.z, rd.x, rd.y, rd.z, ri.x, ri.y, ri.z, rt.x, rt.y};
    }
}
// Per-lane traversal stack. The lowest SS entries of every lane live in shared memory as [entry][thread]: one 8-byte column per
// lane, so a warp whose lanes sit at 32 different depths still pushes or pops in two conflict-free wavefronts. The same access
// to a local-memory array touches up to 32 different 128-byte lines (local memory interleaves the lanes' words per index).
// Entries beyond SS spill to local memory. SS = 0: plain local array.
constexpr int TRAV_BLOCK = 128;
template <class T, int CAP, int SS>
struct LaneStack {
    T local[CAP - SS > 0 ? CAP - SS : 1];
    static __device__ __forceinline__ T* column() {
        __shared__ T sh[SS * TRAV_BLOCK];
        return sh + threadIdx.x;
    }
    __device__ __forceinline__ void put(u32 i, const T& v) {
        if (i < (u32)SS) column()[i * TRAV_BLOCK] = v;
        else local[i - SS] = v;
    }
    __device__ __forceinline__ T get(u32 i) const {
        if (i < (u32)SS) return column()[i * TRAV_BLOCK];
        return local[i - SS];
    }
};
template <class T, int CAP>
struct LaneStack<T, CAP, 0> {
    T local[CAP];
    __device__ __forceinline__ void put(u32 i, const T& v) { local[i] = v; }
    __device__ __forceinline__ T get(u32 i) const { return local[i]; }
};
// CAP = 0: a heap stack (the reference's HeapStack for max_depth beyond the fixed sizes, faststack.rs:44-47): this lane's slice
// of a global arena, interleaved over the threads of the grid
template <class T>
struct LaneStack<T, 0, 0> {
    T* base;
    u32 stride;
    __device__ __forceinline__ void put(u32 i, const T& v) { base[(size_t)i * stride] = v; }
    __device__ __forceinline__ T get(u32 i) const { return base[(size_t)i * stride]; }
};
__device__ __forceinline__ void result_reset(RayResult& o) {
    o.hit_id = 0xffffffffu;  // RayHit::none(), ray.rs:74-83
    o.hit_t = __int_as_float(0x7f800000);
    o.count = 0;
    o.is_miss = true;
}
// hit8: closest hits as 8-byte {primitive_id, t} records (ObvhsRayHit8) instead of the 16-byte RayHit whose geometry_id / instance_id
// this path never sets (they stay RayHit::none()'s INVALID)
template <int MODE>
__device__ __forceinline__ void result_store(const RayResult& o, void* __restrict__ out, size_t i, u32 hit8) {
    if (MODE == 0) {
        if (hit8) reinterpret_cast<uint2*>(out)[i] = make_uint2(o.hit_id, __float_as_uint(o.hit_t));
        else reinterpret_cast<uint4*>(out)[i] = make_uint4(o.hit_id, 0xffffffffu, 0xffffffffu, __float_as_uint(o.hit_t));
    } else if (MODE == 1) reinterpret_cast<u8*>(out)[i] = o.is_miss ? 1 : 0;
    else reinterpret_cast<u32*>(out)[i] = o.count;
}

// CwBvh: cwbvh/mod.rs:169-245 + traverse_macro.rs:59-126
struct CwTree {
    const uint4* nodes;
    const float4* tris;
    u32 root_group;  // 0x80000000, or 0 for an empty tree (cwbvh/mod.rs:147-151)
    u32 magic;       // 0x4B000000 (see byte_f)
    typedef uint2 StackT;
    static constexpr int STACK = 32;  // StackStack<UVec2, 32>, cwbvh/mod.rs:60
    struct State {
        RayRegs r;
        u32 oct_inv4;
        u32 sp;
        uint2 cur, prim;  // cwbvh/mod.rs:84-120 current_group / primitive_group
        RayResult o;
    };
    template <class Stack>
    __device__ __forceinline__ void bind_stack(Stack&) const {}
    __device__ __forceinline__ void begin(State& st, const float4* __restrict__ rays, size_t i, const RayFormat fmt) const {
        ray_load(st.r, rays, i, fmt);
        // cwbvh/mod.rs:1001-1010
        st.oct_inv4 = (st.r.dx < 0.0f ? 0u : 0x04040404u) | (st.r.dy < 0.0f ? 0u : 0x02020202u) | (st.r.dz < 0.0f ? 0u : 0x01010101u);
        st.sp = 0;
        st.cur = make_uint2(0u, root_group);  // cwbvh/mod.rs:146-165
        st.prim = make_uint2(0u, 0u);
        result_reset(st.o);
    }
    // One turn of the traverse! loop (traverse_macro.rs:59-126) is tri_step() for every primitive of the current group, then
    // node_step(): test one node, pop when both groups are empty. The loop has ONE exit: in miss mode the first hit clears all
    // pending work (cwbvh/mod.rs:216-220 returns there) and the state machine runs out on its own. An early `break`/`goto`
    // out of the primitive loop made ptxas (12.9, sm_100a) share convergence-barrier registers between the primitive loop
    // and the node test, and the persistent kernel then dead-locked on scenes where lanes of one warp sat in both at once.
    // traverse_macro.rs:64-72, one primitive of the current group (highest set bit first)
    template <int MODE, bool COUNT>
    __device__ __forceinline__ void tri_step(State& st, u32& tris_tested) const {
        u32 local = 31u - __clz(st.prim.y);
        st.prim.y &= ~(1u << local);
        u32 pid = st.prim.x + local;
        float t = tri_intersect(tris, pid, st.r);
        if (COUNT) tris_tested++;
        if (MODE == 0) {
            if (t < st.r.tmax) {  // cwbvh/mod.rs:184-189
                st.o.hit_id = pid;
                st.o.hit_t = t;
                st.r.tmax = t;
            }
        } else if (MODE == 1) {
            if (t < st.r.tmax) {
                st.o.is_miss = false;
                st.prim.y = 0;
                st.cur.y = 0;
                st.sp = 0;
            }
        } else {
            if (t < __int_as_float(0x7f800000)) st.o.count++;
        }
    }
    // traverse_macro.rs:76-123 with an empty primitive group: next node of the current group (or nothing), then pop / finish.
    // Returns true when the ray is done.
    template <bool COUNT, class Stack>
    __device__ __forceinline__ bool node_step(State& st, Stack& stack, u32& nodes_visited) const {
        bool done = false;
        st.prim = make_uint2(0u, 0u);
        if (st.cur.y & 0xff000000u) {  // traverse_macro.rs:76-103
            u32 hits_imask = st.cur.y;
            u32 child_index_offset = 31u - __clz(hits_imask);
            u32 child_index_base = st.cur.x;
            st.cur.y &= ~(1u << child_index_offset);
            if (st.cur.y & 0xff000000u) {  // faststack.rs:299-303 saturating push
                stack.put(st.sp, st.cur);
                st.sp = min(st.sp + 1u, 31u);
            }
            u32 slot_index = (child_index_offset - 24u) ^ (st.oct_inv4 & 0xffu);
            u32 relative_index = __popc(hits_imask & ~(0xffffffffu << slot_index));
            const uint4* np = nodes + (size_t)(child_index_base + relative_index) * 5;
            uint4 q0 = __ldg(np), q1 = __ldg(np + 1), q2 = __ldg(np + 2), q3 = __ldg(np + 3), q4 = __ldg(np + 4);
            if (COUNT) nodes_visited++;
            u32 hitmask = node_intersect(q0, q1, q2, q3, q4, st.r, st.oct_inv4, magic);
            st.cur.x = q1.x;                                   // child_base_idx
            st.prim.x = q1.y;                                  // primitive_base_idx
            st.cur.y = (hitmask & 0xff000000u) | (q0.w >> 24);  // | imask
            st.prim.y = hitmask & 0x00ffffffu;
        } else {
            st.cur = make_uint2(0u, 0u);
        }
        if (st.prim.y == 0 && (st.cur.y & 0xff000000u) == 0) {  // traverse_macro.rs:112-123
            if (st.sp == 0) done = true;
            else {
                st.sp--;
                st.cur = stack.get(st.sp);
            }
        }
        return done;
    }
    // One turn with a SINGLE memory round trip for the whole warp (persistent kernel, POLICY 1). step<ONE_TRI> runs the triangle
    // lanes and the node lanes of a warp as two divergent paths, each with its own load -> use dependency, so a turn costs two
    // serialized memory latencies. Here every lane first works out what it needs next -- the next triangle of its group (64 bytes)
    // or the next node (80 bytes; a lane whose group ran dry pops its stack first, in the same turn) -- then the warp issues all
    // the 16-byte loads together, and only then splits into the triangle test and the node test. The per-ray sequence of tests is
    // exactly traverse_macro.rs:59-126 (a ray still drains its primitive group before its next node). Returns true when the ray
    // is done. Measured on B200 (10 M triangles): long-scoreboard stalls per issue 6.2 -> 3.3, issue-slot utilisation 67 -> 77 %,
    // but 12 % more instructions (every lane advances one test per turn instead of up to two): +2 % on the soup, -4 % on the
    // bounce set. Kept as a selectable variant, not the default.
    template <int MODE, bool COUNT, class Stack>
    __device__ __forceinline__ bool fused_turn(State& st, Stack& stack, u32& nodes_visited, u32& tris_tested) const {
        const bool tri = st.prim.y != 0;
        const uint4* addr = nullptr;
        u32 pid = 0;
        bool done = false;
        if (tri) {  // traverse_macro.rs:64-72: highest set bit first
            const u32 local = 31u - __clz(st.prim.y);
            st.prim.y &= ~(1u << local);
            pid = st.prim.x + local;
            addr = reinterpret_cast<const uint4*>(tris) + (size_t)pid * RT_TRI_VEC4;
        } else {
            if ((st.cur.y & 0xff000000u) == 0) {  // traverse_macro.rs:112-123: both groups empty -> pop or finish
                if (st.sp == 0) done = true;
                else {
                    st.sp--;
                    st.cur = stack.get(st.sp);
                }
            }
            if (!done) {  // traverse_macro.rs:76-103
                const u32 hits_imask = st.cur.y;
                const u32 child_index_offset = 31u - __clz(hits_imask);
                const u32 child_index_base = st.cur.x;
                st.cur.y &= ~(1u << child_index_offset);
                if (st.cur.y & 0xff000000u) {  // faststack.rs:299-303 saturating push
                    stack.put(st.sp, st.cur);
                    st.sp = min(st.sp + 1u, 31u);
                }
                const u32 slot_index = (child_index_offset - 24u) ^ (st.oct_inv4 & 0xffu);
                const u32 relative_index = __popc(hits_imask & ~(0xffffffffu << slot_index));
                addr = nodes + (size_t)(child_index_base + relative_index) * 5;
            }
        }
        // the warp's loads, issued together (asm volatile: the compiler must not sink them into the two branches below)
        uint4 q0 = make_uint4(0, 0, 0, 0), q1 = q0, q2 = q0, q3 = q0, q4 = q0;
        if (addr) {
            asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(q0.x), "=r"(q0.y), "=r"(q0.z), "=r"(q0.w) : "l"(addr));
            asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4+16];" : "=r"(q1.x), "=r"(q1.y), "=r"(q1.z), "=r"(q1.w) : "l"(addr));
            asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4+32];" : "=r"(q2.x), "=r"(q2.y), "=r"(q2.z), "=r"(q2.w) : "l"(addr));
            asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4+48];" : "=r"(q3.x), "=r"(q3.y), "=r"(q3.z), "=r"(q3.w) : "l"(addr));
            if (!tri) asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4+64];" : "=r"(q4.x), "=r"(q4.y), "=r"(q4.z), "=r"(q4.w) : "l"(addr));
        }
        if (tri) {
            const float t = tri_intersect_regs(q0, q1, q2, q3, st.r);
            if (COUNT) tris_tested++;
            if (MODE == 0) {
                if (t < st.r.tmax) {  // cwbvh/mod.rs:184-189
                    st.o.hit_id = pid;
                    st.o.hit_t = t;
                    st.r.tmax = t;
                }
            } else if (MODE == 1) {
                if (t < st.r.tmax) {  // cwbvh/mod.rs:216-220: the first hit ends the ray
                    st.o.is_miss = false;
                    done = true;
                }
            } else {
                if (t < __int_as_float(0x7f800000)) st.o.count++;
            }
        } else if (!done) {
            if (COUNT) nodes_visited++;
            const u32 hitmask = node_intersect(q0, q1, q2, q3, q4, st.r, st.oct_inv4, magic);
            st.cur.x = q1.x;                                   // child_base_idx
            st.prim.x = q1.y;                                  // primitive_base_idx
            st.cur.y = (hitmask & 0xff000000u) | (q0.w >> 24);  // | imask
            st.prim.y = hitmask & 0x00ffffffu;
            // nothing hit and nothing stacked: finished (otherwise the next turn pops)
            if (st.prim.y == 0 && (st.cur.y & 0xff000000u) == 0 && st.sp == 0) done = true;
        }
        return done;
    }
    // ONE_TRI: at most one triangle per call, and no node test while triangles of the current group are pending. The per-ray
    // sequence of tests is unchanged (a ray still drains its group before its next node), only the interleaving with the other
    // lanes of the warp differs: in an incoherent warp the `while` form makes all lanes wait for the lane with the most
    // triangles before every node test.
    template <int MODE, bool COUNT, bool ONE_TRI, class Stack>
    __device__ __forceinline__ bool step(State& st, Stack& stack, u32& nodes_visited, u32& tris_tested) const {
        for (bool first = true; st.prim.y != 0 && (!ONE_TRI || first); first = false) tri_step<MODE, COUNT>(st, tris_tested);
        if (ONE_TRI && st.prim.y != 0) return false;  // triangles of this group still pending -> next call
        return node_step<COUNT>(st, stack, nodes_visited);
    }
};

// Bvh2: bvh2/mod.rs:148-334 (ray_traverse / ray_traverse_miss / ray_traverse_anyhit over ray_traverse_dynamic), node AABB test
// aabb.rs:186-206. Nodes are the 32-byte device layout; a sibling pair is 64 contiguous bytes (4 x 16-byte loads).
// CAP is the reference's fixed stack size: fast_stack!(u32, (96, 192), max_depth) -> StackStack<u32, 96 | 192>.
template <int CAP>
struct Bvh2Tree {
    const float4* nodes;
    const float4* tris;
    u32 node_count;
    u32* heap;      // CAP == 0 only: arena of heap_cap entries per thread of the launch
    u32 heap_cap;
    typedef u32 StackT;
    static constexpr int STACK = CAP;
    __device__ __forceinline__ u32 cap_m1() const { return CAP ? (u32)(CAP - 1) : heap_cap - 1u; }
    template <class Stack>
    __device__ __forceinline__ void bind_stack(Stack& stack) const {
        if constexpr (CAP == 0) {
            stack.base = heap + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
            stack.stride = gridDim.x * blockDim.x;
        }
    }
    static constexpr u32 AT_ROOT = 0xffffffffu;
    struct State {
        RayRegs r;
        u32 cur;  // left child of the pair to test next, or AT_ROOT before the root test
        u32 sp;
        RayResult o;
        // one-triangle-per-step form only: the leaf being drained and what is still owed to the current pair
        u32 phase;        // 0 = test the next pair, 1 / 2 = draining the nearer / farther leaf of the pair, 3 = draining a leaf root
        u32 pend_first, pend_count;
        float right_t;    // entry distance of the farther child, compared with tmax AFTER the nearer leaf was drained (:310)
        u32 r_count, r_first, l_first;
        bool go_left;
    };
    __device__ __forceinline__ void begin(State& st, const float4* __restrict__ rays, size_t i, const RayFormat fmt) const {
        ray_load(st.r, rays, i, fmt);
        st.cur = AT_ROOT;
        st.sp = 0;
        st.phase = 0;
        st.pend_first = st.pend_count = 0;
        st.right_t = 0.f;
        st.r_count = st.r_first = st.l_first = 0;
        st.go_left = false;
        result_reset(st.o);
    }
    // aabb.rs:186-206; glam sse2 min/max (a<b?a:b) and max_element/min_element pairing (x,z),(y,z)
    static __device__ __forceinline__ float box_t(const float4 lo, const float4 hi, const RayRegs& r) {
        float t1x = (lo.x - r.ox) * r.ix, t1y = (lo.y - r.oy) * r.iy, t1z = (lo.z - r.oz) * r.iz;
        float t2x = (hi.x - r.ox) * r.ix, t2y = (hi.y - r.oy) * r.iy, t2z = (hi.z - r.oz) * r.iz;
        float mnx = smin(t1x, t2x), mny = smin(t1y, t2y), mnz = smin(t1z, t2z);
        float mxx = smax(t1x, t2x), mxy = smax(t1y, t2y), mxz = smax(t1z, t2z);
        float tmin_n = smax(smax(mnx, mnz), smax(mny, mnz));
        float tmax_n = smin(smin(mxx, mxz), smin(mxy, mxz));
        return (tmax_n >= tmin_n && tmax_n >= 0.0f) ? tmin_n : __int_as_float(0x7f800000);
    }
    // the leaf callbacks of bvh2/mod.rs:155-165, 191-200, 226-231. Returns false when the traversal must halt (miss mode).
    template <int MODE, bool COUNT>
    __device__ __forceinline__ bool leaf(State& st, u32 first, u32 count, u32& tris_tested) const {
        bool go_on = true;
        for (u32 k = 0; k < count; k++) {
            const u32 pid = first + k;
            float t = tri_intersect(tris, pid, st.r);
            if (COUNT) tris_tested++;
            if (MODE == 0) {
                if (t < st.r.tmax) {
                    st.o.hit_id = pid;
                    st.o.hit_t = t;
                    st.r.tmax = t;
                }
            } else if (MODE == 1) {
                if (t < st.r.tmax) {
                    st.o.is_miss = false;
                    go_on = false;
                    count = 0;  // single loop exit (see CwTree::step)
                }
            } else {
                if (t < __int_as_float(0x7f800000)) st.o.count++;
            }
        }
        return go_on;
    }
    // The same loop with at most ONE triangle per call (persistent kernel): a leaf is drained over several calls while the rest of
    // the pair's work (the farther child's `right_t < tmax` test with the tmax the nearer leaf left behind, then descend / push /
    // pop) waits in the state. The sequence of box and triangle tests of a ray is exactly the reference's.
    template <int MODE, bool COUNT, class Stack>
    __device__ __forceinline__ bool step_one_tri(State& st, Stack& stack, u32& nodes_tested, u32& tris_tested) const {
        bool done = false;
        u32 stage = 0;  // after this call's test: 0 = nothing more, 1 = decide the farther child, 2 = descend / push / pop
        if (st.phase != 0) {
            const u32 pid = st.pend_first;
            const float t = tri_intersect(tris, pid, st.r);
            if (COUNT) tris_tested++;
            st.pend_first++;
            st.pend_count--;
            bool halt = false;
            if (MODE == 0) {
                if (t < st.r.tmax) {
                    st.o.hit_id = pid;
                    st.o.hit_t = t;
                    st.r.tmax = t;
                }
            } else if (MODE == 1) {
                if (t < st.r.tmax) {
                    st.o.is_miss = false;
                    halt = true;
                }
            } else {
                if (t < __int_as_float(0x7f800000)) st.o.count++;
            }
            if (halt) done = true;
            else if (st.pend_count == 0) {  // the leaf is finished
                if (st.phase == 3) done = true;
                else stage = st.phase;  // 1 -> the farther child is next, 2 -> finish the pair
                st.phase = 0;
            }
        } else if (st.cur == AT_ROOT) {
            if (node_count == 0) done = true;
            else {
                const float4 lo = __ldg(nodes), hi = __ldg(nodes + 1);
                if (COUNT) nodes_tested++;
                const u32 prim_count = __float_as_uint(lo.w), first_index = __float_as_uint(hi.w);
                if (!(box_t(lo, hi, st.r) < st.r.tmax)) done = true;
                else if (prim_count != 0) {
                    st.phase = 3;
                    st.pend_first = first_index;
                    st.pend_count = prim_count;
                } else st.cur = first_index;
            }
        } else {
            const float4* np = nodes + (size_t)st.cur * 2;
            const float4 llo = __ldg(np), lhi = __ldg(np + 1), rlo = __ldg(np + 2), rhi = __ldg(np + 3);
            if (COUNT) nodes_tested += 2;
            float left_t = box_t(llo, lhi, st.r), right_t = box_t(rlo, rhi, st.r);
            u32 l_count = __float_as_uint(llo.w), l_first = __float_as_uint(lhi.w);
            u32 r_count = __float_as_uint(rlo.w), r_first = __float_as_uint(rhi.w);
            if (left_t > right_t) {  // :294-297
                float tf = left_t; left_t = right_t; right_t = tf;
                u32 tu = l_count; l_count = r_count; r_count = tu;
                tu = l_first; l_first = r_first; r_first = tu;
            }
            st.right_t = right_t;
            st.r_count = r_count;
            st.r_first = r_first;
            st.l_first = l_first;
            const bool hit_left = left_t < st.r.tmax;
            if (hit_left && l_count != 0) {
                st.go_left = false;
                st.phase = 1;
                st.pend_first = l_first;
                st.pend_count = l_count;
            } else {
                st.go_left = hit_left;
                stage = 1;
            }
        }
        bool go_right = false;
        if (stage == 1) {
            const bool hit_right = st.right_t < st.r.tmax;  // ray.tmax may have shrunk in the nearer leaf (:310)
            if (hit_right && st.r_count != 0) {
                st.phase = 2;
                st.pend_first = st.r_first;
                st.pend_count = st.r_count;
            } else {
                go_right = hit_right;
                stage = 2;
            }
        }
        if (stage == 2) {
            if (st.go_left) {
                st.cur = st.l_first;
                if (go_right) {  // :321-324, saturating push (faststack.rs:299-303)
                    stack.put(st.sp, st.r_first);
                    st.sp = min(st.sp + 1u, cap_m1());
                }
            } else if (go_right) {
                st.cur = st.r_first;
            } else if (st.sp == 0) {
                st.o.hit_t = st.r.tmax;  // :326 `hit.t = ray.tmax`
                done = true;
            } else {
                st.sp--;
                st.cur = stack.get(st.sp);
            }
        }
        return done;
    }
    // one iteration of ray_traverse_dynamic's loop (:284-331); the first call performs the root test (:273-282)
    template <int MODE, bool COUNT, bool ONE_TRI, class Stack>
    __device__ __forceinline__ bool step(State& st, Stack& stack, u32& nodes_tested, u32& tris_tested) const {
        if (ONE_TRI) return step_one_tri<MODE, COUNT>(st, stack, nodes_tested, tris_tested);
        bool done = false;
        if (st.cur == AT_ROOT) {
            if (node_count == 0) return true;
            const float4 lo = __ldg(nodes), hi = __ldg(nodes + 1);
            if (COUNT) nodes_tested++;
            const u32 prim_count = __float_as_uint(lo.w), first_index = __float_as_uint(hi.w);
            if (!(box_t(lo, hi, st.r) < st.r.tmax)) done = true;
            else if (prim_count != 0) {
                leaf<MODE, COUNT>(st, first_index, prim_count, tris_tested);
                done = true;
            } else st.cur = first_index;
            return done;
        }
        const float4* np = nodes + (size_t)st.cur * 2;
        float4 llo = __ldg(np), lhi = __ldg(np + 1), rlo = __ldg(np + 2), rhi = __ldg(np + 3);
        if (COUNT) nodes_tested += 2;
        float left_t = box_t(llo, lhi, st.r), right_t = box_t(rlo, rhi, st.r);
        u32 l_count = __float_as_uint(llo.w), l_first = __float_as_uint(lhi.w);
        u32 r_count = __float_as_uint(rlo.w), r_first = __float_as_uint(rhi.w);
        if (left_t > right_t) {  // :294-297
            float tf = left_t; left_t = right_t; right_t = tf;
            u32 tu = l_count; l_count = r_count; r_count = tu;
            tu = l_first; l_first = r_first; r_first = tu;
        }
        const bool hit_left = left_t < st.r.tmax;
        bool go_left = hit_left;
        if (hit_left && l_count != 0) {
            done = !leaf<MODE, COUNT>(st, l_first, l_count, tris_tested);
            go_left = false;
        }
        const bool hit_right = !done && right_t < st.r.tmax;  // ray.tmax may have shrunk in the left leaf (:310)
        bool go_right = hit_right;
        if (hit_right && r_count != 0) {
            done = !leaf<MODE, COUNT>(st, r_first, r_count, tris_tested);
            go_right = false;
        }
        if (done) return true;
        if (go_left) {
            st.cur = l_first;
            if (go_right) {  // :321-324, saturating push (faststack.rs:299-303)
                stack.put(st.sp, r_first);
                st.sp = min(st.sp + 1u, cap_m1());
            }
        } else if (go_right) {
            st.cur = r_first;
        } else {
            if (st.sp == 0) {
                st.o.hit_t = st.r.tmax;  // :326 `hit.t = ray.tmax`
                return true;
            }
            st.sp--;
            st.cur = stack.get(st.sp);
        }
        return false;
    }
};

template <bool COUNT>
__device__ __forceinline__ void trav_flush_counters(unsigned long long* __restrict__ counters, u32 nodes_visited, u32 tris_tested) {
    if (!COUNT) return;
    unsigned long long a = nodes_visited, b = tris_tested;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(counters, a);
        atomicAdd(counters + 1, b);
    }
}

// One ray per thread: the fastest form for coherent batches (primary / shadow rays of neighbouring pixels).
template <class Tree, int MODE, bool COUNT>
__global__ void __launch_bounds__(TRAV_BLOCK) traverse_kernel(const Tree tree, const float4* __restrict__ rays, size_t n, const RayFormat fmt,
                                                              void* __restrict__ out, unsigned long long* __restrict__ counters,
                                                              const DeferList defer) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    u32 nodes_visited = 0, tris_tested = 0;
    typename Tree::State st;
    LaneStack<typename Tree::StackT, Tree::STACK, 0> stack;
    tree.bind_stack(stack);
    const bool valid = i < n;
    if (valid) tree.begin(st, rays, i, fmt);
    bool mine = valid;
    if (defer.count) {  // auto mode: vote, hand the block over to the persistent kernel when one of its warps is incoherent
        const bool coherent = warp_is_coherent(st.r, valid, defer.max_dist2);
        if (!__syncthreads_and(coherent)) {
            if (threadIdx.x == 0) defer.blocks[atomicAdd(defer.count, 1u)] = blockIdx.x;
            mine = false;
        }
    }
    if (mine) {
        while (!tree.template step<MODE, COUNT, OBVHS_STATIC_ONE_TRI>(st, stack, nodes_visited, tris_tested)) {
        }
        result_store<MODE>(st.o, out, i, fmt.hit8);
    }
    trav_flush_counters<COUNT>(counters, nodes_visited, tris_tested);
}

// Persistent-warp variant for incoherent batches: a warp keeps its 32 lanes busy by pulling new rays from a global
// cursor (warp-private chunks of consecutive rays, one atomicAdd per chunk) whenever at least `refill` lanes have finished,
// instead of idling until its longest ray ends (measured on the 10M-triangle soup with the one-ray-per-thread kernel:
// 5.8 of 32 lanes active per issued instruction, issue slots 75 % busy -- divergence-bound, not memory-bound). Every ray
// still runs the reference's exact per-ray state machine, so results and counters are identical to traverse_kernel's.
//
// POLICY 0: every turn of the inner loop runs both halves of the state machine, the lanes with a pending triangle the triangle
//           test and the others the node test (two divergent paths per turn, each with part of the warp).
// POLICY 1: CwTree::fused_turn -- the loads of the whole warp are issued before the warp splits (one memory round trip per turn).
// SS      : stack entries per lane kept in shared memory (LaneStack); the rest spills to local memory.
// MINB    : __launch_bounds__ minimum CTAs per SM (register cap).
// Dead ends measured in round 2 on the 10 M-triangle scenes (soup / bounce / terrain Mrays/s against 856 / 5865 / 2230 for the
// default): running ONE of the two halves per turn for the whole warp, the node test only when >= 8 / 16 / 24 lanes wait for it
// (555 / 732 / 800 on the soup: the waiting lanes cost more turns than the fuller node tests save); holding fetched nodes in
// registers until >= 16 lanes hold one (692: no gain in lanes per instruction, 24 % more instructions); 48 registers for ten
// CTAs per SM (785: spills); the shared-memory short stack alone (857 / 5771 / 2171: removes 2.1 GB of local-memory write-through
// per 2 M rays, no time change).
struct PersistArgs {
    u32 chunk, refill;
    RayFormat fmt;
};
template <class Tree, int MODE, bool COUNT, bool DEFER, int POLICY, int SS, int MINB>
__global__ void __launch_bounds__(TRAV_BLOCK, MINB) traverse_persistent_kernel(const Tree tree, const float4* __restrict__ rays, u32 n,
                                                                               void* __restrict__ out, unsigned long long* __restrict__ counters,
                                                                               u32* __restrict__ next_ray, const PersistArgs pa,
                                                                               const DeferList defer) {
    // auto mode: only the rays of the 128-ray blocks the one-ray-per-thread kernel deferred (slot k of the list covers the
    // virtual indices [128 k, 128 k + 128)); otherwise the whole batch
    const u32 n_rays = n;
    if (DEFER) {  // (a separate instantiation: the extra live values cost the plain kernel a CTA per SM)
        n = __ldcg(defer.count) * 128u;
        if (n == 0) return;
    }
    constexpr bool ONE_TRI = OBVHS_PERSISTENT_ONE_TRI;
    const u32 lane = threadIdx.x & 31u;
    const u32 lt_mask = (1u << lane) - 1u;
    u32 nodes_visited = 0, tris_tested = 0;
    bool active = false, exhausted = false;
    u32 my = 0;
    typename Tree::State st = {};
    LaneStack<typename Tree::StackT, Tree::STACK, SS> stack;
    tree.bind_stack(stack);
    const u32 chunk = pa.chunk;
    // the warp owns [chunk_pos, chunk_end): consecutive rays, so refills stay close to the rays still in flight
    u32 chunk_pos = 0, chunk_end = 0;  // (n + warps * chunk < 2^32: the host splits larger batches)
    for (;;) {
        if (!exhausted || chunk_pos < chunk_end) {
            const u32 idle = __ballot_sync(0xffffffffu, !active);
            if (chunk_pos == chunk_end) {
                u32 base = 0;
                if (lane == 0) base = atomicAdd(next_ray, chunk);
                base = __shfl_sync(0xffffffffu, base, 0);
                if (base + chunk >= n) exhausted = true;
                chunk_pos = min(base, n);
                chunk_end = min(base + chunk, n);
            }
            const u32 take = min(chunk_end - chunk_pos, (u32)__popc(idle));
            if (!active) {
                const u32 rank = __popc(idle & lt_mask);
                if (rank < take) {
                    my = chunk_pos + rank;
                    if (DEFER) my = __ldcg(defer.blocks + (my >> 7)) * 128u + (my & 127u);
                    if (!DEFER || my < n_rays) {  // (the last block of a batch may be partial)
                        tree.begin(st, rays, my, pa.fmt);
                        active = true;
                    }
                }
            }
            chunk_pos += take;
        }
        if (!__any_sync(0xffffffffu, active)) break;
        const u32 min_active = (exhausted && chunk_pos == chunk_end) ? 1u : 33u - pa.refill;
        if constexpr (POLICY == 0) {
            do {
                if (active && tree.template step<MODE, COUNT, ONE_TRI>(st, stack, nodes_visited, tris_tested)) {
                    result_store<MODE>(st.o, out, my, pa.fmt.hit8);
                    active = false;
                }
            } while ((u32)__popc(__ballot_sync(0xffffffffu, active)) >= min_active);
        } else {
            do {
                if (active && tree.template fused_turn<MODE, COUNT>(st, stack, nodes_visited, tris_tested)) {
                    result_store<MODE>(st.o, out, my, pa.fmt.hit8);
                    active = false;
                }
            } while ((u32)__popc(__ballot_sync(0xffffffffu, active)) >= min_active);
        }
    }
    trav_flush_counters<COUNT>(counters, nodes_visited, tris_tested);
}

// examples/obj_cwbvh.rs:63-67: bvh_tris[i] = tris[primitive_indices[i]], stored as RtTriangle::new(v0, v1, v2) (rt_triangle.rs:171-183)
__global__ void __launch_bounds__(256) permute_tris_kernel(const float4* __restrict__ tris, const u32* __restrict__ prim_idx,
                                                           float4* __restrict__ out, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4* src = tris + (size_t)prim_idx[i] * 3;
    float4 rt[4];
    rt_triangle_of(__ldg(src), __ldg(src + 1), __ldg(src + 2), rt);
#pragma unroll
    for (int k = 0; k < 4; k++) out[i * RT_TRI_VEC4 + k] = rt[k];
}

__global__ void make_rays_kernel(const float* __restrict__ od, size_t n, float tmin, float tmax, float4* __restrict__ rays) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* p = od + i * 6;
    float ox = p[0], oy = p[1], oz = p[2], dx = p[3], dy = p[4], dz = p[5];
    rays[i * 4 + 0] = make_float4(ox, oy, oz, 0.f);
    rays[i * 4 + 1] = make_float4(dx, dy, dz, 0.f);
    rays[i * 4 + 2] = make_float4(safe_inverse(dx), safe_inverse(dy), safe_inverse(dz), 0.f);
    rays[i * 4 + 3] = make_float4(tmin, tmax, 0.f, 0.f);
}

}  // namespace

// Persistent-kernel variants (POLICY, SS, MINB) selectable per context: obvhs_cuda_set_option("traverse_variant", "<id>").
// Variant 0 is the default. Only CwTree has fused_turn (POLICY 1); Bvh2 trees always run variant 0.
template <class Tree, int MODE, bool COUNT, bool DEFER, int POLICY, int SS, int MINB>
static int launch_persistent_v(ObvhsContext* ctx, const Tree& tree, const float4* rays, size_t n, const RayFormat& fmt, void* d_out, unsigned long long* c,
                               u32* next, const DeferList& defer) {
    auto kernel = traverse_persistent_kernel<Tree, MODE, COUNT, DEFER, POLICY, SS, MINB>;
    int per_sm = 0;
    CU_TRY(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, TRAV_BLOCK, 0));
    const size_t need = (n + TRAV_BLOCK - 1) / TRAV_BLOCK;
    size_t blocks = (size_t)ctx->sm_count * (per_sm < 1 ? 1 : per_sm);
    if (blocks > need) blocks = need;
    ctx->traverse_resident_lanes = (size_t)ctx->sm_count * (per_sm < 1 ? 1 : per_sm) * TRAV_BLOCK;
    const PersistArgs pa{(u32)ctx->traverse_chunk, (u32)ctx->traverse_refill, fmt};
    kernel<<<(unsigned)blocks, TRAV_BLOCK, 0, ctx->stream>>>(tree, rays, (u32)n, d_out, c, next, pa, defer);
    KERNEL_CHECK(ctx);
    return OBVHS_OK;
}
template <class Tree>
struct HasSplitSteps { static constexpr bool value = false; };
template <>
struct HasSplitSteps<CwTree> { static constexpr bool value = true; };

template <class Tree, int MODE, bool COUNT>
static int launch_persistent_t(ObvhsContext* ctx, const Tree& tree, const float4* rays, size_t n, const RayFormat& fmt, void* d_out, unsigned long long* c,
                               u32* next, const DeferList& defer) {
    if (defer.count) return launch_persistent_v<Tree, MODE, COUNT, true, 0, 0, 8>(ctx, tree, rays, n, fmt, d_out, c, next, defer);
    if constexpr (HasSplitSteps<Tree>::value) {
        switch (ctx->traverse_variant) {
#define OBVHS_V(ID, POLICY, SS, MINB) \
    case ID: return launch_persistent_v<Tree, MODE, COUNT, false, POLICY, SS, MINB>(ctx, tree, rays, n, fmt, d_out, c, next, defer);
            OBVHS_V(1, 0, 8, 9)
            OBVHS_V(2, 1, 0, 9)
            OBVHS_V(3, 1, 8, 9)
            OBVHS_V(4, 0, 0, 10)
#undef OBVHS_V
            default: break;
        }
    }
    return launch_persistent_v<Tree, MODE, COUNT, false, 0, 0, 9>(ctx, tree, rays, n, fmt, d_out, c, next, defer);
}
template <class Tree>
static int launch_persistent(ObvhsContext* ctx, const Tree& tree, const float4* rays, size_t n, const RayFormat& fmt, int mode, void* d_out,
                             unsigned long long* c, u32* next, const DeferList& defer) {
    if (c) {
        if (mode == 0) return launch_persistent_t<Tree, 0, true>(ctx, tree, rays, n, fmt, d_out, c, next, defer);
        if (mode == 1) return launch_persistent_t<Tree, 1, true>(ctx, tree, rays, n, fmt, d_out, c, next, defer);
        return launch_persistent_t<Tree, 2, true>(ctx, tree, rays, n, fmt, d_out, c, next, defer);
    }
    if (mode == 0) return launch_persistent_t<Tree, 0, false>(ctx, tree, rays, n, fmt, d_out, c, next, defer);
    if (mode == 1) return launch_persistent_t<Tree, 1, false>(ctx, tree, rays, n, fmt, d_out, c, next, defer);
    return launch_persistent_t<Tree, 2, false>(ctx, tree, rays, n, fmt, d_out, c, next, defer);
}
template <class Tree>
static int launch_static(ObvhsContext* ctx, const Tree& tree, const float4* rays, size_t n, const RayFormat& fmt, int mode, void* d_out,
                         unsigned long long* c, const DeferList& defer) {
    dim3 block(TRAV_BLOCK), grid(div_up(n, TRAV_BLOCK));
    cudaStream_t s = ctx->stream;
    if (c) {
        if (mode == 0) traverse_kernel<Tree, 0, true><<<grid, block, 0, s>>>(tree, rays, n, fmt, d_out, c, defer);
        else if (mode == 1) traverse_kernel<Tree, 1, true><<<grid, block, 0, s>>>(tree, rays, n, fmt, d_out, c, defer);
        else traverse_kernel<Tree, 2, true><<<grid, block, 0, s>>>(tree, rays, n, fmt, d_out, c, defer);
    } else {
        if (mode == 0) traverse_kernel<Tree, 0, false><<<grid, block, 0, s>>>(tree, rays, n, fmt, d_out, c, defer);
        else if (mode == 1) traverse_kernel<Tree, 1, false><<<grid, block, 0, s>>>(tree, rays, n, fmt, d_out, c, defer);
        else traverse_kernel<Tree, 2, false><<<grid, block, 0, s>>>(tree, rays, n, fmt, d_out, c, defer);
    }
    KERNEL_CHECK(ctx);
    return OBVHS_OK;
}

// Kernel choice shared by both tree types. ctx->traverse_mode: 0 static (one ray per thread), 1 persistent refill, 2 auto
// (every 128-ray block votes on the device, see DeferList; small batches are static).
constexpr size_t AUTO_STATIC_MAX_PRIMS = 262144;
template <class Tree>
static int traverse_dispatch(ObvhsContext* ctx, const Tree& tree, const ObvhsAabb& total_aabb, size_t prim_count, const float4* rays, size_t n,
                             const RayFormat& fmt, int mode, void* d_out, u64* d_counters, bool force_persistent = false) {
    const size_t ray_bytes = fmt.bytes();
    unsigned long long* c = reinterpret_cast<unsigned long long*>(d_counters);
    cudaStream_t s = ctx->stream;
    int tm = ctx->traverse_mode;
    if (force_persistent) tm = 1;  // (heap stacks: the arena is sized for the resident lanes of the persistent kernel)
    if (tm == 2 && n < 16384) tm = 0;
    // Large scenes: even camera rays vary a lot in work per ray (depth complexity, cache misses), and the refill kernel wins
    // regardless of coherence (terrain, jittered primary rays: +3 % at 0.3 M triangles, +16 % at 1 M, +41 % at 3 M, +36 % at
    // 10 M; the 57 k-triangle kitchen is 41 % faster one-ray-per-thread). The vote only decides for small scenes.
    if (tm == 2 && prim_count > AUTO_STATIC_MAX_PRIMS) tm = 1;
    // 32-bit ray indices inside the persistent kernel: batches beyond 2^31 rays are split into several launches
    const size_t MAX_LAUNCH = (size_t)1 << 31;
    const size_t out_elem = mode == 0 ? (fmt.hit8 ? 8 : sizeof(ObvhsRayHit)) : (mode == 1 ? 1 : 4);
    const size_t n_launches = (n + MAX_LAUNCH - 1) / MAX_LAUNCH;
    if (tm == 2 && n_launches > 1) tm = 1;
    const DeferList none{nullptr, nullptr, 0.f};
    if (tm == 0) return launch_static(ctx, tree, rays, n, fmt, mode, d_out, c, none);
    // scratch: [0] deferred block count, [1] unused, [2..] one ray cursor per persistent launch, then the deferred block list
    const size_t n_blocks = tm == 2 ? (n + 127) / 128 : 0;
    DevBuf<u32> scratch;
    CU_TRY(ctx, scratch.alloc(2 + n_launches + n_blocks, s));
    CU_TRY(ctx, cudaMemsetAsync(scratch.p, 0, (2 + n_launches) * sizeof(u32), s));
    if (tm == 1) {
        for (size_t l = 0; l < n_launches; l++) {
            const size_t off = l * MAX_LAUNCH, cnt = n - off < MAX_LAUNCH ? n - off : MAX_LAUNCH;
            ST_TRY(launch_persistent(ctx, tree, reinterpret_cast<const float4*>(reinterpret_cast<const char*>(rays) + off * ray_bytes), cnt, fmt, mode, (char*)d_out + off * out_elem, c, scratch.p + 2 + l, none));
        }
        return OBVHS_OK;
    }
    const float dx = total_aabb.max[0] - total_aabb.min[0], dy = total_aabb.max[1] - total_aabb.min[1],
                dz = total_aabb.max[2] - total_aabb.min[2];
    float diag2 = dx * dx + dy * dy + dz * dz;
    if (!(diag2 > 0.0f) || !(diag2 < 3.0e38f)) diag2 = 3.0e38f;  // unknown scene extent (uploaded tree): directions decide
    const DeferList defer{scratch.p, scratch.p + 2 + n_launches, diag2 * 0.0004f};
    ST_TRY(launch_static(ctx, tree, rays, n, fmt, mode, d_out, c, defer));
    return launch_persistent(ctx, tree, rays, n, fmt, mode, d_out, c, scratch.p + 2, defer);
}

// Smallest slice of a host batch worth its own launch (traverse_common pipelines H2D | traversal | D2H slice by slice): 32 Ki
// rays for the one-ray-per-thread kernel. The persistent kernel keeps sm_count * 9 CTAs of 128 lanes resident and pays off
// when lanes are refilled, while a large slice leaves the GPU idle behind its H2D copy. Measured on 2 M-ray host batches over
// 10 M triangles (Mrays/s at 1x / 2x / 4x the resident lanes per slice, slices alternating between two compute streams): the
// kernel-bound soup 445 / 557 / 599, the copy-bound terrain 1081 / 997 / 843, diffuse bounces 1329 / 1321 / 1340 -> 2x.
size_t traverse_host_chunk_min(const ObvhsContext* ctx, size_t prim_count, bool* persistent) {
    *persistent = ctx->traverse_mode == 1 || (ctx->traverse_mode == 2 && prim_count > AUTO_STATIC_MAX_PRIMS);
    if (ctx->host_slice) return ctx->host_slice < 1024 ? 1024 : ctx->host_slice;  // obvhs_cuda_set_option("host_slice", ...)
    const size_t lanes = ctx->traverse_resident_lanes ? ctx->traverse_resident_lanes : (size_t)ctx->sm_count * 9 * TRAV_BLOCK;
    return *persistent ? lanes * 2 : (size_t)32768;
}

int cwbvh_traverse_device(ObvhsContext* ctx, const ObvhsCwBvh* bvh, const void* d_rays, const RayFormat& fmt, size_t n, int mode, void* d_out,
                          u64* d_counters) {
    if (n == 0) return OBVHS_OK;
    if (bvh->node_count > 0 && bvh->prim_count > 0 && !bvh->bvh_tris) {
        OBVHS_SET_ERR(ctx, "CwBvh has no triangles attached (call obvhs_cuda_cwbvh_set_triangles)");
        return OBVHS_ERR_INVALID_ARG;
    }
    CwTree tree;
    tree.nodes = reinterpret_cast<const uint4*>(bvh->nodes);
    tree.tris = reinterpret_cast<const float4*>(bvh->bvh_tris);
    tree.root_group = bvh->node_count ? 0x80000000u : 0u;  // cwbvh/mod.rs:147-151: empty bvh => nothing to visit
    tree.magic = 0x4B000000u;
    return traverse_dispatch(ctx, tree, bvh->total_aabb, bvh->prim_count, reinterpret_cast<const float4*>(d_rays), n, fmt, mode, d_out, d_counters);
}

int bvh2_traverse_device(ObvhsContext* ctx, const ObvhsBvh2* bvh, const void* d_rays, const RayFormat& fmt, size_t n, int mode, void* d_out,
                         u64* d_counters) {
    if (n == 0) return OBVHS_OK;
    if (bvh->node_count > 0 && bvh->prim_count > 0 && !bvh->bvh_tris) {
        OBVHS_SET_ERR(ctx, "Bvh2 has no triangles attached (call obvhs_cuda_bvh2_set_triangles)");
        return OBVHS_ERR_INVALID_ARG;
    }
    ObvhsAabb unknown = {};  // the Bvh2 handle does not keep the scene box: the probe then judges directions only
    const float4* rays = reinterpret_cast<const float4*>(d_rays);
    if (bvh->max_depth <= 96) {  // fast_stack!(u32, (96, 192), self.max_depth, ...) bvh2/mod.rs:166
        Bvh2Tree<96> tree{reinterpret_cast<const float4*>(bvh->nodes), reinterpret_cast<const float4*>(bvh->bvh_tris), (u32)bvh->node_count, nullptr, 0u};
        return traverse_dispatch(ctx, tree, unknown, bvh->prim_count, rays, n, fmt, mode, d_out, d_counters);
    }
    if (bvh->max_depth <= 192) {
        Bvh2Tree<192> tree{reinterpret_cast<const float4*>(bvh->nodes), reinterpret_cast<const float4*>(bvh->bvh_tris), (u32)bvh->node_count, nullptr, 0u};
        return traverse_dispatch(ctx, tree, unknown, bvh->prim_count, rays, n, fmt, mode, d_out, d_counters);
    }
    // beyond 192 the reference allocates HeapStack::new_with_capacity(max_depth) per call: here one arena for the launch, max_depth
    // entries for each of the (at most sm_count * 16 * 128) lanes the persistent kernel keeps resident
    DevBuf<u32> heap;
    CU_TRY(ctx, heap.alloc((size_t)ctx->sm_count * 16 * TRAV_BLOCK * bvh->max_depth, ctx->stream));
    Bvh2Tree<0> tree{reinterpret_cast<const float4*>(bvh->nodes), reinterpret_cast<const float4*>(bvh->bvh_tris), (u32)bvh->node_count, heap.p,
                     (u32)bvh->max_depth};
    return traverse_dispatch(ctx, tree, unknown, bvh->prim_count, rays, n, fmt, mode, d_out, d_counters, true);
}

// bvh_tris[i] = tris[primitive_indices[i]] for a Bvh2 (examples/demoscene.rs:66-70)
int bvh2_permute_tris_device(ObvhsContext* ctx, ObvhsBvh2* bvh, const ObvhsTriangle* d_tris, size_t n_tris) {
    if (bvh->bvh_tris) {
        obvhs_result_free(bvh->owner, bvh->bvh_tris);
        bvh->bvh_tris = nullptr;
    }
    const size_t n = bvh->prim_count;
    if (n == 0) return OBVHS_OK;
    (void)n_tris;  // indices are < n_tris by construction of the builders; uploaded trees are the caller's responsibility
    CU_TRY(ctx, obvhs_result_alloc(ctx, (void**)&bvh->bvh_tris, n * OBVHS_RT_TRIANGLE_BYTES));
    permute_tris_kernel<<<div_up(n, 256), 256, 0, ctx->stream>>>(reinterpret_cast<const float4*>(d_tris), bvh->primitive_indices,
                                                                reinterpret_cast<float4*>(bvh->bvh_tris), n);
    KERNEL_CHECK(ctx);
    return OBVHS_OK;
}

int cwbvh_permute_tris_device(ObvhsContext* ctx, ObvhsCwBvh* bvh, const ObvhsTriangle* d_tris, size_t n) {
    // with spatial splits primitive_indices names some triangles several times (cwbvh/mod.rs:752-755)
    if (n != bvh->prim_count && !(bvh->uses_spatial_splits && n <= bvh->prim_count)) {
        OBVHS_SET_ERR(ctx, "set_triangles: %zu triangles for a CwBvh over %zu primitives", n, bvh->prim_count);
        return OBVHS_ERR_INVALID_ARG;
    }
    if (bvh->bvh_tris) {
        obvhs_result_free(bvh->owner, bvh->bvh_tris);
        bvh->bvh_tris = nullptr;
    }
    n = bvh->prim_count;
    if (n == 0) return OBVHS_OK;
    CU_TRY(ctx, obvhs_result_alloc(ctx, (void**)&bvh->bvh_tris, n * OBVHS_RT_TRIANGLE_BYTES));
    permute_tris_kernel<<<div_up(n, 256), 256, 0, ctx->stream>>>(reinterpret_cast<const float4*>(d_tris), bvh->primitive_indices,
                                                                reinterpret_cast<float4*>(bvh->bvh_tris), n);
    KERNEL_CHECK(ctx);
    return OBVHS_OK;
}

int make_rays_device(ObvhsContext* ctx, const float* d_od, size_t n, float tmin, float tmax, ObvhsRay* d_rays) {
    if (n == 0) return OBVHS_OK;
    make_rays_kernel<<<div_up(n, 256), 256, 0, ctx->stream>>>(d_od, n, tmin, tmax, reinterpret_cast<float4*>(d_rays));
    KERNEL_CHECK(ctx);
    return OBVHS_OK;
}
