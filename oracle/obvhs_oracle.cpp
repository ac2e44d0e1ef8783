// obvhs_oracle.cpp -- TEST INFRASTRUCTURE ONLY (see obvhs_oracle.h for the header comment and parity status).
//
// CPU restatement of the obvhs 0.3.1 hot path. Every function cites the reference file:line it follows
// (paths relative to the reference checkout). Float rules:
//   * glam Vec3A::min/max == _mm_min_ps/_mm_max_ps: min(a,b) = a<b ? a : b, max(a,b) = a>b ? a : b
//     (second operand on ties / NaN). Never std::fmin/fmax.
//   * no FMA contraction: build with -ffp-contract=off.
//   * dot3 = (x*x' + y*y') + z*z'; cross = two rounded products and one rounded subtract per lane.
//   * Rust `as u32/u8` casts saturate and map NaN to 0.
//   * log2f/exp2f come from the platform libm, as Rust std's f32::log2/exp2 do on linux-gnu.
// Sort tie rule: the reference sorts are unstable; the contract used by oracle AND GPU is
// "ties by ascending original index" (a stable sort), one of the valid outputs of the reference.
#include "obvhs_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <vector>

#if defined(__SSE2__)
#include <emmintrin.h>
#endif
#ifdef _OPENMP
#include <omp.h>
#endif

typedef uint8_t u8;
typedef uint32_t u32;
typedef uint64_t u64;
typedef int64_t i64;
typedef unsigned __int128 u128;

static_assert(sizeof(OrcAabb) == 32, "Aabb");
static_assert(sizeof(OrcTriangle) == 48, "Triangle");
static_assert(sizeof(OrcBvh2Node) == 48, "Bvh2Node");
static_assert(sizeof(OrcCwBvhNode) == 80, "CwBvhNode");
static_assert(sizeof(OrcRay) == 64, "Ray");
static_assert(sizeof(OrcRayHit) == 16, "RayHit");

// ---------------------------------------------------------------------------------------------------------
// glam-faithful scalar helpers
// ---------------------------------------------------------------------------------------------------------
static inline float smin(float a, float b) { return a < b ? a : b; }  // _mm_min_ps(a,b)
static inline float smax(float a, float b) { return a > b ? a : b; }  // _mm_max_ps(a,b)

struct V3 {
    float x, y, z;
};
static inline V3 v3(const float* p) { return V3{p[0], p[1], p[2]}; }
static inline V3 operator-(V3 a, V3 b) { return V3{a.x - b.x, a.y - b.y, a.z - b.z}; }
static inline V3 operator+(V3 a, V3 b) { return V3{a.x + b.x, a.y + b.y, a.z + b.z}; }
static inline V3 operator*(V3 a, V3 b) { return V3{a.x * b.x, a.y * b.y, a.z * b.z}; }
static inline V3 operator*(V3 a, float s) { return V3{a.x * s, a.y * s, a.z * s}; }
// glam sse2 dot3_in_x: (x*x' + y*y') + z*z'
static inline float dot(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
// glam sse2 Vec3A::cross: (lhs.zxy*rhs - lhs*rhs.zxy).zxy
static inline V3 cross(V3 a, V3 b) {
    return V3{a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y};
}

// aabb.rs:84-89
static inline OrcAabb aabb_union(const OrcAabb& a, const OrcAabb& b) {
    OrcAabb r;
    for (int i = 0; i < 3; i++) {
        r.min[i] = smin(a.min[i], b.min[i]);
        r.max[i] = smax(a.max[i], b.max[i]);
    }
    r._p0 = 0.f;
    r._p1 = 0.f;
    return r;
}
// aabb.rs:76-80
static inline void aabb_extend(OrcAabb& a, const float* p) {
    for (int i = 0; i < 3; i++) {
        a.min[i] = smin(a.min[i], p[i]);
        a.max[i] = smax(a.max[i], p[i]);
    }
}
// aabb.rs:151-154
static inline float half_area(const OrcAabb& a) {
    float dx = a.max[0] - a.min[0], dy = a.max[1] - a.min[1], dz = a.max[2] - a.min[2];
    return (dx + dy) * dz + dx * dy;
}
// aabb.rs:166-171
static inline OrcAabb aabb_empty() {
    OrcAabb a;
    for (int i = 0; i < 3; i++) {
        a.min[i] = 3.40282347e+38f;
        a.max[i] = -3.40282347e+38f;
    }
    a._p0 = a._p1 = 0.f;
    return a;
}
static inline bool aabb_eq(const OrcAabb& a, const OrcAabb& b) {  // Vec3A PartialEq (value compare, -0 == +0)
    for (int i = 0; i < 3; i++)
        if (!(a.min[i] == b.min[i]) || !(a.max[i] == b.max[i])) return false;
    return true;
}

static inline u32 f2u(float f) {
    u32 u;
    memcpy(&u, &f, 4);
    return u;
}
static inline float u2f(u32 u) {
    float f;
    memcpy(&f, &u, 4);
    return f;
}
// Rust `f as u32` / `as u64` / `as u8`: saturating, NaN -> 0
static inline u32 sat_u32(double p) {
    if (!(p > 0.0)) return 0;
    if (p >= 4294967296.0) return 0xffffffffu;
    return (u32)p;
}
static inline u64 sat_u64(double p) {
    if (!(p > 0.0)) return 0;
    if (p >= 18446744073709551616.0) return ~0ull;
    return (u64)p;
}
static inline u8 sat_u8(float p) {
    if (!(p > 0.0f)) return 0;
    if (p >= 255.0f) return 255;
    return (u8)p;
}

// ---------------------------------------------------------------------------------------------------------
// triangle.rs
// ---------------------------------------------------------------------------------------------------------
// triangle.rs:28-30 + aabb.rs:59-66 (from_points: first point, then extend)
static inline OrcAabb tri_aabb(const OrcTriangle& t) {
    OrcAabb a;
    for (int i = 0; i < 3; i++) a.min[i] = a.max[i] = t.v0[i];
    a._p0 = a._p1 = 0.f;
    aabb_extend(a, t.v1);
    aabb_extend(a, t.v2);
    return a;
}

// triangle.rs:35-76 (Moller-Trumbore, two sided, sign-bit validity test)
static inline float tri_intersect(const OrcTriangle& tri, const OrcRay& ray) {
    V3 v0 = v3(tri.v0), v1 = v3(tri.v1), v2 = v3(tri.v2);
    V3 e1 = v0 - v1;
    V3 e2 = v2 - v0;
    V3 n = cross(e1, e2);
    V3 o = v3(ray.origin), d = v3(ray.direction);
    V3 c = v0 - o;
    V3 r = cross(d, c);
    float inv_det = 1.0f / dot(n, d);
    float u = dot(r, e2) * inv_det;
    float v = dot(r, e1) * inv_det;
    float w = 1.0f - u - v;
    u32 hit = f2u(u) | f2u(v) | f2u(w);
    bool valid = (inv_det != 0.0f) && ((hit & 0x80000000u) == 0);
    if (valid) {
        float t = dot(n, c) * inv_det;
        if (t >= ray.tmin && t <= ray.tmax) return t;
    }
    return INFINITY;
}

// triangle.rs:20-24: e1.cross(e2).normalize_or_zero(); glam normalize_or_zero: rcp = 1/sqrt(dot); if rcp finite && >0
static inline V3 tri_normal(const OrcTriangle& tri) {
    V3 e1 = v3(tri.v1) - v3(tri.v0);
    V3 e2 = v3(tri.v2) - v3(tri.v0);
    V3 c = cross(e1, e2);
    float rcp = 1.0f / sqrtf(dot(c, c));
    if (isfinite(rcp) && rcp > 0.0f) return c * rcp;
    return V3{0.f, 0.f, 0.f};
}

// ray.rs:6-12
static inline float safe_inverse(float x) {
    const float EPS = 1.1920929e-07f;  // f32::EPSILON
    if (fabsf(x) <= EPS) return copysignf(1.0f, x) / EPS;
    return 1.0f / x;
}

// ---------------------------------------------------------------------------------------------------------
// ploc/morton.rs
// ---------------------------------------------------------------------------------------------------------
// morton.rs:35-44
static inline u64 split_by_3_u64(u32 a) {
    u64 x = (u64)a & 0x1fffffull;
    x = (x | x << 32) & 0x1f00000000ffffull;
    x = (x | x << 16) & 0x1f0000ff0000ffull;
    x = (x | x << 8) & 0x100f00f00f00f00full;
    x = (x | x << 4) & 0x10c30c30c30c30c3ull;
    x = (x | x << 2) & 0x1249249249249249ull;
    return x;
}
// morton.rs:46-58
static inline u64 morton_encode_u64_unorm(const double p[3]) {
    u32 x = sat_u32(p[0] * 2097152.0), y = sat_u32(p[1] * 2097152.0), z = sat_u32(p[2] * 2097152.0);
    return split_by_3_u64(x) | split_by_3_u64(y) << 1 | split_by_3_u64(z) << 2;
}
static inline u128 mk128(u64 hi, u64 lo) { return ((u128)hi << 64) | lo; }
// morton.rs:65-75
static inline u128 split_by_3_u128(u64 a) {
    u128 x = (u128)a & mk128(0, 0x3ffffffffffull);
    x = (x | x << 64) & mk128(0x3ff00000000ull, 0x00000000ffffffffull);
    x = (x | x << 32) & mk128(0x3ff00000000ull, 0xffff00000000ffffull);
    x = (x | x << 16) & mk128(0x30000ff0000ff00ull, 0x00ff0000ff0000ffull);
    x = (x | x << 8) & mk128(0x300f00f00f00f00ull, 0xf00f00f00f00f00full);
    x = (x | x << 4) & mk128(0x30c30c30c30c30cull, 0x30c30c30c30c30c3ull);
    x = (x | x << 2) & mk128(0x924924924924924ull, 0x9249249249249249ull);
    return x;
}
// morton.rs:77-89
static inline u128 morton_encode_u128_unorm(const double p[3]) {
    const double s = 4398046511104.0;  // (1u64 << 42) as f64
    u64 x = sat_u64(p[0] * s), y = sat_u64(p[1] * s), z = sat_u64(p[2] * s);
    return split_by_3_u128(x) | split_by_3_u128(y) << 1 | split_by_3_u128(z) << 2;
}

// ---------------------------------------------------------------------------------------------------------
// containers
// ---------------------------------------------------------------------------------------------------------
struct OrcBvh2 {
    std::vector<OrcBvh2Node> nodes;
    std::vector<u32> primitive_indices;
    std::vector<u32> parents;
    bool children_are_ordered_after_parents = false;
    size_t max_depth = 96;  // bvh2/mod.rs:87 DEFAULT_MAX_STACK_DEPTH
    size_t ploc_iterations = 0;
    size_t last_applied = 0;
    bool uses_spatial_splits = false;  // bvh2/mod.rs:84
};
struct OrcCwBvh {
    std::vector<OrcCwBvhNode> nodes;
    std::vector<u32> primitive_indices;
    OrcAabb total_aabb;
    bool uses_spatial_splits = false;  // cwbvh/mod.rs:54
    std::vector<OrcAabb> exact_node_aabbs;
};

static inline OrcBvh2Node make_node(const OrcAabb& a, u32 prim_count, u32 first_index) {  // bvh2/node.rs:78-87
    OrcBvh2Node n;
    n.aabb = a;
    n.aabb._p0 = n.aabb._p1 = 0.f;
    n.prim_count = prim_count;
    n.first_index = first_index;
    n.meta1 = n.meta2 = 0;
    return n;
}
static inline bool is_leaf(const OrcBvh2Node& n) { return n.prim_count != 0; }
static inline size_t sibling_id(size_t id) { return (id % 2 == 1) ? id + 1 : id - 1; }  // bvh2/node.rs:154-164
static inline size_t left_sibling_id(size_t id) { return (id % 2 == 1) ? id : id - 1; }

static int clamp_threads(int threads) {
#ifdef _OPENMP
    if (threads <= 0) threads = omp_get_max_threads();
    return threads;
#else
    (void)threads;
    return 1;
#endif
}

// ---------------------------------------------------------------------------------------------------------
// stable LSD radix sort on byte digits (stands in for rdst / rayon sorts with the H1 tie rule)
// ---------------------------------------------------------------------------------------------------------
template <class Rec, class GetByte>
static void radix_sort_stable(std::vector<Rec>& a, int levels, int threads, GetByte get_byte) {
    size_t n = a.size();
    if (n < 2) return;
    std::vector<Rec> tmp(n);
    Rec* src = a.data();
    Rec* dst = tmp.data();
    threads = std::max(1, threads);
    if (n < 65536) threads = 1;
    std::vector<size_t> hist((size_t)threads * 256);
    for (int level = 0; level < levels; level++) {
        std::fill(hist.begin(), hist.end(), 0);
        size_t chunk = (n + threads - 1) / threads;
#pragma omp parallel for num_threads(threads) schedule(static, 1)
        for (int t = 0; t < threads; t++) {
            size_t b = (size_t)t * chunk, e = std::min(n, b + chunk);
            size_t* h = &hist[(size_t)t * 256];
            for (size_t i = b; i < e; i++) h[get_byte(src[i], level)]++;
        }
        // skip a level where every key has the same digit
        bool trivial = false;
        for (int d = 0; d < 256; d++) {
            size_t s = 0;
            for (int t = 0; t < threads; t++) s += hist[(size_t)t * 256 + d];
            if (s == n) trivial = true;
        }
        if (trivial) continue;
        size_t run = 0;
        for (int d = 0; d < 256; d++)
            for (int t = 0; t < threads; t++) {
                size_t c = hist[(size_t)t * 256 + d];
                hist[(size_t)t * 256 + d] = run;
                run += c;
            }
#pragma omp parallel for num_threads(threads) schedule(static, 1)
        for (int t = 0; t < threads; t++) {
            size_t b = (size_t)t * chunk, e = std::min(n, b + chunk);
            size_t* h = &hist[(size_t)t * 256];
            for (size_t i = b; i < e; i++) dst[h[get_byte(src[i], level)]++] = src[i];
        }
        std::swap(src, dst);
    }
    if (src != a.data()) memcpy(a.data(), src, n * sizeof(Rec));
}

// ---------------------------------------------------------------------------------------------------------
// PLOC (ploc/mod.rs)
// ---------------------------------------------------------------------------------------------------------
struct Morton {  // ploc/mod.rs:686-708 (Morton128 / Morton64 share this record here)
    u64 lo, hi;
    u32 index;
};

// ploc/mod.rs:187-243 leaf init + total AABB. The serial path (prim_count < 500k or no rayon) is the contract.
static OrcAabb init_leaves(const OrcAabb* aabbs, const u32* indices, size_t n, std::vector<OrcBvh2Node>& cur, int threads) {
    cur.resize(n);
    OrcAabb total = aabb_empty();
    if (threads > 1 && n >= 500000) {
        // ploc/mod.rs:205-231: chunked parallel init, then a serial fold of the per-chunk boxes.
        size_t chunk = (n + threads - 1) / threads;
        std::vector<OrcAabb> local((n + chunk - 1) / chunk, aabb_empty());
#pragma omp parallel for num_threads(threads) schedule(static, 1)
        for (long c = 0; c < (long)local.size(); c++) {
            size_t b = (size_t)c * chunk, e = std::min(n, b + chunk);
            for (size_t i = b; i < e; i++) {
                aabb_extend(local[c], aabbs[i].min);
                aabb_extend(local[c], aabbs[i].max);
                cur[i] = make_node(aabbs[i], 1, indices[i]);
            }
        }
        for (auto& l : local) {
            aabb_extend(total, l.min);
            aabb_extend(total, l.max);
        }
    } else {
        for (size_t i = 0; i < n; i++) {  // ploc/mod.rs:188-194 init_node
            aabb_extend(total, aabbs[i].min);
            aabb_extend(total, aabbs[i].max);
            cur[i] = make_node(aabbs[i], 1, indices[i]);
        }
    }
    return total;
}

// ploc/mod.rs:287-288, 782-785: codes for the nodes in their current order
static void gen_mortons(const std::vector<OrcBvh2Node>& nodes, const OrcAabb& total, int precision, std::vector<Morton>& m,
                        int threads) {
    size_t n = nodes.size();
    m.resize(n);
    double scale[3], offset[3];
    for (int i = 0; i < 3; i++) {
        float diag = total.max[i] - total.min[i];  // aabb.rs:106-108 (f32), then as_dvec3
        scale[i] = 1.0 / (double)diag;
        offset[i] = -(double)total.min[i] * scale[i];
    }
#pragma omp parallel for num_threads(threads) schedule(static) if (threads > 1 && n > 100000)
    for (long i = 0; i < (long)n; i++) {
        const OrcAabb& a = nodes[i].aabb;
        double p[3];
        for (int k = 0; k < 3; k++) {
            float c = (a.max[k] + a.min[k]) * 0.5f;  // aabb.rs:113-115
            p[k] = (double)c * scale[k] + offset[k];
        }
        Morton r;
        r.index = (u32)i;
        if (precision == 128) {
            u128 code = morton_encode_u128_unorm(p);
            r.lo = (u64)code;
            r.hi = (u64)(code >> 64);
        } else {
            r.lo = morton_encode_u64_unorm(p);
            r.hi = 0;
        }
        m[i] = r;
    }
}

// ploc/mod.rs:811-827 with the H1 tie rule (stable by original index)
static void sort_mortons(std::vector<Morton>& m, int precision, int threads) {
    int levels = precision == 128 ? 16 : 8;  // ploc/mod.rs:695,711
    radix_sort_stable(m, levels, threads, [](const Morton& r, int level) -> u8 {
        return level < 8 ? (u8)(r.lo >> (8 * level)) : (u8)(r.hi >> (8 * (level - 8)));
    });
}

// ploc/mod.rs:508-528 / 624-650: window search, `<=` so the last minimum wins. cost(lo,hi) keeps the cache's
// operand order (union(nodes[lower], nodes[higher])), ploc/mod.rs:577-585,641-642.
static inline int8_t find_best_node(size_t index, const OrcBvh2Node* nodes, size_t count, size_t R) {
    size_t best = index;
    float best_cost = INFINITY;
    size_t begin = index - std::min(R, index);
    size_t end = std::min(index + R + 1, count);
    for (size_t other = begin; other < index; other++) {
        float c = half_area(aabb_union(nodes[other].aabb, nodes[index].aabb));
        if (c <= best_cost) {
            best = other;
            best_cost = c;
        }
    }
    for (size_t other = index + 1; other < end; other++) {
        float c = half_area(aabb_union(nodes[index].aabb, nodes[other].aabb));
        if (c <= best_cost) {
            best = other;
            best_cost = c;
        }
    }
    return (int8_t)((i64)best - (i64)index);
}

// ploc/mod.rs:258-503 (REBUILD = false)
static inline bool node_valid(const OrcBvh2Node& n) { return (n.prim_count & 0x80000000u) == 0; }  // bvh2/node.rs:137-139
// ploc/mod.rs:265-503. rebuild == the REBUILD const generic (partial rebuilds): children go into the slot pairs that
// partial_rebuild freed (marked invalid), searched downwards from the end of bvh.nodes (:449-462).
static void build_ploc_from_leaves(OrcBvh2& bvh, std::vector<OrcBvh2Node>& cur, const OrcAabb& total, size_t R, int precision,
                                   size_t search_depth_threshold, int threads, bool rebuild = false) {
    size_t prim_count = cur.size();
    if (prim_count == 0) return;
    size_t nodes_count = 2 * prim_count - 1;
    size_t insert_index;
    if (rebuild) {  // :283-288
        if (bvh.nodes.empty()) return;
        if (bvh.nodes.size() < nodes_count) abort();
        insert_index = bvh.nodes.size() - 1;
    } else {
        bvh.nodes.assign(nodes_count, OrcBvh2Node{});
        insert_index = nodes_count;
    }

    std::vector<Morton> mortons;
    gen_mortons(cur, total, precision, mortons, threads);
    sort_mortons(mortons, precision, threads);
    std::vector<OrcBvh2Node> next(prim_count);
#pragma omp parallel for num_threads(threads) schedule(static) if (threads > 1 && prim_count > 100000)
    for (long i = 0; i < (long)prim_count; i++) next[i] = cur[mortons[i].index];  // ploc/mod.rs:829
    std::swap(cur, next);

    std::vector<int8_t> merge(prim_count);
    size_t depth = 0, count = prim_count;
    while (count > 1) {
        if (R == 1 || depth < search_depth_threshold) {
            // ploc/mod.rs:329-382
            size_t count_m1 = count - 1;
            if (threads > 1 && count >= 300000) {
                size_t chunk = (count_m1 + threads - 1) / threads;
#pragma omp parallel for num_threads(threads) schedule(static, 1)
                for (int t = 0; t < threads; t++) {
                    size_t start = (size_t)t * chunk, e = std::min(count_m1, start + chunk);
                    if (start >= e) continue;
                    float last_cost = start == 0 ? INFINITY : half_area(aabb_union(cur[start - 1].aabb, cur[start].aabb));
                    for (size_t i = start; i < e; i++) {
                        float cost = half_area(aabb_union(cur[i].aabb, cur[i + 1].aabb));
                        merge[i] = last_cost < cost ? -1 : 1;
                        last_cost = cost;
                    }
                }
            } else {
                float last_cost = INFINITY;
                for (size_t i = 0; i < count_m1; i++) {
                    float cost = half_area(aabb_union(cur[i].aabb, cur[i + 1].aabb));
                    merge[i] = last_cost < cost ? -1 : 1;
                    last_cost = cost;
                }
            }
            merge[count_m1] = -1;
        } else {
            // ploc/mod.rs:383-418
#pragma omp parallel for num_threads(threads) schedule(static) if (threads > 1 && count >= 4000)
            for (long i = 0; i < (long)count; i++) merge[i] = find_best_node((size_t)i, cur.data(), count, R);
        }
        // ploc/mod.rs:420-487 sequential merge sweep
        size_t next_idx = 0, index = 0;
        while (index < count) {
            i64 index_offset = merge[index];
            size_t best_index = (size_t)((i64)index + index_offset);
            if ((i64)best_index + (i64)merge[best_index] != (i64)index) {
                next[next_idx++] = cur[index];
                index++;
                continue;
            }
            if (best_index > index) {
                index++;
                continue;
            }
            OrcBvh2Node left = cur[index];
            OrcBvh2Node right = cur[best_index];
            size_t first_child;
            if (rebuild) {  // :449-462
                for (;;) {
                    OrcBvh2Node& left_slot = bvh.nodes[insert_index - 1];
                    if (!node_valid(left_slot)) {
                        left_slot = left;
                        bvh.nodes[insert_index] = right;
                        first_child = insert_index - 1;
                        insert_index -= 2;
                        break;
                    }
                    insert_index -= 2;
                }
            } else {
                insert_index -= 2;
                bvh.nodes[insert_index] = left;
                bvh.nodes[insert_index + 1] = right;
                first_child = insert_index;
            }
            next[next_idx++] = make_node(aabb_union(left.aabb, right.aabb), 0, (u32)first_child);
            if (R == 1 && index_offset == 1)
                index += 2;
            else
                index += 1;
        }
        std::swap(next, cur);
        count = next_idx;
        depth++;
    }
    bvh.nodes[0] = cur[0];
    bvh.max_depth = std::max<size_t>(96, depth + 1);     // ploc/mod.rs:501
    bvh.children_are_ordered_after_parents = !rebuild;  // ploc/mod.rs:502
    bvh.ploc_iterations = depth;
}

// bvh2/mod.rs:586-619
static void compute_parents(OrcBvh2& bvh) {
    bvh.parents.assign(bvh.nodes.size(), 0);
    for (size_t i = 0; i < bvh.nodes.size(); i++) {
        const OrcBvh2Node& n = bvh.nodes[i];
        if (!is_leaf(n)) {
            bvh.parents[n.first_index] = (u32)i;
            bvh.parents[n.first_index + 1] = (u32)i;
        }
    }
}

// ploc/rebuild.rs:12-43
static void compute_rebuild_path_flags(const OrcBvh2& bvh, const u32* leaves, size_t n_leaves, std::vector<u8>& flags) {
    if (bvh.nodes.size() < 2) return;
    if (bvh.parents.empty()) abort();  // the reference panics: parents must be initialised first
    flags.assign(bvh.nodes.size(), 0);
    for (size_t k = 0; k < n_leaves; k++) {
        size_t index = leaves[k];
        flags[index] = 1;
        while (index > 0) {
            index = bvh.parents[index];
            if (flags[index]) break;
            flags[index] = 1;
        }
    }
}
// ploc/rebuild.rs:137-183
static void rebuild_from_leaves(OrcBvh2& bvh, std::vector<OrcBvh2Node>& cur, bool partial, size_t R, int precision, size_t thr, int threads) {
    if (bvh.nodes.size() < 2) return;
    bool had_parents = !bvh.parents.empty();
    OrcAabb total = bvh.nodes[0].aabb;  // :153: the (possibly stale) root box only scales the Morton codes
    build_ploc_from_leaves(bvh, cur, total, R, precision, thr, threads, partial);
    if (had_parents) compute_parents(bvh);  // :177-179
}
// ploc/rebuild.rs:56-80
static void full_rebuild(OrcBvh2& bvh, size_t R, int precision, size_t thr, int threads) {
    if (bvh.nodes.size() < 2) return;
    std::vector<OrcBvh2Node> cur;
    for (const OrcBvh2Node& n : bvh.nodes)
        if (is_leaf(n)) cur.push_back(n);
    rebuild_from_leaves(bvh, cur, false, R, precision, thr, threads);
}
// ploc/rebuild.rs:101-135. The reference collects the surviving subtrees in the order of its stack walk and then sorts them
// with an UNSTABLE sort, so the order among equal Morton codes is unspecified there. Tie rule of this restatement (and of
// the CUDA path): equal codes keep ascending NODE INDEX, i.e. the collected nodes are put in node-index order before the
// stable sort.
static void partial_rebuild(OrcBvh2& bvh, const u8* should_remove, size_t R, int precision, size_t thr, int threads) {
    if (bvh.nodes.size() < 2) return;
    std::vector<std::pair<u32, OrcBvh2Node>> collected;
    std::vector<u32> stack;
    stack.push_back(bvh.nodes[0].first_index);
    while (!stack.empty()) {
        u32 left_node_index = stack.back();
        stack.pop_back();
        for (u32 node_index : {left_node_index, left_node_index + 1}) {
            OrcBvh2Node& node = bvh.nodes[node_index];
            if (!should_remove[node_index] || is_leaf(node)) collected.push_back({node_index, node});
            else stack.push_back(node.first_index);
            node.prim_count |= 0x80000000u;  // set_invalid, bvh2/node.rs:143-145
        }
    }
    std::sort(collected.begin(), collected.end(), [](const auto& a, const auto& b) { return a.first < b.first; });
    std::vector<OrcBvh2Node> cur;
    cur.reserve(collected.size());
    for (auto& c : collected) cur.push_back(c.second);
    rebuild_from_leaves(bvh, cur, true, R, precision, thr, threads);
}

// bvh2/mod.rs:527-569
static void refit_all(OrcBvh2& bvh) {
    if (bvh.nodes.empty()) return;
    auto refit_one = [&](size_t id) {
        OrcBvh2Node& n = bvh.nodes[id];
        if (!is_leaf(n)) {
            OrcAabb a = aabb_union(bvh.nodes[n.first_index].aabb, bvh.nodes[n.first_index + 1].aabb);
            n.aabb = a;
        }
    };
    if (bvh.children_are_ordered_after_parents) {
        for (size_t id = bvh.nodes.size(); id-- > 0;) refit_one(id);
    } else {
        std::vector<u32> stack, reverse_stack;
        reverse_stack.reserve(bvh.nodes.size());
        stack.push_back(0);
        reverse_stack.push_back(0);
        while (!stack.empty()) {
            u32 cur = stack.back();
            stack.pop_back();
            const OrcBvh2Node& n = bvh.nodes[cur];
            if (!is_leaf(n)) {
                reverse_stack.push_back(n.first_index);
                reverse_stack.push_back(n.first_index + 1);
                stack.push_back(n.first_index);
                stack.push_back(n.first_index + 1);
            }
        }
        for (size_t k = reverse_stack.size(); k-- > 0;) refit_one(reverse_stack[k]);
    }
}

// ---------------------------------------------------------------------------------------------------------
// Reinsertion (bvh2/reinsertion.rs)
// ---------------------------------------------------------------------------------------------------------
struct Reinsertion {
    u32 from, to;
    float area_diff;
};
struct Candidate {
    u32 node_id;
    float cost;
};

// rdst 0.20 RadixKey for f32 (third-party, not vendored; published algorithm): map the IEEE bits to an unsigned
// key whose order matches the float order: negative -> flip all bits, non-negative -> flip the sign bit.
static inline u32 f32_radix_key(float f) {
    u32 u = f2u(f);
    u32 mask = (u32)((int32_t)u >> 31) | 0x80000000u;
    return u ^ mask;
}

// bvh2/reinsertion.rs:233-334. FastStack semantics from faststack.rs:282-310 / 104-140.
struct PairStack {
    std::vector<std::pair<float, u32>> data;
    size_t index = 0;
    bool saturating = true;
    void init(size_t max_depth) {
        size_t size = max_depth * 2;
        if (size <= 96) {
            data.resize(96);
            saturating = true;
        } else if (size <= 192) {
            data.resize(192);
            saturating = true;
        } else {
            data.resize(size);
            saturating = false;
        }
        index = 0;
    }
    inline void push(float a, u32 b) {
        if (saturating) {
            data[index] = {a, b};
            index = std::min(index + 1, data.size() - 1);
        } else {
            if (index >= data.size()) {
                fprintf(stderr, "oracle: HeapStack overflow\n");
                abort();
            }
            data[index++] = {a, b};
        }
    }
    inline std::pair<float, u32> pop_fast() {
        index = index > 0 ? index - 1 : 0;
        return data[index];
    }
};

static Reinsertion find_reinsertion(const OrcBvh2& bvh, size_t node_id, PairStack& stack) {
    const std::vector<OrcBvh2Node>& nodes = bvh.nodes;
    const std::vector<u32>& parents = bvh.parents;
    Reinsertion best{(u32)node_id, 0, 0.0f};
    float node_area = half_area(nodes[node_id].aabb);
    float parent_area = half_area(nodes[parents[node_id]].aabb);
    float area_diff = parent_area;
    size_t sib = sibling_id(node_id);
    OrcAabb pivot_bbox = nodes[sib].aabb;
    size_t parent_id = parents[node_id];
    size_t pivot_id = parent_id;
    const OrcAabb aabb = nodes[node_id].aabb;
    stack.index = 0;
    for (;;) {
        stack.push(area_diff, (u32)sib);
        while (stack.index != 0) {
            std::pair<float, u32> top = stack.pop_fast();
            float top_area_diff = top.first;
            u32 top_sibling_id = top.second;
            if (top_area_diff - node_area <= best.area_diff) continue;
            const OrcBvh2Node& dst = nodes[top_sibling_id];
            float merged_area = half_area(aabb_union(dst.aabb, aabb));
            float reinsert_area = top_area_diff - merged_area;
            if (reinsert_area > best.area_diff) {
                best.to = top_sibling_id;
                best.area_diff = reinsert_area;
            }
            if (!is_leaf(dst)) {
                float child_area = reinsert_area + half_area(dst.aabb);
                stack.push(child_area, dst.first_index);
                stack.push(child_area, dst.first_index + 1);
            }
        }
        if (pivot_id != parent_id) {
            pivot_bbox = aabb_union(pivot_bbox, nodes[sib].aabb);
            area_diff += half_area(nodes[pivot_id].aabb) - half_area(pivot_bbox);
        }
        if (pivot_id == 0) break;
        sib = sibling_id(pivot_id);
        pivot_id = parents[pivot_id];
    }
    if (best.to == (u32)sibling_id(best.from) || best.to == parents[best.from]) best = Reinsertion{0, 0, 0.0f};
    return best;
}

// bvh2/mod.rs:722-751 (release semantics: early out after two unchanged nodes). full=true walks to the root
// (debug-assertions semantics); tests check both give identical trees.
static bool g_refit_full = false;
static void refit_from_fast(OrcBvh2& bvh, size_t index) {
    int same_count = 0;
    for (;;) {
        OrcBvh2Node& node = bvh.nodes[index];
        if (!is_leaf(node)) {
            OrcAabb na = aabb_union(bvh.nodes[node.first_index].aabb, bvh.nodes[node.first_index + 1].aabb);
            if (aabb_eq(node.aabb, na)) {
                same_count++;
                if (same_count == 2 && !g_refit_full) return;
            }
            node.aabb = na;
        }
        if (index == 0) break;
        index = bvh.parents[index];
    }
}

// bvh2/reinsertion.rs:336-382
static void reinsert_node(OrcBvh2& bvh, size_t from, size_t to) {
    size_t sib = sibling_id(from);
    size_t parent_id = bvh.parents[from];
    OrcBvh2Node sibling_node = bvh.nodes[sib];
    OrcBvh2Node dst_node = bvh.nodes[to];
    bvh.nodes[to].prim_count = 0;  // make_inner
    bvh.nodes[to].first_index = (u32)left_sibling_id(from);
    bvh.nodes[sib] = dst_node;
    bvh.nodes[parent_id] = sibling_node;
    const OrcBvh2Node& s = bvh.nodes[sib];
    if (!is_leaf(s)) {
        bvh.parents[s.first_index] = (u32)sib;
        bvh.parents[s.first_index + 1] = (u32)sib;
    }
    const OrcBvh2Node& p = bvh.nodes[parent_id];
    if (!is_leaf(p)) {
        bvh.parents[p.first_index] = (u32)parent_id;
        bvh.parents[p.first_index + 1] = (u32)parent_id;
    }
    bvh.parents[sib] = (u32)to;
    bvh.parents[from] = (u32)to;
    refit_from_fast(bvh, to);
    refit_from_fast(bvh, parent_id);
}

struct ReinsertionOptimizer {
    std::vector<Candidate> candidates;
    std::vector<Reinsertion> reinsertions;
    std::vector<u8> touched;
    float batch_size_ratio = 0.f;
    int threads = 1;
    size_t applied = 0;

    // reinsertion.rs:121-139
    void find_candidates(OrcBvh2& bvh, size_t node_count) {
        candidates.clear();
        size_t take = std::min(bvh.nodes.size(), node_count * 2);
        for (size_t i = 1; i < take; i++) candidates.push_back(Candidate{(u32)i, half_area(bvh.nodes[i].aabb)});
        radix_sort_stable(candidates, 4, threads,
                          [](const Candidate& c, int level) -> u8 { return (u8)(f32_radix_key(-c.cost) >> (8 * level)); });
    }
    // reinsertion.rs:142-196 (`parallel` flavour: all results kept, non-positive gains skipped at apply time)
    void optimize_candidates(OrcBvh2& bvh, size_t count) {
        std::fill(touched.begin(), touched.end(), 0);
        reinsertions.resize(count);
#pragma omp parallel num_threads(threads) if (threads > 1 && count > 256)
        {
            PairStack stack;
            stack.init(bvh.max_depth);
#pragma omp for schedule(dynamic, 64)
            for (long i = 0; i < (long)count; i++) reinsertions[i] = find_reinsertion(bvh, candidates[i].node_id, stack);
        }
        // descending by area_diff, ties by candidate rank (H1)
        std::stable_sort(reinsertions.begin(), reinsertions.end(),
                         [](const Reinsertion& a, const Reinsertion& b) { return a.area_diff > b.area_diff; });
        for (size_t i = 0; i < reinsertions.size(); i++) {
            const Reinsertion r = reinsertions[i];
            if (r.area_diff <= 0.0f) continue;
            size_t conflicts[5] = {r.to, r.from, sibling_id(r.from), bvh.parents[r.to], bvh.parents[r.from]};
            bool any = false;
            for (size_t c : conflicts) any |= touched[c] != 0;
            if (any) continue;
            for (size_t c : conflicts) touched[c] = 1;
            reinsert_node(bvh, r.from, r.to);
            applied++;
        }
    }
    // reinsertion.rs:66-90, 113-118: the given ids, in the given order, are the candidates of every iteration
    void run_with_candidates(OrcBvh2& bvh, const u32* ids, size_t n, u32 iterations) {
        if (bvh.nodes.empty() || is_leaf(bvh.nodes[0])) return;
        if (bvh.parents.empty()) compute_parents(bvh);
        candidates.clear();
        for (size_t i = 0; i < n; i++) candidates.push_back(Candidate{ids[i], half_area(bvh.nodes[ids[i]].aabb)});
        touched.assign(bvh.nodes.size(), 0);
        bvh.children_are_ordered_after_parents = false;
        for (u32 k = 0; k < iterations; k++) optimize_candidates(bvh, candidates.size());
    }
    // reinsertion.rs:40-57, 92-111
    void run(OrcBvh2& bvh, float ratio, const float* seq, size_t n_seq) {
        if (bvh.nodes.empty() || is_leaf(bvh.nodes[0]) || ratio <= 0.0f) return;
        if (bvh.parents.empty()) compute_parents(bvh);
        touched.assign(bvh.nodes.size(), 0);
        batch_size_ratio = ratio;
        bvh.children_are_ordered_after_parents = false;
        std::vector<float> default_seq;
        if (!seq) {
            for (int n = 1; n < 32; n += 2) default_seq.push_back(1.0f / (float)n);
            seq = default_seq.data();
            n_seq = default_seq.size();
        }
        for (size_t k = 0; k < n_seq; k++) {
            float r = seq[k];
            float f = ((float)bvh.nodes.size() * batch_size_ratio) * r;
            size_t batch_size = std::max<size_t>(1, (size_t)sat_u64((double)f));  // `as usize` saturating
            size_t node_count = std::min(bvh.nodes.size(), batch_size + 1);
            find_candidates(bvh, node_count);
            optimize_candidates(bvh, node_count - 1);
        }
    }
};

// ---------------------------------------------------------------------------------------------------------
// BVH2 -> CWBVH (cwbvh/bvh2_to_cwbvh.rs)
// ---------------------------------------------------------------------------------------------------------
enum : u8 { KIND_LEAF = 0, KIND_INTERNAL = 1, KIND_DISTRIBUTE = 2 };
struct Decision {  // bvh2_to_cwbvh.rs:478-484 (Default: cost 0, kind DISTRIBUTE, 0, 0)
    float cost = 0.f;
    u8 kind = KIND_DISTRIBUTE;
    u8 distribute_left = 0;
    u8 distribute_right = 0;
};
static const u8 INVALID8 = 0xff;
static const u32 INVALID32 = 0xffffffffu;
static const float PRIM_COST = 0.3f;       // bvh2_to_cwbvh.rs:30
static const float DENOM = 1.0f / 255.0f;  // cwbvh/mod.rs:36-38

struct Bvh2Converter {
    const OrcBvh2& bvh2;
    std::vector<OrcCwBvhNode> nodes;
    std::vector<u32> primitive_indices;
    std::vector<Decision> decisions;
    std::vector<OrcAabb> exact;
    bool order_children_during_build, include_exact;
    V3 direction_lut[8];

    Bvh2Converter(const OrcBvh2& b, bool order, bool ex) : bvh2(b), order_children_during_build(order), include_exact(ex) {
        nodes.reserve(b.primitive_indices.size());
        nodes.push_back(OrcCwBvhNode{});
        primitive_indices.reserve(b.primitive_indices.size());
        decisions.assign(b.nodes.size() * 7, Decision{});
        for (int s = 0; s < 8; s++)  // bvh2_to_cwbvh.rs:40-50
            direction_lut[s] = V3{(s & 4) ? -1.f : 1.f, (s & 2) ? -1.f : 1.f, (s & 1) ? -1.f : 1.f};
        if (ex) exact.assign(b.nodes.size(), aabb_empty());
    }

    // bvh2_to_cwbvh.rs:220-344
    u32 calculate_cost_impl(size_t node_index, u32 max_prims_per_leaf) {
        const OrcBvh2Node& node = bvh2.nodes[node_index];
        float ha = half_area(node.aabb);
        u32 first_index = node.first_index;
        u32 prim_count = node.prim_count;
        size_t node_dec_idx = node_index * 7;
        size_t first_index_7 = (size_t)first_index * 7;
        size_t next_index_7 = ((size_t)first_index + 1) * 7;
        u32 num_primitives;
        if (prim_count != 0) {
            num_primitives = prim_count;
            if (num_primitives != 1) {
                fprintf(stderr, "oracle: BVH8 builder expects 1 primitive per leaf\n");
                abort();
            }
            float cost_leaf = ha * (float)num_primitives * PRIM_COST;
            for (int i = 0; i < 7; i++) {
                decisions[node_dec_idx + i].kind = KIND_LEAF;
                decisions[node_dec_idx + i].cost = cost_leaf;
            }
        } else {
            num_primitives = calculate_cost_impl(first_index, max_prims_per_leaf) +
                             calculate_cost_impl((size_t)first_index + 1, max_prims_per_leaf);
            {
                float cost_leaf = num_primitives <= max_prims_per_leaf ? (float)num_primitives * ha * PRIM_COST : INFINITY;
                float cost_distribute = INFINITY;
                u8 dl = INVALID8, dr = INVALID8;
                for (int k = 0; k < 7; k++) {
                    float c = decisions[first_index_7 + k].cost + decisions[next_index_7 + 6 - k].cost;
                    if (c < cost_distribute) {
                        cost_distribute = c;
                        dl = (u8)k;
                        dr = (u8)(6 - k);
                    }
                }
                float cost_internal = cost_distribute + ha;
                Decision& d = decisions[node_dec_idx];
                if (cost_leaf < cost_internal) {
                    d.kind = KIND_LEAF;
                    d.cost = cost_leaf;
                } else {
                    d.kind = KIND_INTERNAL;
                    d.cost = cost_internal;
                }
                d.distribute_left = dl;
                d.distribute_right = dr;
            }
            for (int i = 1; i < 7; i++) {
                size_t node_i = node_dec_idx + i;
                float cost_distribute = decisions[node_i - 1].cost;
                u8 dl = INVALID8, dr = INVALID8;
                for (int k = 0; k < i; k++) {
                    float c = decisions[first_index_7 + k].cost + decisions[next_index_7 + i - k - 1].cost;
                    if (c < cost_distribute) {
                        cost_distribute = c;
                        dl = (u8)k;
                        dr = (u8)(i - k - 1);
                    }
                }
                Decision& d = decisions[node_i];
                d.cost = cost_distribute;
                if (dl != INVALID8) {
                    d.kind = KIND_DISTRIBUTE;
                    d.distribute_left = dl;
                    d.distribute_right = dr;
                } else {
                    decisions[node_i] = decisions[node_i - 1];
                }
            }
        }
        return num_primitives;
    }

    // bvh2_to_cwbvh.rs:346-397
    void get_children(size_t node_index, u32 children[8], u32& child_count, size_t i) {
        const OrcBvh2Node& node = bvh2.nodes[node_index];
        if (is_leaf(node)) {
            children[child_count++] = (u32)node_index;
            return;
        }
        const Decision& d = decisions[node_index * 7 + i];
        u8 dl = d.distribute_left, dr = d.distribute_right;
        if (decisions[(size_t)node.first_index * 7 + dl].kind == KIND_DISTRIBUTE)
            get_children(node.first_index, children, child_count, dl);
        else
            children[child_count++] = node.first_index;
        if (decisions[((size_t)node.first_index + 1) * 7 + dr].kind == KIND_DISTRIBUTE)
            get_children((size_t)node.first_index + 1, children, child_count, dr);
        else
            children[child_count++] = node.first_index + 1;
    }

    // bvh2_to_cwbvh.rs:402-467
    void order_children(size_t node_index, u32 children[8], size_t child_count) {
        const OrcAabb& na = bvh2.nodes[node_index].aabb;
        V3 p = (v3(na.max) + v3(na.min)) * 0.5f;
        float cost[8][8];
        for (int c = 0; c < 8; c++)
            for (int s = 0; s < 8; s++) cost[c][s] = 3.40282347e+38f;
        for (int s = 0; s < 8; s++) {
            V3 d = direction_lut[s];
            for (size_t c = 0; c < child_count; c++) {
                const OrcAabb& ca = bvh2.nodes[children[c]].aabb;
                V3 v = (v3(ca.max) + v3(ca.min)) * 0.5f - p;
                cost[c][s] = dot(d, v);
            }
        }
        size_t assignment[8];
        bool slot_filled[8];
        for (int i = 0; i < 8; i++) {
            assignment[i] = (size_t)INVALID32;
            slot_filled[i] = false;
        }
        for (;;) {
            float min_cost = 3.40282347e+38f;
            size_t min_slot = (size_t)INVALID32, min_index = (size_t)INVALID32;
            for (size_t c = 0; c < child_count; c++) {
                if (assignment[c] == (size_t)INVALID32) {
                    for (size_t s = 0; s < 8; s++) {
                        float cs = cost[c][s];
                        if (!slot_filled[s] && cs < min_cost) {
                            min_cost = cs;
                            min_slot = s;
                            min_index = c;
                        }
                    }
                }
            }
            if (min_slot == (size_t)INVALID32) break;
            slot_filled[min_slot] = true;
            assignment[min_index] = min_slot;
        }
        u32 original[8];
        for (int i = 0; i < 8; i++) {
            original[i] = children[i];
            children[i] = INVALID32;
        }
        for (size_t i = 0; i < child_count; i++) {
            if (assignment[i] == (size_t)INVALID32) {
                fprintf(stderr, "oracle: order_children left a child unassigned (NaN/inf centre) -- reference panics\n");
                abort();
            }
            children[assignment[i]] = original[i];
        }
    }

    // bvh2_to_cwbvh.rs:197-211
    u32 count_primitives(size_t node_index) {
        const OrcBvh2Node& node = bvh2.nodes[node_index];
        if (is_leaf(node)) {
            primitive_indices.push_back(bvh2.primitive_indices[node.first_index]);
            return node.prim_count;
        }
        return count_primitives(node.first_index) + count_primitives((size_t)node.first_index + 1);
    }

    // bvh2_to_cwbvh.rs:75-193
    void convert_impl(size_t node_index_bvh8, size_t node_index_bvh2) {
        OrcCwBvhNode node = nodes[node_index_bvh8];
        const OrcAabb& aabb = bvh2.nodes[node_index_bvh2].aabb;
        if (include_exact) {
            if (exact.size() <= node_index_bvh8) exact.resize(node_index_bvh8 + 1, aabb_empty());
            exact[node_index_bvh8] = aabb;
        }
        float rcp_e[3];
        for (int k = 0; k < 3; k++) {
            node.p[k] = aabb.min[k];
            float ext = smax(aabb.max[k] - aabb.min[k], 1e-20f);  // Vec3A::max(splat(1e-20))
            float v = exp2f(ceilf(log2f(ext * DENOM)));
            rcp_e[k] = 1.0f / v;
            node.e[k] = (u8)(f2u(v) >> 23);
        }
        u32 children[8];
        for (int i = 0; i < 8; i++) children[i] = INVALID32;
        u32 child_count = 0;
        get_children(node_index_bvh2, children, child_count, 0);
        if (order_children_during_build) order_children(node_index_bvh2, children, child_count);
        node.imask = 0;
        node.primitive_base_idx = (u32)primitive_indices.size();
        node.child_base_idx = (u32)nodes.size();
        size_t num_internal_nodes = 0;
        u32 num_primitives = 0;
        for (int i = 0; i < 8; i++) {
            u32 child_index = children[i];
            if (child_index == INVALID32) continue;
            const OrcAabb& ca = bvh2.nodes[child_index].aabb;
            float cmin[3], cmax[3];
            for (int k = 0; k < 3; k++) {
                float lo = floorf(((ca.min[k] - node.p[k]) - 0.0f) * rcp_e[k]);
                float hi = ceilf(((ca.max[k] - node.p[k]) + 0.0f) * rcp_e[k]);
                // glam clamp = self.max(min).min(max) with SSE operand semantics
                lo = smin(smax(lo, 0.0f), 255.0f);
                hi = smin(smax(hi, 0.0f), 255.0f);
                cmin[k] = lo;
                cmax[k] = hi;
            }
            node.child_min_x[i] = sat_u8(cmin[0]);
            node.child_min_y[i] = sat_u8(cmin[1]);
            node.child_min_z[i] = sat_u8(cmin[2]);
            node.child_max_x[i] = sat_u8(cmax[0]);
            node.child_max_y[i] = sat_u8(cmax[1]);
            node.child_max_z[i] = sat_u8(cmax[2]);
            u8 kind = decisions[(size_t)child_index * 7].kind;
            if (kind == KIND_LEAF) {
                u32 pc = count_primitives(child_index);
                u8 unary = pc == 1 ? 0x20 : pc == 2 ? 0x60 : pc == 3 ? 0xe0 : 0;
                if (!unary) {
                    fprintf(stderr, "oracle: incorrect leaf primitive count %u\n", pc);
                    abort();
                }
                node.child_meta[i] = (u8)num_primitives | unary;
                num_primitives += pc;
            } else if (kind == KIND_INTERNAL) {
                node.imask |= (u8)(1u << i);
                node.child_meta[i] = (u8)((24 + i) | 0x20);
                num_internal_nodes++;
            } else {
                fprintf(stderr, "oracle: unreachable DISTRIBUTE child\n");
                abort();
            }
        }
        nodes.resize(nodes.size() + num_internal_nodes, OrcCwBvhNode{});
        nodes[node_index_bvh8] = node;
        u32 offset = 0;
        for (int i = 0; i < 8; i++) {
            if (children[i] != INVALID32 && (node.imask & (1u << i))) {
                convert_impl(node.child_base_idx + offset, children[i]);
                offset++;
            }
        }
    }
};

static OrcCwBvh* bvh2_to_cwbvh(const OrcBvh2& bvh2, u32 max_prims_per_leaf, bool order_children, bool include_exact) {
    OrcCwBvh* out = new OrcCwBvh();
    out->total_aabb = OrcAabb{};
    out->uses_spatial_splits = bvh2.uses_spatial_splits;  // bvh2_to_cwbvh.rs:508
    if (bvh2.nodes.empty()) return out;  // CwBvh::default()
    Bvh2Converter conv(bvh2, order_children, include_exact);
    conv.calculate_cost_impl(0, max_prims_per_leaf);
    conv.convert_impl(0, 0);
    out->nodes.swap(conv.nodes);
    out->primitive_indices.swap(conv.primitive_indices);
    out->total_aabb = bvh2.nodes[0].aabb;
    out->total_aabb._p0 = out->total_aabb._p1 = 0.f;
    out->exact_node_aabbs.swap(conv.exact);
    return out;
}

// ---------------------------------------------------------------------------------------------------------
// CWBVH traversal (cwbvh/node.rs, simd.rs, traverse_macro.rs, mod.rs)
// ---------------------------------------------------------------------------------------------------------
static const float NODE_EPSILON = 0.0001f;  // cwbvh/node.rs:82

// cwbvh/node.rs:207-231
static inline void get_child_and_index_bits(const OrcCwBvhNode& n, u32 oct_inv4, u64& child_bits8, u64& bit_index8) {
    u64 oct_inv8 = oct_inv4;
    oct_inv8 |= oct_inv8 << 32;
    u64 meta8;
    memcpy(&meta8, n.child_meta, 8);  // from_le_bytes (little-endian host)
    const u64 inner_mask = 0x1010101010101010ull;
    u64 is_inner8 = (meta8 & (meta8 << 1)) & inner_mask;
    u64 inner_mask8 = (is_inner8 >> 4) * 0xffull;
    const u64 index_mask = 0x1f1f1f1f1f1f1f1full;
    bit_index8 = (meta8 ^ (oct_inv8 & inner_mask8)) & index_mask;
    const u64 child_mask = 0x0707070707070707ull;
    child_bits8 = (meta8 >> 5) & child_mask;
}
static inline u32 extract_byte64(u64 x, int b) { return (u32)(x >> (b * 8)) & 0xffu; }

// cwbvh/node.rs:106-154 intersect_ray_basic
static inline u32 node_intersect_basic(const OrcCwBvhNode& n, const OrcRay& ray, u32 oct_inv4) {
    float ext[3], adj_dir[3], adj_org[3];
    for (int k = 0; k < 3; k++) {
        ext[k] = u2f((u32)n.e[k] << 23);  // node.rs:269-275
        adj_dir[k] = ext[k] * ray.inv_direction[k];
        adj_org[k] = (n.p[k] - ray.origin[k]) * ray.inv_direction[k];
    }
    bool rdx = ray.direction[0] < 0.0f, rdy = ray.direction[1] < 0.0f, rdz = ray.direction[2] < 0.0f;
    u64 child_bits8, bit_index8;
    get_child_and_index_bits(n, oct_inv4, child_bits8, bit_index8);
    u32 hit_mask = 0;
    for (int c = 0; c < 8; c++) {
        float x_min = rdx ? n.child_max_x[c] : n.child_min_x[c], x_max = rdx ? n.child_min_x[c] : n.child_max_x[c];
        float y_min = rdy ? n.child_max_y[c] : n.child_min_y[c], y_max = rdy ? n.child_min_y[c] : n.child_max_y[c];
        float z_min = rdz ? n.child_max_z[c] : n.child_min_z[c], z_max = rdz ? n.child_min_z[c] : n.child_max_z[c];
        float tminx = x_min * adj_dir[0] + adj_org[0], tmaxx = x_max * adj_dir[0] + adj_org[0];
        float tminy = y_min * adj_dir[1] + adj_org[1], tmaxy = y_max * adj_dir[1] + adj_org[1];
        float tminz = z_min * adj_dir[2] + adj_org[2], tmaxz = z_max * adj_dir[2] + adj_org[2];
        // simd.rs:81-84 nesting (identical to node.rs for finite inputs)
        float tmin = smax(tminx, smax(tminy, tminz));
        float tmax = smin(tmaxx, smin(tmaxy, tmaxz));
        tmin = smax(tmin, NODE_EPSILON);
        tmax = smin(tmax, ray.tmax);
        if (tmin <= tmax) hit_mask |= extract_byte64(child_bits8, c) << extract_byte64(bit_index8, c);
    }
    return hit_mask;
}

#if defined(__SSE2__)
// cwbvh/simd.rs:17-100
static inline u32 node_intersect_simd(const OrcCwBvhNode& n, const OrcRay& ray, u32 oct_inv4) {
    float adj_dir[3], adj_org[3];
    for (int k = 0; k < 3; k++) {
        float ext = u2f((u32)n.e[k] << 23);
        adj_dir[k] = ext * ray.inv_direction[k];
        adj_org[k] = (n.p[k] - ray.origin[k]) * ray.inv_direction[k];
    }
    __m128 dx = _mm_set1_ps(adj_dir[0]), dy = _mm_set1_ps(adj_dir[1]), dz = _mm_set1_ps(adj_dir[2]);
    __m128 ox = _mm_set1_ps(adj_org[0]), oy = _mm_set1_ps(adj_org[1]), oz = _mm_set1_ps(adj_org[2]);
    bool rdx = ray.direction[0] < 0.0f, rdy = ray.direction[1] < 0.0f, rdz = ray.direction[2] < 0.0f;
    u64 child_bits8, bit_index8;
    get_child_and_index_bits(n, oct_inv4, child_bits8, bit_index8);
    auto get_q = [](const u8* v, int i) {
        return _mm_set_ps((float)v[i * 4 + 3], (float)v[i * 4 + 2], (float)v[i * 4 + 1], (float)v[i * 4]);
    };
    u32 hit_mask = 0;
    for (int i = 0; i < 2; i++) {
        __m128 qlx = get_q(n.child_min_x, i), qhx = get_q(n.child_max_x, i);
        __m128 x_min = rdx ? qhx : qlx, x_max = rdx ? qlx : qhx;
        __m128 tmin_x = _mm_add_ps(_mm_mul_ps(x_min, dx), ox), tmax_x = _mm_add_ps(_mm_mul_ps(x_max, dx), ox);
        __m128 qly = get_q(n.child_min_y, i), qhy = get_q(n.child_max_y, i);
        __m128 y_min = rdy ? qhy : qly, y_max = rdy ? qly : qhy;
        __m128 tmin_y = _mm_add_ps(_mm_mul_ps(y_min, dy), oy), tmax_y = _mm_add_ps(_mm_mul_ps(y_max, dy), oy);
        __m128 qlz = get_q(n.child_min_z, i), qhz = get_q(n.child_max_z, i);
        __m128 z_min = rdz ? qhz : qlz, z_max = rdz ? qlz : qhz;
        __m128 tmin_z = _mm_add_ps(_mm_mul_ps(z_min, dz), oz), tmax_z = _mm_add_ps(_mm_mul_ps(z_max, dz), oz);
        __m128 tmin = _mm_max_ps(tmin_x, _mm_max_ps(tmin_y, tmin_z));
        __m128 tmax = _mm_min_ps(tmax_x, _mm_min_ps(tmax_y, tmax_z));
        tmin = _mm_max_ps(tmin, _mm_set1_ps(NODE_EPSILON));
        tmax = _mm_min_ps(tmax, _mm_set1_ps(ray.tmax));
        int mask = _mm_movemask_ps(_mm_cmple_ps(tmin, tmax));
        for (int j = 0; j < 4; j++) {
            int offset = i * 4 + j;
            if (mask & (1 << j)) hit_mask |= extract_byte64(child_bits8, offset) << extract_byte64(bit_index8, offset);
        }
    }
    return hit_mask;
}
#else
static inline u32 node_intersect_simd(const OrcCwBvhNode& n, const OrcRay& ray, u32 oct_inv4) {
    return node_intersect_basic(n, ray, oct_inv4);
}
#endif

static inline u32 firstbithigh(u32 v) { return 31 - (u32)__builtin_clz(v); }  // cwbvh/mod.rs:996-998
// cwbvh/mod.rs:1001-1010
static inline u32 ray_get_octant_inv4(const float* dir) {
    return (dir[0] < 0.0f ? 0 : 0x04040404u) | (dir[1] < 0.0f ? 0 : 0x02020202u) | (dir[2] < 0.0f ? 0 : 0x01010101u);
}

// traverse_macro.rs:59-126 with the three primitive blocks of cwbvh/mod.rs:169-245.
// MODE 0 = closest hit, 1 = miss (first hit terminates), 2 = all hits (count).
template <int MODE, bool SIMD>
static inline void traverse_one(const OrcCwBvh& bvh, const OrcTriangle* tris, const OrcRay& ray_in, OrcRayHit* hit, u8* miss,
                                u32* anycount, u64& nodes_visited, u64& tris_tested) {
    OrcRay ray = ray_in;  // traverse_ray
    struct G {
        u32 x, y;
    };
    G stack[32];  // StackStack<UVec2, 32>, cwbvh/mod.rs:60
    size_t sp = 0;
    G current_group = bvh.nodes.empty() ? G{0, 0} : G{0, 0x80000000u};  // cwbvh/mod.rs:146-165
    G primitive_group = G{0, 0};
    u32 oct_inv4 = ray_get_octant_inv4(ray.direction);
    bool is_miss = true;
    u32 count = 0;
    for (;;) {
        while (primitive_group.y != 0) {
            u32 local = firstbithigh(primitive_group.y);
            primitive_group.y &= ~(1u << local);
            u32 primitive_id = primitive_group.x + local;
            float t = tri_intersect(tris[primitive_id], ray);
            tris_tested++;
            if (MODE == 0) {
                if (t < ray.tmax) {
                    hit->primitive_id = primitive_id;
                    hit->t = t;
                    ray.tmax = t;
                }
            } else if (MODE == 1) {
                if (t < ray.tmax) {
                    is_miss = false;
                    goto done;
                }
            } else {
                if (t < INFINITY) count++;
            }
        }
        primitive_group = G{0, 0};
        if (current_group.y & 0xff000000u) {
            u32 hits_imask = current_group.y;
            u32 child_index_offset = firstbithigh(hits_imask);
            u32 child_index_base = current_group.x;
            current_group.y &= ~(1u << child_index_offset);
            if (current_group.y & 0xff000000u) {  // push, saturating (faststack.rs:299-303)
                stack[sp] = current_group;
                sp = std::min<size_t>(sp + 1, 31);
            }
            u32 slot_index = (child_index_offset - 24) ^ (oct_inv4 & 0xff);
            u32 relative_index = (u32)__builtin_popcount(hits_imask & ~(0xffffffffu << slot_index));
            u32 child_node_index = child_index_base + relative_index;
            const OrcCwBvhNode& node = bvh.nodes[child_node_index];
            nodes_visited++;
            u32 hitmask = SIMD ? node_intersect_simd(node, ray, oct_inv4) : node_intersect_basic(node, ray, oct_inv4);
            current_group.x = node.child_base_idx;
            primitive_group.x = node.primitive_base_idx;
            current_group.y = (hitmask & 0xff000000u) | (u32)node.imask;
            primitive_group.y = hitmask & 0x00ffffffu;
        } else {
            current_group = G{0, 0};
        }
        if (primitive_group.y == 0 && (current_group.y & 0xff000000u) == 0) {
            if (sp == 0) break;
            sp = sp - 1;  // pop_fast
            current_group = stack[sp];
        }
    }
done:
    if (MODE == 1) *miss = is_miss ? 1 : 0;
    if (MODE == 2) *anycount = count;
}

template <int MODE>
static void traverse_batch(const OrcCwBvh& bvh, const OrcTriangle* tris, const OrcRay* rays, size_t n, OrcRayHit* hits, u8* miss,
                           u32* counts, int threads, int use_simd, u64* counters) {
    threads = clamp_threads(threads);
    u64 nv = 0, tt = 0;
#pragma omp parallel for num_threads(threads) schedule(dynamic, 256) reduction(+ : nv, tt)
    for (long i = 0; i < (long)n; i++) {
        OrcRayHit h{0xffffffffu, 0xffffffffu, 0xffffffffu, INFINITY};  // RayHit::none(), ray.rs:74-83
        u8 m = 1;
        u32 c = 0;
        u64 a = 0, b = 0;
        if (use_simd)
            traverse_one<MODE, true>(bvh, tris, rays[i], &h, &m, &c, a, b);
        else
            traverse_one<MODE, false>(bvh, tris, rays[i], &h, &m, &c, a, b);
        if (MODE == 0) hits[i] = h;
        if (MODE == 1) miss[i] = m;
        if (MODE == 2) counts[i] = c;
        nv += a;
        tt += b;
    }
    if (counters) {
        counters[0] += nv;
        counters[1] += tt;
    }
}

// ---------------------------------------------------------------------------------------------------------
// SAH leaf collapse (bvh2/leaf_collapser.rs:21-192)
// ---------------------------------------------------------------------------------------------------------
static void collapse(OrcBvh2& bvh, u32 max_prims, float traversal_cost) {
    const size_t nodes_qty = bvh.nodes.size();
    if (max_prims <= 1 || (u32)nodes_qty <= max_prims * 2 + 1) return;                                        // :25-27
    if (!bvh.primitive_indices.empty() && (u32)bvh.primitive_indices.size() <= max_prims) return;            // :29-31
    if (bvh.nodes.empty() || is_leaf(bvh.nodes[0])) return;                                                   // :33-35
    const bool previously_had_parents = !bvh.parents.empty();
    if (bvh.parents.empty()) compute_parents(bvh);
    std::vector<u32> node_counts(nodes_qty, 1), prim_counts(nodes_qty, 0);
    // bottom-up traversal (:48-87 + bottom_up_traverse :195-236): every inner node after both of its children
    auto process = [&](bool leaf, size_t i) {
        if (leaf) {
            prim_counts[i] = bvh.nodes[i].prim_count;
            return;
        }
        const OrcBvh2Node& node = bvh.nodes[i];
        const size_t first_child = node.first_index;
        const u32 left_count = prim_counts[first_child], right_count = prim_counts[first_child + 1];
        const u32 total_count = left_count + right_count;
        if (left_count > 0 && right_count > 0 && total_count <= max_prims) {
            const OrcBvh2Node& left = bvh.nodes[first_child];
            const OrcBvh2Node& right = bvh.nodes[first_child + 1];
            const float collapse_cost = half_area(node.aabb) * ((float)total_count - traversal_cost);
            const float base_cost = half_area(left.aabb) * (float)left_count + half_area(right.aabb) * (float)right_count;
            const bool both_have_same_prim = (left.first_index == right.first_index) && total_count == 2;
            if (collapse_cost <= base_cost || both_have_same_prim) {
                prim_counts[i] = total_count;
                prim_counts[first_child] = 0;
                prim_counts[first_child + 1] = 0;
                node_counts[first_child] = 0;
                node_counts[first_child + 1] = 0;
            }
        }
    };
    {
        std::vector<u8> flags(nodes_qty, 0);
        for (size_t i = 1; i < nodes_qty; i++) {
            if (!is_leaf(bvh.nodes[i])) continue;
            process(true, i);
            size_t j = i;
            while (j != 0) {
                j = bvh.parents[j];
                u8 prev = flags[j];
                flags[j] = (u8)std::min<int>(prev + 1, 255);
                if (prev != 1) break;
                flags[j] = 0;
                process(false, j);
            }
        }
    }
    // inclusive prefix sums (:89-98)
    for (size_t i = 1; i < nodes_qty; i++) {
        node_counts[i] += node_counts[i - 1];
        prim_counts[i] += prim_counts[i - 1];
    }
    std::vector<u32> indices_copy;
    std::vector<OrcBvh2Node> nodes_copy;
    const u32 node_count = node_counts[nodes_qty - 1];
    const bool root_became_leaf = prim_counts[0] > 0;
    if (root_became_leaf) {  // :104-110 (unreachable for trees whose leaves all hold >= 1 primitive; kept for fidelity)
        bvh.nodes[0].first_index = 0;
        bvh.nodes[0].prim_count = prim_counts[0];
        std::swap(bvh.primitive_indices, indices_copy);
        std::swap(bvh.nodes, nodes_copy);
    } else {
        nodes_copy.assign(node_count, OrcBvh2Node{});
        indices_copy.assign(prim_counts[nodes_qty - 1], 0);
        nodes_copy[0] = bvh.nodes[0];
        nodes_copy[0].first_index = node_counts[nodes_copy[0].first_index - 1];
    }
    auto top_down = [&](size_t i) {  // :122-151: the primitives of the subtree of i, depth first, left to right
        u32 first_prim = prim_counts[i - 1];
        size_t j = i;
        for (;;) {
            const OrcBvh2Node node = bvh.nodes[j];
            if (is_leaf(node)) {
                for (u32 n = 0; n < node.prim_count; n++) indices_copy[first_prim + n] = bvh.primitive_indices[node.first_index + n];
                first_prim += node.prim_count;
                while (!(j % 2 == 1) && j != i) j = bvh.parents[j];  // !is_left_sibling(j)
                if (j == i) break;
                j = sibling_id(j);
            } else {
                j = node.first_index;
            }
        }
    };
    for (size_t i = 1; i < bvh.nodes.size(); i++) {  // :153-174 (bvh.nodes is empty here when the root became a leaf)
        const size_t node_id = node_counts[i - 1];
        if (node_id == node_counts[i]) continue;
        nodes_copy[node_id] = bvh.nodes[i];
        const u32 first_prim = prim_counts[i - 1];
        if (first_prim == prim_counts[i]) {
            nodes_copy[node_id].first_index = node_counts[nodes_copy[node_id].first_index - 1];
        } else {
            nodes_copy[node_id].prim_count = prim_counts[i] - first_prim;
            nodes_copy[node_id].first_index = first_prim;
            top_down(i);
        }
    }
    std::swap(bvh.nodes, nodes_copy);
    std::swap(bvh.primitive_indices, indices_copy);
    if (previously_had_parents) compute_parents(bvh);  // update_parents (:183-186)
    else bvh.parents.clear();
}

// ---------------------------------------------------------------------------------------------------------
// Bvh2 ray traversal (bvh2/mod.rs:148-334, aabb.rs:186-206)
// ---------------------------------------------------------------------------------------------------------
static inline float aabb_intersect_ray(const OrcAabb& a, const OrcRay& ray) {
    float t1[3], t2[3], tmn[3], tmx[3];
    for (int k = 0; k < 3; k++) {
        t1[k] = (a.min[k] - ray.origin[k]) * ray.inv_direction[k];
        t2[k] = (a.max[k] - ray.origin[k]) * ray.inv_direction[k];
        tmn[k] = smin(t1[k], t2[k]);
        tmx[k] = smax(t1[k], t2[k]);
    }
    // glam sse2 Vec3A::max_element / min_element: combine (x,z) and (y,z), then the two results
    float tmin_n = smax(smax(tmn[0], tmn[2]), smax(tmn[1], tmn[2]));
    float tmax_n = smin(smin(tmx[0], tmx[2]), smin(tmx[1], tmx[2]));
    if (tmax_n >= tmin_n && tmax_n >= 0.0f) return tmin_n;
    return INFINITY;
}

// MODE 0 closest hit (ray_traverse :148-172), 1 miss (:185-213), 2 counting any-hit (:225-238)
template <int MODE>
static inline void bvh2_traverse_one(const OrcBvh2& bvh, const OrcTriangle* tris, const OrcRay& ray_in, OrcRayHit* hit, u8* miss,
                                     u32* anycount, u64& nodes_tested, u64& tris_tested) {
    OrcRay ray = ray_in;
    bool is_miss = true;
    u32 count = 0;
    // the leaf callback; returns false to halt the traversal
    auto intersect_prims = [&](const OrcBvh2Node& node) -> bool {
        for (u32 primitive_id = node.first_index; primitive_id < node.first_index + node.prim_count; primitive_id++) {
            float t = tri_intersect(tris[primitive_id], ray);
            tris_tested++;
            if (MODE == 0) {
                if (t < ray.tmax) {
                    hit->primitive_id = primitive_id;
                    hit->t = t;
                    ray.tmax = t;
                }
            } else if (MODE == 1) {
                if (t < ray.tmax) {
                    is_miss = false;
                    return false;
                }
            } else {
                if (t < INFINITY) count++;
            }
        }
        return true;
    };
    auto finish = [&]() {
        if (MODE == 1) *miss = is_miss ? 1 : 0;
        if (MODE == 2) *anycount = count;
    };
    // ray_traverse_dynamic :260-334
    if (bvh.nodes.empty()) return finish();
    const OrcBvh2Node& root = bvh.nodes[0];
    nodes_tested++;
    if (!(aabb_intersect_ray(root.aabb, ray) < ray.tmax)) return finish();
    if (is_leaf(root)) {
        intersect_prims(root);
        return finish();
    }
    std::vector<u32> heap_stack;  // fast_stack!(u32, (96, 192), max_depth): saturating fixed stacks, heap beyond 192
    const size_t cap = bvh.max_depth <= 96 ? 96 : (bvh.max_depth <= 192 ? 192 : 0);
    u32 fixed[192];
    size_t sp = 0;
    u32 current = root.first_index;
    for (;;) {
        const OrcBvh2Node* left = &bvh.nodes[current];
        const OrcBvh2Node* right = &bvh.nodes[(size_t)current + 1];
        float left_t = aabb_intersect_ray(left->aabb, ray), right_t = aabb_intersect_ray(right->aabb, ray);
        nodes_tested += 2;
        if (left_t > right_t) {
            std::swap(left_t, right_t);
            std::swap(left, right);
        }
        const bool hit_left = left_t < ray.tmax;
        bool go_left = hit_left;
        if (hit_left && is_leaf(*left)) {
            if (!intersect_prims(*left)) return finish();
            go_left = false;
        }
        const bool hit_right = right_t < ray.tmax;
        bool go_right = hit_right;
        if (hit_right && is_leaf(*right)) {
            if (!intersect_prims(*right)) return finish();
            go_right = false;
        }
        if (go_left && go_right) {
            current = left->first_index;
            if (cap) {
                fixed[sp] = right->first_index;
                sp = std::min(sp + 1, cap - 1);
            } else {
                heap_stack.push_back(right->first_index);
            }
        } else if (go_left) {
            current = left->first_index;
        } else if (go_right) {
            current = right->first_index;
        } else {
            if (cap ? sp == 0 : heap_stack.empty()) {
                hit->t = ray.tmax;  // :326
                return finish();
            }
            if (cap) current = fixed[--sp];
            else {
                current = heap_stack.back();
                heap_stack.pop_back();
            }
        }
    }
}

template <int MODE>
static void bvh2_traverse_batch(const OrcBvh2& bvh, const OrcTriangle* tris, const OrcRay* rays, size_t n, OrcRayHit* hits, u8* miss,
                                u32* counts, int threads, u64* counters) {
    threads = clamp_threads(threads);
    u64 nv = 0, tt = 0;
#pragma omp parallel for num_threads(threads) schedule(dynamic, 256) reduction(+ : nv, tt)
    for (long i = 0; i < (long)n; i++) {
        OrcRayHit h{0xffffffffu, 0xffffffffu, 0xffffffffu, INFINITY};
        u8 m = 1;
        u32 c = 0;
        u64 a = 0, b = 0;
        bvh2_traverse_one<MODE>(bvh, tris, rays[i], &h, &m, &c, a, b);
        if (MODE == 0) hits[i] = h;
        if (MODE == 1) miss[i] = m;
        if (MODE == 2) counts[i] = c;
        nv += a;
        tt += b;
    }
    if (counters) {
        counters[0] += nv;
        counters[1] += tt;
    }
}

// ---------------------------------------------------------------------------------------------------------
// validation (bvh2/mod.rs:786-981, cwbvh/mod.rs:747-908) -- invariants restated, not the stats
// ---------------------------------------------------------------------------------------------------------
// ---------------------------------------------------------------------------------------------------------
// Broad-phase queries: Bvh2::aabb_traverse / point_traverse (bvh2/mod.rs:365-456) and the traverse! macro over
// CwBvhNode::intersect_aabb / contains_point (cwbvh/node.rs:157-200), with an `eval` that always returns true.
// ---------------------------------------------------------------------------------------------------------
// aabb.rs:181-183: (self.min.cmpgt(other.max) | self.max.cmplt(other.min)).bitmask() == 0   (three lanes)
static inline bool aabb_intersect_aabb(const float* smin_, const float* smax_, const float* omin, const float* omax) {
    for (int k = 0; k < 3; k++)
        if (smin_[k] > omax[k] || smax_[k] < omin[k]) return false;
    return true;
}
// aabb.rs:70-72: (point.cmpge(self.min) & point.cmple(self.max)).all()
static inline bool aabb_contains_point(const float* mn, const float* mx, const float* p) {
    for (int k = 0; k < 3; k++)
        if (!(p[k] >= mn[k]) || !(p[k] <= mx[k])) return false;
    return true;
}
// QUERY 0: q = Aabb (min at q[0..3], max at q[4..7]); QUERY 1: q = point (Vec3A)
template <int QUERY>
static inline bool bvh2_node_test(const OrcBvh2Node& n, const float* q) {
    return QUERY == 0 ? aabb_intersect_aabb(n.aabb.min, n.aabb.max, q, q + 4) : aabb_contains_point(n.aabb.min, n.aabb.max, q);
}
template <int QUERY, class Emit>
static void bvh2_query_one(const OrcBvh2& bvh, const float* q, Emit emit) {
    if (bvh.nodes.empty()) return;
    const OrcBvh2Node& root = bvh.nodes[0];
    if (is_leaf(root)) {  // :370-376
        if (bvh2_node_test<QUERY>(root, q)) emit(0u);
        return;
    }
    const size_t cap = bvh.max_depth <= 96 ? 96 : (bvh.max_depth <= 192 ? 192 : bvh.max_depth);  // fast_stack!(u32, (96, 192), max_depth)
    std::vector<u32> stack(cap);
    size_t sp = 0;
    auto push = [&](u32 v) {  // StackStack::push saturates at the last slot (faststack.rs:299-303)
        stack[sp] = v;
        sp = std::min(sp + 1, cap - 1);
    };
    push(root.first_index);
    while (sp > 0) {
        u32 node_index = stack[--sp];
        for (u32 k = 0; k < 2; k++) {  // left, then right
            const OrcBvh2Node& node = bvh.nodes[node_index + k];
            if (bvh2_node_test<QUERY>(node, q)) {
                if (is_leaf(node)) emit(node_index + k);
                else push(node.first_index);
            }
        }
    }
}
// cwbvh/node.rs:157-200
template <int QUERY>
static inline u32 cw_node_query(const OrcCwBvhNode& n, const float* q, u32 oct_inv4) {
    float rcp[3], lo[3], hi[3];
    for (int k = 0; k < 3; k++) {
        float e = u2f((u32)n.e[k] << 23);
        rcp[k] = 1.0f / e;
        lo[k] = (q[k] - n.p[k]) * rcp[k];
        if (QUERY == 0) hi[k] = (q[4 + k] - n.p[k]) * rcp[k];
    }
    u64 child_bits8, bit_index8;
    get_child_and_index_bits(n, oct_inv4, child_bits8, bit_index8);
    u32 hit_mask = 0;
    for (int child = 0; child < 8; child++) {
        float cmin[3] = {(float)n.child_min_x[child], (float)n.child_min_y[child], (float)n.child_min_z[child]};
        float cmax[3] = {(float)n.child_max_x[child], (float)n.child_max_y[child], (float)n.child_max_z[child]};
        bool hit = QUERY == 0 ? aabb_intersect_aabb(cmin, cmax, lo, hi) : aabb_contains_point(cmin, cmax, lo);
        if (hit) hit_mask |= extract_byte64(child_bits8, child) << extract_byte64(bit_index8, child);
    }
    return hit_mask;
}
// traverse_macro.rs:59-126 with $node_intersection = node.intersect_aabb / contains_point and a primitive block that
// reports state.primitive_id
template <int QUERY, class Emit>
static void cwbvh_query_one(const OrcCwBvh& bvh, const float* q, const float* dir, Emit emit) {
    struct G {
        u32 x, y;
    };
    G stack[32];
    size_t sp = 0;
    G current_group = bvh.nodes.empty() ? G{0, 0} : G{0, 0x80000000u};
    G primitive_group = G{0, 0};
    u32 oct_inv4 = ray_get_octant_inv4(dir);
    for (;;) {
        while (primitive_group.y != 0) {
            u32 local = firstbithigh(primitive_group.y);
            primitive_group.y &= ~(1u << local);
            emit(primitive_group.x + local);
        }
        primitive_group = G{0, 0};
        if (current_group.y & 0xff000000u) {
            u32 hits_imask = current_group.y;
            u32 child_index_offset = firstbithigh(hits_imask);
            u32 child_index_base = current_group.x;
            current_group.y &= ~(1u << child_index_offset);
            if (current_group.y & 0xff000000u) {
                stack[sp] = current_group;
                sp = std::min<size_t>(sp + 1, 31);
            }
            u32 slot_index = (child_index_offset - 24) ^ (oct_inv4 & 0xff);
            u32 relative_index = (u32)__builtin_popcount(hits_imask & ~(0xffffffffu << slot_index));
            const OrcCwBvhNode& node = bvh.nodes[child_index_base + relative_index];
            u32 hitmask = cw_node_query<QUERY>(node, q, oct_inv4);
            current_group.x = node.child_base_idx;
            primitive_group.x = node.primitive_base_idx;
            current_group.y = (hitmask & 0xff000000u) | (u32)node.imask;
            primitive_group.y = hitmask & 0x00ffffffu;
        } else {
            current_group = G{0, 0};
        }
        if (primitive_group.y == 0 && (current_group.y & 0xff000000u) == 0) {
            if (sp == 0) break;
            current_group = stack[--sp];
        }
    }
}
// batch drivers: counts[i] = reports of query i; ids_out receives them query after query, in call order, up to cap; returns the total
template <class One>
static size_t query_batch(size_t n, u32* counts, u32* ids_out, size_t cap, One one) {
    size_t total = 0;
    for (size_t i = 0; i < n; i++) {
        u32 c = 0;
        one(i, [&](u32 id) {
            if (ids_out && total < cap) ids_out[total] = id;
            total++;
            c++;
        });
        if (counts) counts[i] = c;
    }
    return total;
}

// ---------------------------------------------------------------------------------------------------------
// splits.rs: spatial pre-splits of large triangles
// ---------------------------------------------------------------------------------------------------------
// aabb.rs:124-134
static inline int largest_axis(const OrcAabb& a) {
    float dx = a.max[0] - a.min[0], dy = a.max[1] - a.min[1], dz = a.max[2] - a.min[2];
    if (dx < dy) return dy < dz ? 2 : 1;
    return dx < dz ? 2 : 0;
}
// aabb.rs:97-103: min = self.min.max(other.min), max = self.max.min(other.max)
static inline OrcAabb aabb_intersection(const OrcAabb& a, const OrcAabb& b) {
    OrcAabb r;
    for (int i = 0; i < 3; i++) {
        r.min[i] = smax(a.min[i], b.min[i]);
        r.max[i] = smin(a.max[i], b.max[i]);
    }
    r._p0 = r._p1 = 0.f;
    return r;
}
// splits.rs:129-158 split_triangle(dim, pos, [v0,v1,v2,v0]). Aabb::INVALID (aabb.rs:23-26) = (f32::MAX, f32::MIN);
// Vec3A::mul_add is a fused multiply-add per lane on every glam backend (_mm_fmadd_ps or f32::mul_add).
static inline void split_triangle(int dim, float pos, const OrcTriangle& t, OrcAabb& left, OrcAabb& right) {
    left = aabb_empty();
    right = aabb_empty();
    const float* v[4] = {t.v0, t.v1, t.v2, t.v0};
    for (int i = 0; i < 3; i++) {
        const float* v0 = v[i];
        const float* v1 = v[i + 1];
        float v0d = v0[dim], v1d = v1[dim];
        if (v0d <= pos) aabb_extend(left, v0);
        if (v0d >= pos) aabb_extend(right, v0);
        if ((v0d < pos && pos < v1d) || (v1d < pos && pos < v0d)) {
            float inv_length = 1.0f / (v1d - v0d);
            float tt = (pos - v0d) * inv_length;
            float c[3];
            for (int k = 0; k < 3; k++) c[k] = fmaf(tt, v1[k] - v0[k], v0[k]);
            aabb_extend(left, c);
            aabb_extend(right, c);
        }
    }
}
// splits.rs:49-125
static void split_aabbs_precise(std::vector<OrcAabb>& aabbs, std::vector<u32>& indices, const OrcTriangle* triangles, float area_thresh_low,
                                float area_thresh_high, float split_factor_low, float split_factor_high, u32 max_iterations,
                                u32 split_tests) {
    std::vector<size_t> candidates;
    for (size_t i = 0; i < aabbs.size(); i++)
        if (half_area(aabbs[i]) > area_thresh_low) candidates.push_back(i);
    size_t old_candidates_len = candidates.size();
    for (u32 it = 0; it < max_iterations; it++) {
        const size_t count = candidates.size();  // the range 0..candidates.len() is evaluated once (:71)
        for (size_t ci = 0; ci < count; ci++) {
            OrcAabb aabb = aabbs[candidates[ci]];
            u32 index = indices[candidates[ci]];
            int axis = largest_axis(aabb);
            const OrcTriangle& tri = triangles[index];
            float best_cost = 3.40282347e+38f;
            OrcAabb left = aabb, right = aabb;
            for (u32 i = 1; i < split_tests; i++) {
                float n = (float)i / (float)split_tests;
                float pos = aabb.min[axis] * n + aabb.max[axis] * (1.0f - n);
                OrcAabb tmp_left = aabb, tmp_right = aabb;
                tmp_left.max[axis] = pos;
                tmp_right.min[axis] = pos;
                OrcAabb t_left, t_right;
                split_triangle(axis, pos, tri, t_left, t_right);
                tmp_left = aabb_intersection(t_left, tmp_left);
                tmp_right = aabb_intersection(t_right, tmp_right);
                float area = half_area(tmp_left) + half_area(tmp_right);
                if (area < best_cost) {
                    best_cost = area;
                    left = tmp_left;
                    right = tmp_right;
                }
            }
            float old_cost = half_area(aabb);
            if ((area_thresh_high > old_cost && best_cost * split_factor_high < old_cost) || best_cost * split_factor_low < old_cost) {
                aabbs[candidates[ci]] = left;
                candidates.push_back(aabbs.size());
                aabbs.push_back(right);
                indices.push_back(index);
            }
        }
        if (old_candidates_len == candidates.size()) break;
        size_t w = 0;  // Vec::retain keeps order (:121)
        for (size_t k = 0; k < candidates.size(); k++)
            if (half_area(aabbs[candidates[k]]) > area_thresh_low) candidates[w++] = candidates[k];
        candidates.resize(w);
        old_candidates_len = candidates.size();
    }
}
// splits.rs:16-34
static void split_aabbs_preset(std::vector<OrcAabb>& aabbs, std::vector<u32>& indices, const OrcTriangle* triangles, float avg_half_area,
                               float largest_half_area) {
    float hi_a = avg_half_area * 4.0f, hi_b = avg_half_area * 0.9f + largest_half_area * 0.1f;
    float hi = fmaxf(hi_a, hi_b);  // f32::max
    split_aabbs_precise(aabbs, indices, triangles, avg_half_area * 3.0f, hi, 1.8f, 1.6f, 12, 12);
}
// cwbvh/builder.rs:28-54 == bvh2/builder.rs:25-51: triangle AABBs, sequential f32 sum and f32::max of their half areas
static void presplit_inputs(const OrcTriangle* tris, size_t n, std::vector<OrcAabb>& aabbs, std::vector<u32>& indices, float* avg_out,
                            float* largest_out) {
    float largest_half_area = 0.0f, avg_area = 0.0f;
    aabbs.resize(n);
    indices.resize(n);
    for (size_t i = 0; i < n; i++) {
        aabbs[i] = tri_aabb(tris[i]);
        float h = half_area(aabbs[i]);
        largest_half_area = fmaxf(h, largest_half_area);
        avg_area += h;
        indices[i] = (u32)i;
    }
    avg_area /= (float)n;
    *avg_out = avg_area;
    *largest_out = largest_half_area;
}

#define VFAIL(code, ...)                     \
    do {                                     \
        if (msg) snprintf(msg, 256, __VA_ARGS__); \
        return code;                         \
    } while (0)

static int bvh2_validate(const OrcBvh2& bvh, const OrcAabb* prim_aabbs, size_t n, bool tight_fit, char* msg) {
    if (msg) msg[0] = 0;
    if (n == 0) {
        if (!bvh.nodes.empty()) VFAIL(-1, "empty input but %zu nodes", bvh.nodes.size());
        return 0;
    }
    {  // 2*leaves - 1 nodes; leaves == n unless leaves were collapsed (leaf_collapser.rs)
        size_t leaves = 0;
        for (const OrcBvh2Node& nd : bvh.nodes) leaves += is_leaf(nd) ? 1 : 0;
        if (bvh.nodes.size() != 2 * leaves - 1 || (leaves > n && !bvh.uses_spatial_splits)) VFAIL(-2, "node count %zu != 2*leaves-1 (leaves=%zu, n=%zu)", bvh.nodes.size(), leaves, n);
    }
    // with spatial splits a primitive shows up in several leaves and may extend outside them (bvh2/mod.rs:823-829,928-945)
    const bool splits = bvh.uses_spatial_splits;
    const size_t n_slots = bvh.primitive_indices.size();
    if (!splits && n_slots != n) VFAIL(-3, "primitive_indices len %zu != %zu", n_slots, n);
    std::vector<u8> seen_node(bvh.nodes.size(), 0), seen_prim(n, 0), seen_slot2(n_slots, 0);
    std::vector<std::pair<u32, u32>> stack;  // node, depth
    stack.push_back({0, 0});
    size_t max_depth = 0, visited = 0;
    while (!stack.empty()) {
        auto [id, depth] = stack.back();
        stack.pop_back();
        if (id >= bvh.nodes.size()) VFAIL(-4, "node index %u out of range", id);
        if (seen_node[id]) VFAIL(-5, "node %u reached twice", id);
        seen_node[id] = 1;
        visited++;
        max_depth = std::max<size_t>(max_depth, depth);
        const OrcBvh2Node& nd = bvh.nodes[id];
        if (is_leaf(nd)) {
            for (u32 k = 0; k < nd.prim_count; k++) {
                u32 slot = nd.first_index + k;
                if (slot >= n_slots) VFAIL(-6, "leaf %u prim slot %u out of range", id, slot);
                if (seen_slot2[slot]) VFAIL(-8, "primitive slot %u referenced twice", slot);
                seen_slot2[slot] = 1;
                u32 prim = bvh.primitive_indices[slot];
                if (prim >= n) VFAIL(-7, "primitive id %u out of range", prim);
                if (seen_prim[prim] && !splits) VFAIL(-8, "primitive %u referenced twice", prim);
                seen_prim[prim] = 1;
                if (splits) continue;
                for (int a = 0; a < 3; a++) {
                    if (!(prim_aabbs[prim].min[a] >= nd.aabb.min[a]) || !(prim_aabbs[prim].max[a] <= nd.aabb.max[a]))
                        VFAIL(-9, "primitive %u not inside leaf %u", prim, id);
                }
                if (tight_fit && nd.prim_count == 1 && !aabb_eq(prim_aabbs[prim], nd.aabb)) VFAIL(-10, "leaf %u not tight", id);
            }
        } else {
            u32 f = nd.first_index;
            if (f % 2 != 1) VFAIL(-11, "node %u first child %u is not odd", id, f);
            if ((size_t)f + 1 >= bvh.nodes.size()) VFAIL(-12, "node %u child out of range", id);
            OrcAabb u = aabb_union(bvh.nodes[f].aabb, bvh.nodes[f + 1].aabb);
            for (int a = 0; a < 3; a++)
                if (!(u.min[a] >= nd.aabb.min[a]) || !(u.max[a] <= nd.aabb.max[a])) VFAIL(-13, "children of %u not inside parent", id);
            if (tight_fit && !aabb_eq(u, nd.aabb)) VFAIL(-14, "node %u is not a tight union of its children", id);
            if (!bvh.parents.empty() && (bvh.parents[f] != id || bvh.parents[f + 1] != id)) VFAIL(-15, "parents of children of %u wrong", id);
            if (bvh.children_are_ordered_after_parents && f <= id) VFAIL(-16, "child %u not after parent %u", f, id);
            stack.push_back({f, depth + 1});
            stack.push_back({f + 1, depth + 1});
        }
    }
    if (visited != bvh.nodes.size()) VFAIL(-17, "visited %zu of %zu nodes", visited, bvh.nodes.size());
    for (size_t i = 0; i < n; i++)
        if (!seen_prim[i]) VFAIL(-18, "primitive %zu unreachable", i);
    if (max_depth >= bvh.max_depth) VFAIL(-19, "depth %zu >= max_depth %zu", max_depth, bvh.max_depth);
    return 0;
}

static int cwbvh_validate(const OrcCwBvh& bvh, const OrcAabb* prim_aabbs, size_t n, char* msg) {
    if (msg) msg[0] = 0;
    const bool splits = bvh.uses_spatial_splits;  // cwbvh/mod.rs:752-755,893
    const size_t n_slots = bvh.primitive_indices.size();
    if (!splits && n_slots != n) VFAIL(-1, "primitive_indices len %zu != %zu", n_slots, n);
    if (bvh.nodes.empty()) {
        if (n != 0) VFAIL(-2, "no nodes for %zu primitives", n);
        return 0;
    }
    std::vector<u8> seen_node(bvh.nodes.size(), 0), seen_slot(n_slots, 0), seen_prim(n, 0);
    struct Item {
        u32 node;
        u32 depth;
        OrcAabb bounds;
    };
    std::vector<Item> stack;
    OrcAabb largest;
    for (int a = 0; a < 3; a++) {
        largest.min[a] = -3.40282347e+38f;
        largest.max[a] = 3.40282347e+38f;
    }
    stack.push_back({0, 0, largest});
    size_t visited = 0, max_depth = 0;
    while (!stack.empty()) {
        Item it = stack.back();
        stack.pop_back();
        if (it.node >= bvh.nodes.size()) VFAIL(-3, "node %u out of range", it.node);
        if (seen_node[it.node]) VFAIL(-4, "node %u reached twice", it.node);
        seen_node[it.node] = 1;
        visited++;
        max_depth = std::max<size_t>(max_depth, it.depth);
        const OrcCwBvhNode& nd = bvh.nodes[it.node];
        float e[3];
        for (int a = 0; a < 3; a++) {
            e[a] = u2f((u32)nd.e[a] << 23);
            if (!isfinite(nd.p[a])) VFAIL(-5, "node %u p not finite", it.node);
            if (!(nd.p[a] >= it.bounds.min[a] - 1.0e-5f) || !(nd.p[a] <= it.bounds.max[a] + 1.0e-5f)) VFAIL(-6, "node %u p outside parent bounds", it.node);
        }
        for (int ch = 0; ch < 8; ch++) {
            u8 meta = nd.child_meta[ch];
            if (meta == 0) {
                if (nd.imask & (1u << ch)) VFAIL(-7, "node %u empty slot %d has imask bit", it.node, ch);
                continue;
            }
            OrcAabb q;
            const u8* lo[3] = {nd.child_min_x, nd.child_min_y, nd.child_min_z};
            const u8* hi[3] = {nd.child_max_x, nd.child_max_y, nd.child_max_z};
            for (int a = 0; a < 3; a++) {
                q.min[a] = (float)lo[a][ch] * e[a] + nd.p[a];
                q.max[a] = (float)hi[a][ch] * e[a] + nd.p[a];
            }
            bool inner = (nd.imask & (1u << ch)) != 0;
            if (inner != ((meta & 0x1f) >= 24)) VFAIL(-8, "node %u slot %d imask/meta mismatch", it.node, ch);
            if (inner) {
                u32 slot_index = (meta & 0x1f) - 24;
                if (slot_index != (u32)ch || (meta >> 5) != 1) VFAIL(-9, "node %u slot %d bad inner meta %02x", it.node, ch, meta);
                u32 rel = (u32)__builtin_popcount((u32)nd.imask & ~(0xffffffffu << slot_index));
                OrcAabb b;
                for (int a = 0; a < 3; a++) {  // intersection with parent bounds, aabb.rs:98-103
                    b.min[a] = smax(q.min[a], it.bounds.min[a]);
                    b.max[a] = smin(q.max[a], it.bounds.max[a]);
                }
                stack.push_back({nd.child_base_idx + rel, it.depth + 1, b});
            } else {
                u32 first = nd.primitive_base_idx + (meta & 0x1f);
                u32 unary = meta >> 5;
                if (unary != 1 && unary != 3 && unary != 7) VFAIL(-10, "node %u slot %d bad unary %u", it.node, ch, unary);
                for (int i = 0; i < 3; i++) {
                    if (!(meta & (0x20 << i))) continue;
                    u32 slot = first + i;
                    if (slot >= n_slots) VFAIL(-11, "prim slot %u out of range", slot);
                    if (seen_slot[slot]) VFAIL(-12, "prim slot %u referenced twice", slot);
                    seen_slot[slot] = 1;
                    u32 prim = bvh.primitive_indices[slot];
                    if (prim >= n) VFAIL(-13, "prim id %u out of range", prim);
                    if (seen_prim[prim] && !splits) VFAIL(-14, "primitive %u referenced twice", prim);
                    seen_prim[prim] = 1;
                    if (splits) continue;
                    for (int a = 0; a < 3; a++) {
                        if (!(prim_aabbs[prim].min[a] >= it.bounds.min[a] - 1.0e-5f) || !(prim_aabbs[prim].max[a] <= it.bounds.max[a] + 1.0e-5f))
                            VFAIL(-15, "primitive %u does not fit in node %u bounds", prim, it.node);
                        if (!(prim_aabbs[prim].min[a] >= q.min[a] - 1.0e-5f * fmaxf(1.f, fabsf(q.min[a]))) ||
                            !(prim_aabbs[prim].max[a] <= q.max[a] + 1.0e-5f * fmaxf(1.f, fabsf(q.max[a]))))
                            VFAIL(-16, "primitive %u does not fit quantised child box (node %u slot %d)", prim, it.node, ch);
                    }
                }
            }
        }
    }
    if (visited != bvh.nodes.size()) VFAIL(-17, "visited %zu of %zu nodes", visited, bvh.nodes.size());
    for (size_t i = 0; i < n; i++)
        if (!seen_prim[i]) VFAIL(-18, "primitive %zu unreachable", i);
    if (max_depth >= 32) VFAIL(-19, "depth %zu >= traversal stack size 32", max_depth);
    return 0;
}

// ---------------------------------------------------------------------------------------------------------
// C surface
// ---------------------------------------------------------------------------------------------------------
extern "C" {

int orc_max_threads(void) { return clamp_threads(0); }

void orc_tri_aabbs(const OrcTriangle* tris, size_t n, OrcAabb* out) {
    for (size_t i = 0; i < n; i++) out[i] = tri_aabb(tris[i]);
}

void orc_make_rays(const float* od6, size_t n, float tmin, float tmax, OrcRay* out) {
    for (size_t i = 0; i < n; i++) {
        OrcRay r;
        memset(&r, 0, sizeof(r));
        for (int k = 0; k < 3; k++) {
            r.origin[k] = od6[i * 6 + k];
            r.direction[k] = od6[i * 6 + 3 + k];
            r.inv_direction[k] = safe_inverse(r.direction[k]);
        }
        r.tmin = tmin;
        r.tmax = tmax;
        out[i] = r;
    }
}

void orc_morton_sort(const OrcAabb* aabbs, size_t n, int precision, u64* codes_lo, u64* codes_hi, u32* order_out,
                     OrcAabb* total_aabb_out) {
    std::vector<OrcBvh2Node> cur;
    std::vector<u32> idx(n);
    for (size_t i = 0; i < n; i++) idx[i] = (u32)i;
    OrcAabb total = init_leaves(aabbs, idx.data(), n, cur, 1);
    if (total_aabb_out) *total_aabb_out = total;
    if (n == 0) return;
    std::vector<Morton> m;
    gen_mortons(cur, total, precision, m, 1);
    for (size_t i = 0; i < n; i++) {
        if (codes_lo) codes_lo[i] = m[i].lo;
        if (codes_hi) codes_hi[i] = m[i].hi;
    }
    sort_mortons(m, precision, 1);
    if (order_out)
        for (size_t i = 0; i < n; i++) order_out[i] = m[i].index;
}

OrcBvh2* orc_ploc_build(const OrcAabb* aabbs, const u32* indices, size_t n, u32 search_distance, int precision,
                        size_t search_depth_threshold, int threads) {
    threads = clamp_threads(threads);
    OrcBvh2* bvh = new OrcBvh2();
    // bvh2/mod.rs:106-121 reset_for_reuse
    bvh->primitive_indices.assign(indices, indices + n);
    if (n == 0) return bvh;
    std::vector<OrcBvh2Node> cur;
    OrcAabb total = init_leaves(aabbs, indices, n, cur, threads);
    build_ploc_from_leaves(*bvh, cur, total, search_distance, precision, search_depth_threshold, threads);
    return bvh;
}
void orc_bvh2_free(OrcBvh2* b) { delete b; }
size_t orc_bvh2_node_count(const OrcBvh2* b) { return b->nodes.size(); }
size_t orc_bvh2_prim_count(const OrcBvh2* b) { return b->primitive_indices.size(); }
size_t orc_bvh2_max_depth(const OrcBvh2* b) { return b->max_depth; }
size_t orc_bvh2_ploc_iterations(const OrcBvh2* b) { return b->ploc_iterations; }
void orc_bvh2_get(const OrcBvh2* b, OrcBvh2Node* nodes, u32* primitive_indices, u32* parents) {
    if (nodes && !b->nodes.empty()) memcpy(nodes, b->nodes.data(), b->nodes.size() * sizeof(OrcBvh2Node));
    if (primitive_indices && !b->primitive_indices.empty())
        memcpy(primitive_indices, b->primitive_indices.data(), b->primitive_indices.size() * 4);
    if (parents && !b->parents.empty()) memcpy(parents, b->parents.data(), b->parents.size() * 4);
}
OrcBvh2* orc_bvh2_from(const OrcBvh2Node* nodes, size_t n_nodes, const u32* primitive_indices, size_t n_prims, size_t max_depth) {
    OrcBvh2* b = new OrcBvh2();
    b->nodes.assign(nodes, nodes + n_nodes);
    b->primitive_indices.assign(primitive_indices, primitive_indices + n_prims);
    b->max_depth = max_depth;
    return b;
}
int orc_bvh2_validate(const OrcBvh2* b, const OrcAabb* prim_aabbs, size_t n, int tight_fit, char* msg) {
    return bvh2_validate(*b, prim_aabbs, n, tight_fit != 0, msg);
}
void orc_bvh2_compute_parents(OrcBvh2* b) { compute_parents(*b); }
// Bvh2::reorder_in_stack_traversal_order (bvh2/mod.rs:462-500): a pure re-indexing (parents before children, sibling pairs in the
// pop order of a stack that receives the first child's pair before the second's). Not on the GPU path yet;
// restated so that the reference's test_reinsertion (bvh2/reinsertion.rs:394-436) runs here exactly as written.
void orc_bvh2_reorder_in_stack_traversal_order(OrcBvh2* b) {
    OrcBvh2& bvh = *b;
    if (bvh.nodes.size() < 2) return;
    std::vector<OrcBvh2Node> new_nodes;
    new_nodes.reserve(bvh.nodes.size());
    std::vector<u32> mapping(bvh.nodes.size(), 0);
    std::vector<u32> stack;
    stack.push_back(bvh.nodes[0].first_index);
    new_nodes.push_back(bvh.nodes[0]);
    while (!stack.empty()) {
        const u32 cur = stack.back();
        stack.pop_back();
        const OrcBvh2Node node_a = bvh.nodes[cur], node_b = bvh.nodes[(size_t)cur + 1];
        if (!is_leaf(node_a)) stack.push_back(node_a.first_index);
        if (!is_leaf(node_b)) stack.push_back(node_b.first_index);
        const u32 new_idx = (u32)new_nodes.size();
        mapping[cur] = new_idx;
        mapping[(size_t)cur + 1] = new_idx + 1;
        new_nodes.push_back(node_a);
        new_nodes.push_back(node_b);
    }
    for (OrcBvh2Node& n : new_nodes)
        if (!is_leaf(n)) n.first_index = mapping[n.first_index];
    bvh.nodes = std::move(new_nodes);
    if (!bvh.parents.empty()) compute_parents(bvh);  // update_parents
    bvh.children_are_ordered_after_parents = true;
}
void orc_bvh2_collapse(OrcBvh2* b, u32 max_prims, float traversal_cost) { collapse(*b, max_prims, traversal_cost); }
int orc_bvh2_has_parents(const OrcBvh2* b) { return b->parents.empty() ? 0 : 1; }
void orc_bvh2_ray_traverse(const OrcBvh2* b, const OrcTriangle* bvh_tris, const OrcRay* rays, size_t n, OrcRayHit* hits, int threads,
                           u64* counters) {
    bvh2_traverse_batch<0>(*b, bvh_tris, rays, n, hits, nullptr, nullptr, threads, counters);
}
void orc_bvh2_ray_traverse_miss(const OrcBvh2* b, const OrcTriangle* bvh_tris, const OrcRay* rays, size_t n, u8* miss, int threads,
                                u64* counters) {
    bvh2_traverse_batch<1>(*b, bvh_tris, rays, n, nullptr, miss, nullptr, threads, counters);
}
void orc_bvh2_ray_traverse_anyhit_count(const OrcBvh2* b, const OrcTriangle* bvh_tris, const OrcRay* rays, size_t n, u32* counts,
                                        int threads) {
    bvh2_traverse_batch<2>(*b, bvh_tris, rays, n, nullptr, nullptr, counts, threads, nullptr);
}
// bvh2/builder.rs:17-91: [pre-splits ->] PLOC -> reinsertion -> collapse -> reinsertion
OrcBvh2* orc_build_bvh2_from_tris(const OrcTriangle* tris, size_t n, u32 search_distance, size_t search_depth_threshold,
                                  float reinsertion_batch_ratio, float post_collapse_multiplier, int precision, u32 max_prims_per_leaf,
                                  float collapse_traversal_cost, int pre_split, int threads, double* core_seconds) {
    threads = clamp_threads(threads);
    std::vector<OrcAabb> aabbs(n);
    std::vector<u32> indices(n);
    auto t0 = std::chrono::steady_clock::now();
    if (pre_split) {  // bvh2/builder.rs:24-58
        float avg, largest;
        presplit_inputs(tris, n, aabbs, indices, &avg, &largest);
        t0 = std::chrono::steady_clock::now();
        split_aabbs_preset(aabbs, indices, tris, avg, largest);
    } else {
#pragma omp parallel for num_threads(threads) if (n > 100000)
        for (long i = 0; i < (long)n; i++) {
            aabbs[i] = tri_aabb(tris[i]);
            indices[i] = (u32)i;
        }
    }
    OrcBvh2* bvh2 = orc_ploc_build(aabbs.data(), indices.data(), aabbs.size(), search_distance, precision, search_depth_threshold, threads);
    bvh2->uses_spatial_splits = pre_split != 0;  // :69
    orc_reinsertion_run(bvh2, reinsertion_batch_ratio, nullptr, 0, threads);
    collapse(*bvh2, std::min<u32>(std::max<u32>(max_prims_per_leaf, 1), 255), collapse_traversal_cost);
    orc_reinsertion_run(bvh2, reinsertion_batch_ratio * post_collapse_multiplier, nullptr, 0, threads);
    auto t1 = std::chrono::steady_clock::now();
    if (core_seconds) *core_seconds += std::chrono::duration<double>(t1 - t0).count();
    return bvh2;
}
void orc_bvh2_refit_all(OrcBvh2* b) { refit_all(*b); }
void orc_ploc_full_rebuild(OrcBvh2* b, u32 search_distance, int precision, size_t search_depth_threshold, int threads) {
    full_rebuild(*b, search_distance, precision, search_depth_threshold, clamp_threads(threads));
}
void orc_ploc_partial_rebuild(OrcBvh2* b, const u8* should_remove, u32 search_distance, int precision, size_t search_depth_threshold,
                              int threads) {
    partial_rebuild(*b, should_remove, search_distance, precision, search_depth_threshold, clamp_threads(threads));
}
void orc_compute_rebuild_path_flags(const OrcBvh2* b, const u32* leaves, size_t n, u8* flags_out) {
    std::vector<u8> f;
    compute_rebuild_path_flags(*b, leaves, n, f);
    if (!f.empty()) memcpy(flags_out, f.data(), f.size());
}
void orc_bvh2_set_node_aabbs(OrcBvh2* b, const u32* node_ids, const OrcAabb* aabbs, size_t n) {  // Bvh2Node::set_aabb, physics.rs:446
    for (size_t k = 0; k < n; k++) {
        b->nodes[node_ids[k]].aabb = aabbs[k];
        b->nodes[node_ids[k]].aabb._p0 = b->nodes[node_ids[k]].aabb._p1 = 0.f;
    }
}
void orc_bvh2_set_leaf_aabbs(OrcBvh2* b, const OrcAabb* prim_aabbs) {
    for (auto& nd : b->nodes)
        if (is_leaf(nd)) {
            OrcAabb a = prim_aabbs[b->primitive_indices[nd.first_index]];
            for (u32 k = 1; k < nd.prim_count; k++) a = aabb_union(a, prim_aabbs[b->primitive_indices[nd.first_index + k]]);
            nd.aabb = a;
            nd.aabb._p0 = nd.aabb._p1 = 0.f;
        }
}

void orc_reinsertion_run(OrcBvh2* b, float ratio, const float* seq, size_t n_seq, int threads) {
    ReinsertionOptimizer opt;
    opt.threads = clamp_threads(threads);
    opt.run(*b, ratio, seq, n_seq);
    b->last_applied = opt.applied;
}
void orc_reinsertion_run_with_candidates(OrcBvh2* b, const u32* ids, size_t n, u32 iterations, int threads) {
    ReinsertionOptimizer opt;
    opt.threads = clamp_threads(threads);
    opt.run_with_candidates(*b, ids, n, iterations);
    b->last_applied = opt.applied;
}
size_t orc_reinsertion_last_applied(const OrcBvh2* b) { return b->last_applied; }
void orc_set_refit_full(int full) { g_refit_full = full != 0; }

size_t orc_bvh2_aabb_traverse(const OrcBvh2* b, const OrcAabb* queries, size_t n, u32* counts, u32* leaf_ids, size_t cap) {
    return query_batch(n, counts, leaf_ids, cap, [&](size_t i, auto emit) { bvh2_query_one<0>(*b, queries[i].min, emit); });
}
size_t orc_bvh2_point_traverse(const OrcBvh2* b, const float* points4, size_t n, u32* counts, u32* leaf_ids, size_t cap) {
    return query_batch(n, counts, leaf_ids, cap, [&](size_t i, auto emit) { bvh2_query_one<1>(*b, points4 + 4 * i, emit); });
}
size_t orc_cwbvh_aabb_traverse(const OrcCwBvh* c, const OrcAabb* queries, size_t n, const float* dir3, u32* counts, u32* prim_ids, size_t cap) {
    return query_batch(n, counts, prim_ids, cap, [&](size_t i, auto emit) { cwbvh_query_one<0>(*c, queries[i].min, dir3, emit); });
}
size_t orc_cwbvh_point_traverse(const OrcCwBvh* c, const float* points4, size_t n, const float* dir3, u32* counts, u32* prim_ids, size_t cap) {
    return query_batch(n, counts, prim_ids, cap, [&](size_t i, auto emit) { cwbvh_query_one<1>(*c, points4 + 4 * i, dir3, emit); });
}

// splits.rs:49-125 on caller arrays with room for `cap` entries; returns the new count (which may exceed cap: nothing is
// written past cap, call again with more room)
size_t orc_split_aabbs_precise(OrcAabb* aabbs, u32* indices, size_t n, size_t cap, const OrcTriangle* tris, float area_thresh_low,
                               float area_thresh_high, float split_factor_low, float split_factor_high, u32 max_iterations,
                               u32 split_tests) {
    std::vector<OrcAabb> a(aabbs, aabbs + n);
    std::vector<u32> idx(indices, indices + n);
    split_aabbs_precise(a, idx, tris, area_thresh_low, area_thresh_high, split_factor_low, split_factor_high, max_iterations, split_tests);
    size_t m = std::min(cap, a.size());
    memcpy(aabbs, a.data(), m * sizeof(OrcAabb));
    memcpy(indices, idx.data(), m * 4);
    return a.size();
}
// the builders' pre-split prologue (cwbvh/builder.rs:27-54): triangle AABBs, avg / largest half area, split_aabbs_preset
size_t orc_presplit_tris(const OrcTriangle* tris, size_t n, OrcAabb* aabbs_out, u32* indices_out, size_t cap, float* avg_half_area,
                         float* largest_half_area) {
    std::vector<OrcAabb> a;
    std::vector<u32> idx;
    float avg = 0.f, largest = 0.f;
    presplit_inputs(tris, n, a, idx, &avg, &largest);
    if (avg_half_area) *avg_half_area = avg;
    if (largest_half_area) *largest_half_area = largest;
    split_aabbs_preset(a, idx, tris, avg, largest);
    size_t m = std::min(cap, a.size());
    if (aabbs_out) memcpy(aabbs_out, a.data(), m * sizeof(OrcAabb));
    if (indices_out) memcpy(indices_out, idx.data(), m * 4);
    return a.size();
}
void orc_bvh2_set_uses_spatial_splits(OrcBvh2* b, int v) { b->uses_spatial_splits = v != 0; }
void orc_cwbvh_set_uses_spatial_splits(OrcCwBvh* c, int v) { c->uses_spatial_splits = v != 0; }

OrcCwBvh* orc_bvh2_to_cwbvh(const OrcBvh2* b, u32 max_prims_per_leaf, int order_children, int include_exact) {
    return bvh2_to_cwbvh(*b, max_prims_per_leaf, order_children != 0, include_exact != 0);
}
OrcCwBvh* orc_cwbvh_from(const OrcCwBvhNode* nodes, size_t n_nodes, const u32* primitive_indices, size_t n_prims, const OrcAabb* total) {
    OrcCwBvh* c = new OrcCwBvh();
    c->nodes.assign(nodes, nodes + n_nodes);
    c->primitive_indices.assign(primitive_indices, primitive_indices + n_prims);
    c->total_aabb = total ? *total : OrcAabb{};
    return c;
}
void orc_cwbvh_free(OrcCwBvh* c) { delete c; }
size_t orc_cwbvh_node_count(const OrcCwBvh* c) { return c->nodes.size(); }
size_t orc_cwbvh_prim_count(const OrcCwBvh* c) { return c->primitive_indices.size(); }
void orc_cwbvh_get(const OrcCwBvh* c, OrcCwBvhNode* nodes, u32* primitive_indices, OrcAabb* total) {
    if (nodes && !c->nodes.empty()) memcpy(nodes, c->nodes.data(), c->nodes.size() * sizeof(OrcCwBvhNode));
    if (primitive_indices && !c->primitive_indices.empty())
        memcpy(primitive_indices, c->primitive_indices.data(), c->primitive_indices.size() * 4);
    if (total) *total = c->total_aabb;
}
int orc_cwbvh_validate(const OrcCwBvh* c, const OrcAabb* prim_aabbs, size_t n, char* msg) { return cwbvh_validate(*c, prim_aabbs, n, msg); }

// ---- CwBvh::order_children / order_node_children as a separate pass (cwbvh/mod.rs:520-735) -------------------------------
// Not on the GPU path yet; restated here so that the next widening step has a pinned checker
// (tests/mod.rs:351-385 order_children_cwbvh, :431-445).
namespace {
inline bool cw_is_child_empty(const OrcCwBvhNode& n, int ch) { return n.child_meta[ch] == 0; }          // node.rs:284-286
inline bool cw_is_leaf(const OrcCwBvhNode& n, int ch) { return (n.imask & (1u << ch)) == 0; }           // node.rs:279-281
inline u32 cw_child_node_index(const OrcCwBvhNode& n, int ch) {                                        // node.rs:299-304
    const u32 slot_index = (u32)(n.child_meta[ch] & 0b11111) - 24u;
    return n.child_base_idx + (u32)__builtin_popcount((u32)n.imask & ~(0xffffffffu << slot_index));
}
inline OrcAabb cw_node_aabb_compressed(const OrcCwBvhNode& n) {                                        // node.rs:261-275
    OrcAabb a = aabb_empty();
    for (int k = 0; k < 3; k++) {
        u32 bits = (u32)n.e[k] << 23;
        float e;
        memcpy(&e, &bits, 4);
        a.min[k] = n.p[k];
        a.max[k] = n.p[k] + e * 255.0f;  // p + e * NQ_SCALE
    }
    return a;
}
inline V3 aabb_center(const OrcAabb& a) { return (v3(a.max) + v3(a.min)) * 0.5f; }                      // aabb.rs:113-115

void cwbvh_order_node_children(OrcCwBvh& bvh, const OrcAabb* prim_aabbs, size_t node_index, bool direct_layout) {
    const OrcCwBvhNode old_node = bvh.nodes[node_index];
    const bool have_exact = !bvh.exact_node_aabbs.empty();
    auto node_aabb = [&](size_t i) { return have_exact ? bvh.exact_node_aabbs[i] : cw_node_aabb_compressed(bvh.nodes[i]); };  // :737-745
    const V3 center = aabb_center(cw_node_aabb_compressed(old_node));
    float cost[8][8];
    for (auto& row : cost)
        for (float& c : row) c = 3.40282347e+38f;
    size_t child_inner_count = 0;
    for (int ch = 0; ch < 8; ch++)
        if (!cw_is_child_empty(old_node, ch) && !cw_is_leaf(old_node, ch)) child_inner_count++;
    V3 old_child_centers[8] = {};
    for (int ch = 0; ch < 8; ch++) {
        if (cw_is_child_empty(old_node, ch)) continue;
        if (cw_is_leaf(old_node, ch)) {
            const u32 meta = old_node.child_meta[ch];  // child_primitives, node.rs:290-295
            const u32 start = old_node.primitive_base_idx + (meta & 0b11111u), count = (u32)__builtin_popcount(meta & 0b11100000u);
            OrcAabb a = aabb_empty();
            for (u32 i = 0; i < count; i++) {
                size_t prim_index = start + i;
                if (!direct_layout) prim_index = bvh.primitive_indices[prim_index];
                a = aabb_union(a, prim_aabbs[prim_index]);
            }
            old_child_centers[ch] = aabb_center(a);
        } else {
            old_child_centers[ch] = aabb_center(node_aabb(cw_child_node_index(old_node, ch)));
        }
    }
    for (int s = 0; s < 8; s++) {
        const V3 d = V3{(s & 0b100) ? -1.0f : 1.0f, (s & 0b010) ? -1.0f : 1.0f, (s & 0b001) ? -1.0f : 1.0f};
        for (int ch = 0; ch < 8; ch++) {
            if (cw_is_child_empty(old_node, ch)) continue;
            cost[ch][s] = dot(d, old_child_centers[ch] - center);
        }
    }
    size_t assignment[8];
    bool slot_filled[8] = {};
    for (size_t& a : assignment) a = (size_t)INVALID32;
    for (;;) {  // greedy: cheapest unfilled slot of any unassigned child
        float min_cost = 3.40282347e+38f;
        size_t min_slot = (size_t)INVALID32, min_index = (size_t)INVALID32;
        for (int ch = 0; ch < 8; ch++) {
            if (cw_is_child_empty(old_node, ch) || assignment[ch] != (size_t)INVALID32) continue;
            for (int sl = 0; sl < 8; sl++)
                if (!slot_filled[sl] && cost[ch][sl] < min_cost) {
                    min_cost = cost[ch][sl];
                    min_slot = (size_t)sl;
                    min_index = (size_t)ch;
                }
        }
        if (min_slot == (size_t)INVALID32) break;
        slot_filled[min_slot] = true;
        assignment[min_index] = min_slot;
    }
    OrcCwBvhNode new_node = old_node;
    new_node.imask = 0;
    for (int ch = 0; ch < 8; ch++) new_node.child_meta[ch] = 0;
    for (int ch = 0; ch < 8; ch++) {
        if (cw_is_child_empty(old_node, ch)) continue;
        const size_t new_ch = assignment[ch];
        if (new_ch >= 8) {
            fprintf(stderr, "oracle: order_node_children left a child unassigned -- the reference asserts\n");
            abort();
        }
        if (cw_is_leaf(old_node, ch)) new_node.child_meta[new_ch] = old_node.child_meta[ch];
        else {
            new_node.imask |= (uint8_t)(1u << new_ch);
            new_node.child_meta[new_ch] = (uint8_t)((24 + new_ch) | 0b00100000);
        }
        new_node.child_min_x[new_ch] = old_node.child_min_x[ch];
        new_node.child_max_x[new_ch] = old_node.child_max_x[ch];
        new_node.child_min_y[new_ch] = old_node.child_min_y[ch];
        new_node.child_max_y[new_ch] = old_node.child_max_y[ch];
        new_node.child_min_z[new_ch] = old_node.child_min_z[ch];
        new_node.child_max_z[new_ch] = old_node.child_max_z[ch];
    }
    if (child_inner_count == 0) {
        bvh.nodes[node_index] = new_node;
        return;
    }
    OrcCwBvhNode old_child_nodes[8];
    OrcAabb old_child_exact[8];
    for (int ch = 0; ch < 8; ch++) {
        if (cw_is_child_empty(old_node, ch) || cw_is_leaf(old_node, ch)) continue;
        const u32 ci = cw_child_node_index(old_node, ch);
        old_child_nodes[ch] = bvh.nodes[ci];
        if (have_exact) old_child_exact[ch] = bvh.exact_node_aabbs[ci];
    }
    for (int ch = 0; ch < 8; ch++) {
        if (cw_is_child_empty(old_node, ch) || assignment[ch] == (size_t)INVALID32 || cw_is_leaf(old_node, ch)) continue;
        const size_t new_idx = cw_child_node_index(new_node, (int)assignment[ch]);
        bvh.nodes[new_idx] = old_child_nodes[ch];
        if (have_exact) bvh.exact_node_aabbs[new_idx] = old_child_exact[ch];
    }
    bvh.nodes[node_index] = new_node;
}
}  // namespace
void orc_cwbvh_order_node_children(OrcCwBvh* c, const OrcAabb* prim_aabbs, size_t node_index, int direct_layout) {
    cwbvh_order_node_children(*c, prim_aabbs, node_index, direct_layout != 0);
}
void orc_cwbvh_order_children(OrcCwBvh* c, const OrcAabb* prim_aabbs, int direct_layout) {  // cwbvh/mod.rs:520-524
    for (size_t i = 0; i < c->nodes.size(); i++) cwbvh_order_node_children(*c, prim_aabbs, i, direct_layout != 0);
}
// CwBvh::exact_node_aabbs (cwbvh/mod.rs:47): bvh2.nodes.len() entries, Aabb::empty() beyond the wide nodes; returns the length
size_t orc_cwbvh_exact_node_aabbs(const OrcCwBvh* c, OrcAabb* out, size_t cap) {
    size_t m = std::min(cap, c->exact_node_aabbs.size());
    if (out && m) memcpy(out, c->exact_node_aabbs.data(), m * sizeof(OrcAabb));
    return c->exact_node_aabbs.size();
}

OrcCwBvh* orc_build_cwbvh_from_tris(const OrcTriangle* tris, size_t n, u32 search_distance, size_t search_depth_threshold,
                                    float reinsertion_batch_ratio, int precision, u32 max_prims_per_leaf, int pre_split, int threads,
                                    double* core_seconds) {
    threads = clamp_threads(threads);
    auto t0 = std::chrono::steady_clock::now();  // cwbvh/builder.rs:62
    // PlocBuilder::build over &[Triangle]: aabbs[i].aabb() is evaluated inside the leaf-init loop (ploc/mod.rs:221,240)
    std::vector<OrcAabb> aabbs(n);
    std::vector<u32> indices(n);
    if (pre_split) {  // cwbvh/builder.rs:27-61; the clock starts after the AABB / average pass (:45)
        float avg, largest;
        presplit_inputs(tris, n, aabbs, indices, &avg, &largest);
        t0 = std::chrono::steady_clock::now();
        split_aabbs_preset(aabbs, indices, tris, avg, largest);
    } else {
#pragma omp parallel for num_threads(threads) schedule(static) if (threads > 1 && n >= 500000)
        for (long i = 0; i < (long)n; i++) {
            aabbs[i] = tri_aabb(tris[i]);
            indices[i] = (u32)i;
        }
    }
    OrcBvh2* bvh2 = orc_ploc_build(aabbs.data(), indices.data(), aabbs.size(), search_distance, precision, search_depth_threshold, threads);
    bvh2->uses_spatial_splits = pre_split != 0;  // :72
    orc_reinsertion_run(bvh2, reinsertion_batch_ratio, nullptr, 0, threads);
    u32 mp = std::min<u32>(std::max<u32>(max_prims_per_leaf, 1), 3);  // cwbvh/builder.rs:74 clamp(1,3)
    OrcCwBvh* c = bvh2_to_cwbvh(*bvh2, mp, true, false);
    auto t1 = std::chrono::steady_clock::now();
    if (core_seconds) *core_seconds += std::chrono::duration<double>(t1 - t0).count();
    delete bvh2;
    return c;
}

void orc_cwbvh_ray_traverse(const OrcCwBvh* c, const OrcTriangle* bvh_tris, const OrcRay* rays, size_t n, OrcRayHit* hits,
                            int threads, int use_simd, u64* counters) {
    traverse_batch<0>(*c, bvh_tris, rays, n, hits, nullptr, nullptr, threads, use_simd, counters);
}
void orc_cwbvh_ray_traverse_miss(const OrcCwBvh* c, const OrcTriangle* bvh_tris, const OrcRay* rays, size_t n, u8* miss,
                                 int threads, int use_simd, u64* counters) {
    traverse_batch<1>(*c, bvh_tris, rays, n, nullptr, miss, nullptr, threads, use_simd, counters);
}
void orc_cwbvh_ray_traverse_anyhit_count(const OrcCwBvh* c, const OrcTriangle* bvh_tris, const OrcRay* rays, size_t n,
                                         u32* counts, int threads) {
    traverse_batch<2>(*c, bvh_tris, rays, n, nullptr, nullptr, counts, threads, 1, nullptr);
}
float orc_triangle_intersect(const OrcTriangle* tri, const OrcRay* ray) { return tri_intersect(*tri, *ray); }
// the Aabb primitives the path is built on, exported so that the reference's own unit tests (aabb.rs:222-360) pin them
float orc_aabb_half_area(const OrcAabb* a) { return half_area(*a); }
void orc_aabb_union(const OrcAabb* a, const OrcAabb* b, OrcAabb* out) { *out = aabb_union(*a, *b); }
float orc_aabb_intersect_ray(const OrcAabb* a, const OrcRay* ray) { return aabb_intersect_ray(*a, *ray); }
int orc_aabb_intersect_aabb(const OrcAabb* a, const OrcAabb* b) { return aabb_intersect_aabb(a->min, a->max, b->min, b->max); }
int orc_aabb_contains_point(const OrcAabb* a, const float* p3) { return aabb_contains_point(a->min, a->max, p3); }
void orc_triangle_normal(const OrcTriangle* tri, float* out3) {
    V3 nrm = tri_normal(*tri);
    out3[0] = nrm.x;
    out3[1] = nrm.y;
    out3[2] = nrm.z;
}

// naive bit interleave used by tests to cross-check the magic-mask versions above
u64 orc_test_split3_64(u32 a) { return split_by_3_u64(a); }
void orc_test_split3_128(u64 a, u64* lo, u64* hi) {
    u128 x = split_by_3_u128(a);
    *lo = (u64)x;
    *hi = (u64)(x >> 64);
}
}  // extern "C"
