// Exercises include/obvhs.hpp (the C++ host side above the C ABI) on a GPU: builds, traversals, queries and rebuilds of a procedural
// scene, checked against a brute-force CPU intersection written here (double precision Moeller-Trumbore) and against each other.
// Built by obvhs_b200/build.py with g++ (no CUDA headers needed), run by tests/test_gpu_cpp_host.py. Exit code 0 = all checks hold.
// Without a CUDA device the Context constructor throws: the program prints the reason and exits 3 (there is no CPU fallback).
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../include/obvhs.hpp"

using namespace obvhs;

static uint32_t lcg_state = 12345u;
static float frand() {  // [0,1)
    lcg_state = lcg_state * 1664525u + 1013904223u;
    return (float)(lcg_state >> 8) * (1.0f / 16777216.0f);
}

static Triangle tri(const float a[3], const float b[3], const float c[3]) {
    Triangle t;
    std::memset(&t, 0, sizeof(t));
    for (int k = 0; k < 3; k++) {
        t.v0[k] = a[k];
        t.v1[k] = b[k];
        t.v2[k] = c[k];
    }
    return t;
}

static float height(int i, int j) { return 0.08f * std::sin(0.37f * (float)i) * std::cos(0.23f * (float)j) + 0.02f * std::sin(1.7f * (float)(i + j)); }

static std::vector<Triangle> make_scene(int grid, int extra) {
    std::vector<Triangle> t;
    const float s = 1.0f / (float)grid;
    for (int i = 0; i < grid; i++)
        for (int j = 0; j < grid; j++) {
            float p00[3] = {i * s, height(i, j), j * s}, p10[3] = {(i + 1) * s, height(i + 1, j), j * s};
            float p01[3] = {i * s, height(i, j + 1), (j + 1) * s}, p11[3] = {(i + 1) * s, height(i + 1, j + 1), (j + 1) * s};
            t.push_back(tri(p00, p10, p11));
            t.push_back(tri(p00, p11, p01));
        }
    for (int e = 0; e < extra; e++) {  // small floating triangles above the terrain
        float c[3] = {frand(), 0.2f + 0.5f * frand(), frand()}, v[3][3];
        for (auto& p : v)
            for (int k = 0; k < 3; k++) p[k] = c[k] + 0.03f * (frand() - 0.5f);
        t.push_back(tri(v[0], v[1], v[2]));
    }
    return t;
}

// two-sided ray / triangle test in double precision; returns +inf on a miss
static double brute_intersect(const Triangle& t, const Ray& r) {
    double e1[3], e2[3], d[3], o[3], p[3], s[3], q[3];
    for (int k = 0; k < 3; k++) {
        e1[k] = (double)t.v1[k] - t.v0[k];
        e2[k] = (double)t.v2[k] - t.v0[k];
        d[k] = r.direction[k];
        o[k] = r.origin[k];
    }
    p[0] = d[1] * e2[2] - d[2] * e2[1], p[1] = d[2] * e2[0] - d[0] * e2[2], p[2] = d[0] * e2[1] - d[1] * e2[0];
    const double det = e1[0] * p[0] + e1[1] * p[1] + e1[2] * p[2];
    if (std::fabs(det) < 1e-14) return INFINITY;
    for (int k = 0; k < 3; k++) s[k] = o[k] - t.v0[k];
    const double u = (s[0] * p[0] + s[1] * p[1] + s[2] * p[2]) / det;
    q[0] = s[1] * e1[2] - s[2] * e1[1], q[1] = s[2] * e1[0] - s[0] * e1[2], q[2] = s[0] * e1[1] - s[1] * e1[0];
    const double v = (d[0] * q[0] + d[1] * q[1] + d[2] * q[2]) / det;
    if (u < 0 || v < 0 || u + v > 1) return INFINITY;
    const double tt = (e2[0] * q[0] + e2[1] * q[1] + e2[2] * q[2]) / det;
    return (tt >= r.tmin && tt <= r.tmax) ? tt : INFINITY;
}

static int failures = 0;
#define CHECK(cond, ...)                               \
    do {                                               \
        if (!(cond)) {                                 \
            std::printf("FAILED %s:%d: ", __FILE__, __LINE__); \
            std::printf(__VA_ARGS__);                  \
            std::printf("\n");                         \
            failures++;                                \
        }                                              \
    } while (0)

int main() {
    std::vector<Triangle> tris = make_scene(48, 600);
    const size_t n = tris.size();
    std::vector<Ray> rays;
    std::vector<RayNew> args;
    for (int i = 0; i < 6000; i++) {
        float o[3], d[3];
        if (i % 2) {  // from above, roughly downwards
            o[0] = frand(), o[1] = 1.5f, o[2] = frand();
            d[0] = 0.3f * (frand() - 0.5f), d[1] = -1.0f, d[2] = 0.3f * (frand() - 0.5f);
        } else {  // through the scene box in any direction
            o[0] = 2.0f * frand() - 0.5f, o[1] = frand() - 0.1f, o[2] = 2.0f * frand() - 0.5f;
            d[0] = 0.5f - o[0] + 0.4f * (frand() - 0.5f), d[1] = 0.2f - o[1] + 0.4f * (frand() - 0.5f), d[2] = 0.5f - o[2] + 0.4f * (frand() - 0.5f);
        }
        const float len = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
        for (float& x : d) x /= len;
        if (i % 7 == 0) d[0] = 0.0f;  // safe_inverse path
        rays.push_back(ray_new_inf(o, d));
        RayNew a = {{o[0], o[1], o[2]}, 0.0f, {d[0], d[1], d[2]}, INFINITY};
        args.push_back(a);
    }
    const size_t m = rays.size();

    {  // host-only pieces of the header (these run without a GPU too)
        CHECK(safe_inverse(0.0f) == 8388608.0f && safe_inverse(-0.0f) == -8388608.0f && safe_inverse(4.0f) == 0.25f, "safe_inverse");
        std::vector<Bvh2Node> tree(3);
        std::memset(tree.data(), 0, tree.size() * sizeof(Bvh2Node));
        tree[0].first_index = 1;                             // root: inner, children 1 and 2
        tree[1].prim_count = 2, tree[1].first_index = 0;     // leaf over slots 0..1
        tree[2].prim_count = 1, tree[2].first_index = 2;     // leaf over slot 2
        const std::vector<uint32_t> p2n = compute_primitives_to_nodes(tree, {2u, 0u, 1u});
        CHECK(p2n.size() == 3 && p2n[2] == 1 && p2n[0] == 1 && p2n[1] == 2, "compute_primitives_to_nodes");
        if (failures) return 1;
    }

    try {
        Context ctx(0);
        const BvhBuildParams params = BvhBuildParams::fast_build();
        double core = 0.0;
        CwBvh cw = build_cwbvh_from_tris(ctx, tris.data(), n, params, &core);
        CHECK(core > 0.0, "core_build_time not incremented");
        CHECK(cw.prim_count() == n, "prim_count %zu != %zu", cw.prim_count(), n);
        std::vector<CwBvhNode> nodes;
        std::vector<uint32_t> prims;
        Aabb total;
        cw.download(&nodes, &prims, &total);
        std::vector<uint32_t> sorted = prims;
        std::sort(sorted.begin(), sorted.end());
        for (size_t i = 0; i < n; i++) CHECK(sorted[i] == i, "primitive_indices is not a permutation at %zu", i);

        // closest hits against brute force
        std::vector<RayHit> hits(m);
        cw.ray_traverse(rays.data(), m, hits.data());
        size_t n_hit = 0, disagree = 0;
        for (size_t i = 0; i < m; i++) {
            double best = INFINITY;
            for (size_t k = 0; k < n; k++) best = std::min(best, brute_intersect(tris[k], rays[i]));
            const bool hit = hits[i].primitive_id != INVALID_ID;
            n_hit += hit;
            if (hit != (best < INFINITY)) disagree++;  // (edge-grazing rays may differ between f32 and f64)
            else if (hit) {
                CHECK(std::fabs((double)hits[i].t - best) <= 1e-4 * std::max(1.0, best), "ray %zu: t %g, brute force %g", i, hits[i].t, best);
                CHECK(brute_intersect(tris[prims[hits[i].primitive_id]], rays[i]) < INFINITY, "ray %zu: reported triangle is not hit", i);
            }
        }
        CHECK(n_hit > m / 3, "only %zu of %zu rays hit", n_hit, m);
        CHECK(disagree * 200 <= m, "%zu of %zu rays disagree with brute force on hit / miss", disagree, m);

        // the same rays as Ray::new arguments: bit-identical hits; constructor alone equals the host constructor
        std::vector<RayHit> hits2(m);
        cw.ray_traverse(args.data(), m, hits2.data());
        CHECK(std::memcmp(hits.data(), hits2.data(), m * sizeof(RayHit)) == 0, "Ray::new path differs from the Ray-struct path");
        // ... and as (origin, direction) records with one pair of bounds for the batch (all rays here are Ray::new_inf)
        std::vector<RayOd> od(m);
        bool uniform = true;
        for (size_t i = 0; i < m; i++) {
            std::memcpy(od[i].origin, args[i].origin, 12);
            std::memcpy(od[i].direction, args[i].direction, 12);
            uniform = uniform && args[i].tmin == args[0].tmin && args[i].tmax == args[0].tmax;
        }
        if (uniform) {
            std::vector<RayHit> hits3(m);
            cw.ray_traverse(od.data(), m, hits3.data(), args[0].tmin, args[0].tmax);
            CHECK(std::memcmp(hits.data(), hits3.data(), m * sizeof(RayHit)) == 0, "24-byte ray path differs from the Ray-struct path");
        }
        std::vector<Ray> built(m);
        ray_new_batch(ctx, args.data(), m, built.data());
        CHECK(std::memcmp(built.data(), rays.data(), m * sizeof(Ray)) == 0, "device Ray::new differs from the host constructor");

        // miss / all-hit flavours agree with the closest hit
        std::vector<uint8_t> miss(m);
        std::vector<uint32_t> counts(m);
        cw.ray_traverse_miss(rays.data(), m, miss.data());
        cw.ray_traverse_anyhit_count(rays.data(), m, counts.data());
        for (size_t i = 0; i < m; i++) {
            CHECK((miss[i] != 0) == (hits[i].primitive_id == INVALID_ID), "ray %zu: miss flag", i);
            CHECK((counts[i] > 0) == (hits[i].primitive_id != INVALID_ID), "ray %zu: any-hit count %u", i, counts[i]);
        }

        // staged path == one call (cwbvh/builder.rs:20-85), byte for byte
        PlocBuilder ploc(ctx);
        ReinsertionOptimizer opt(ctx);
        Bvh2 b = ploc.build((PlocSearchDistance)params.ploc_search_distance, tris.data(), n, (SortPrecision)params.sort_precision,
                            (size_t)params.search_depth_threshold);
        CHECK(b.node_count() == 2 * n - 1, "Bvh2 node count %zu", b.node_count());
        opt.run(b, params.reinsertion_batch_ratio);
        CwBvh cw2 = bvh2_to_cwbvh(b, std::min(3u, std::max(1u, params.max_prims_per_leaf)), true, false);
        std::vector<CwBvhNode> nodes2;
        std::vector<uint32_t> prims2;
        cw2.download(&nodes2, &prims2);
        CHECK(nodes2.size() == nodes.size() && std::memcmp(nodes2.data(), nodes.data(), nodes.size() * sizeof(CwBvhNode)) == 0, "staged CWBVH differs");
        CHECK(prims2 == prims, "staged primitive_indices differ");

        // Bvh2 builder + traversal sees the same distances
        Bvh2 b2 = build_bvh2_from_tris(ctx, tris.data(), n, BvhBuildParams::medium_build());
        std::vector<RayHit> hb(m);
        b2.ray_traverse(rays.data(), m, hb.data());
        size_t both = 0, same = 0;
        for (size_t i = 0; i < m; i++)
            if (hb[i].primitive_id != INVALID_ID && hits[i].primitive_id != INVALID_ID) {
                both++;
                same += std::memcmp(&hb[i].t, &hits[i].t, 4) == 0;
            }
        CHECK(both * 100 >= n_hit * 99 && same * 100 >= both * 99, "Bvh2 / CwBvh agree on %zu of %zu common hits (%zu CwBvh hits)", same, both, n_hit);

        // broad-phase queries: the box around everything reports every primitive once
        Aabb all = total;
        for (int k = 0; k < 3; k++) all.min[k] -= 1.0f, all.max[k] += 1.0f;
        const float dir[3] = {1.0f, 1.0f, 1.0f};
        QueryResult q = cw.aabb_traverse(&all, 1, dir);
        CHECK(q.counts[0] == n && q.ids.size() == n, "aabb_traverse reported %u of %zu primitives", q.counts[0], n);
        QueryResult q2 = b2.aabb_traverse(&all, 1);
        std::vector<Bvh2Node> bn;
        b2.download(&bn, nullptr);
        size_t leaves = 0;
        for (const Bvh2Node& nd : bn) leaves += nd.prim_count != 0;
        CHECK(q2.counts[0] == leaves, "Bvh2 aabb_traverse reported %u of %zu leaves", q2.counts[0], leaves);

        // rebuild: flag the paths of a few leaves, rebuild partially, then fully; the tree keeps every primitive
        b.compute_parents();
        std::vector<Bvh2Node> pn;
        b.download(&pn, nullptr);
        std::vector<uint32_t> some;
        for (uint32_t i = 0; i < pn.size() && some.size() < 50; i += 97)
            if (pn[i].prim_count != 0) some.push_back(i);
        std::vector<uint8_t> flags = compute_rebuild_path_flags(b, some.data(), some.size());
        ploc.partial_rebuild(b, flags.data(), PlocSearchDistance::Minimum, SortPrecision::U64, 0);
        ploc.full_rebuild(b, PlocSearchDistance::Low, SortPrecision::U128, 2);
        std::vector<uint32_t> bp;
        b.download(nullptr, &bp);
        std::sort(bp.begin(), bp.end());
        CHECK(bp.size() == n && bp.front() == 0 && bp.back() == n - 1, "rebuilt tree lost primitives");

        // layout passes: Bvh2::reorder_in_stack_traversal_order keeps the hits and puts every child pair behind its parent;
        // CwBvh::order_children on a tree converted WITHOUT ordering keeps the hits and is idempotent
        {
            Bvh2 b3 = build_bvh2_from_tris(ctx, tris.data(), n, BvhBuildParams::fast_build());
            std::vector<RayHit> h0(m), h1(m);
            b3.ray_traverse(rays.data(), m, h0.data());
            b3.reorder_in_stack_traversal_order();
            CHECK(b3.children_are_ordered_after_parents(), "reorder did not set children_are_ordered_after_parents");
            b3.ray_traverse(rays.data(), m, h1.data());
            size_t diff_t = 0;
            for (size_t i = 0; i < m; i++) diff_t += std::memcmp(&h0[i].t, &h1[i].t, 4) != 0;
            CHECK(diff_t == 0, "reorder changed %zu hit distances", diff_t);
            std::vector<Bvh2Node> rn;
            b3.download(&rn, nullptr);
            size_t out_of_order = 0;
            for (size_t i = 0; i < rn.size(); i++) out_of_order += rn[i].prim_count == 0 && rn[i].first_index <= i;
            CHECK(out_of_order == 0, "%zu inner nodes point backwards after the reorder", out_of_order);

            std::vector<Aabb> boxes(n);
            for (size_t i = 0; i < n; i++)
                for (int k = 0; k < 3; k++) {
                    boxes[i].min[k] = std::min(tris[i].v0[k], std::min(tris[i].v1[k], tris[i].v2[k]));
                    boxes[i].max[k] = std::max(tris[i].v0[k], std::max(tris[i].v1[k], tris[i].v2[k]));
                }
            CwBvh cu = bvh2_to_cwbvh(b, 3, false, true);
            cu.set_triangles(tris.data(), n);
            cu.ray_traverse(rays.data(), m, h0.data());
            cu.order_children(boxes.data(), n, false);
            cu.ray_traverse(rays.data(), m, h1.data());
            CHECK(std::memcmp(h0.data(), h1.data(), m * sizeof(RayHit)) == 0, "order_children changed the hits");
            std::vector<CwBvhNode> o1, o2;
            cu.download(&o1, nullptr);
            cu.order_children(boxes.data(), n, false);
            cu.download(&o2, nullptr);
            CHECK(std::memcmp(o1.data(), o2.data(), o1.size() * sizeof(CwBvhNode)) == 0, "order_children is not idempotent");
            CHECK(cu.exact_node_aabbs().size() >= o1.size(), "exact boxes lost");  // vec![Aabb::empty(); bvh2.nodes.len()], bvh2_to_cwbvh.rs:60-62
        }

        // errors surface as exceptions with the library's message
        bool threw = false;
        try {
            CwBvh bare = CwBvh::upload(ctx, nodes.data(), nodes.size(), prims.data(), prims.size(), total);
            bare.ray_traverse(rays.data(), m, hits2.data());  // no triangles attached
        } catch (const Error& e) {
            threw = e.status == OBVHS_ERR_INVALID_ARG && std::strlen(e.what()) > 0;
        }
        CHECK(threw, "traversal without triangles did not throw");

        std::printf("host_api: %zu tris, %zu CWBVH nodes, %zu rays, %zu hits, %llu kernel launches, core build %.3f ms -> %s\n", n, nodes.size(), m, n_hit,
                    (unsigned long long)ctx.launch_count(), core * 1e3, failures ? "FAILED" : "ok");
    } catch (const Error& e) {
        std::printf("host_api: obvhs::Error %d: %s\n", e.status, e.what());
        return 3;
    }
    return failures ? 1 : 0;
}
