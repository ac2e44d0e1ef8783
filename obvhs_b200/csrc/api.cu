// api.cu -- the extern "C" surface declared in include/obvhs_cuda.h. Thin: argument checks, host<->device staging,
// and the orchestration of build_cwbvh_from_tris (src/cwbvh/builder.rs:20-85).
#include <string.h>

#include "common.cuh"

int bvh2_set_leaf_aabbs_device(ObvhsContext* ctx, ObvhsBvh2* bvh, const ObvhsAabb* d_aabbs);
int bvh2_expand_nodes_device(ObvhsContext* ctx, const Node32* in, size_t n, ObvhsBvh2Node* d_out);
int bvh2_pack_nodes_device(ObvhsContext* ctx, const ObvhsBvh2Node* d_in, size_t n, Node32* out);

#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <type_traits>
cudaError_t obvhs_result_alloc(ObvhsContext* ctx, void** p, size_t bytes) {
    const size_t GRAIN = (size_t)2 << 20;
    size_t cap = (bytes + GRAIN - 1) / GRAIN * GRAIN;
    if (cap == 0) cap = GRAIN;
    {
        std::lock_guard<std::mutex> lock(ctx->result_mu);
        // best fit among cached blocks that waste at most 25 % (+ one grain)
        size_t best = (size_t)-1;
        for (size_t i = 0; i < ctx->result_cache.size(); i++) {
            const size_t c = ctx->result_cache[i].cap;
            if (c >= cap && c <= cap + cap / 4 + GRAIN && (best == (size_t)-1 || c < ctx->result_cache[best].cap)) best = i;
        }
        if (best != (size_t)-1) {
            *p = ctx->result_cache[best].p;
            ctx->result_live[*p] = ctx->result_cache[best].cap;
            ctx->result_cache.erase(ctx->result_cache.begin() + best);
            return cudaSuccess;
        }
    }
    double t0 = ctx->trace ? TraceScope::now() : 0.0;
    cudaError_t e = cudaMalloc(p, cap);
    if (e != cudaSuccess) {  // make room: drop the cache and retry once
        cudaGetLastError();
        std::vector<ObvhsContext::ResultBlock> drop;
        {
            std::lock_guard<std::mutex> lock(ctx->result_mu);
            drop.swap(ctx->result_cache);
        }
        for (auto& b : drop) cudaFree(b.p);
        e = cudaMalloc(p, cap);
    }
    if (ctx->trace) fprintf(stderr, "[obvhs trace]     result cudaMalloc %zu MB took %.3f ms\n", cap >> 20, (TraceScope::now() - t0) * 1e3);
    if (e == cudaSuccess) {
        std::lock_guard<std::mutex> lock(ctx->result_mu);
        ctx->result_live[*p] = cap;
    }
    return e;
}
// The block goes back to the context's cache. Work that still reads it was enqueued on ctx->stream, and so is whatever
// reuses it, so stream order makes the hand-over safe without a synchronisation.
void obvhs_result_free(ObvhsContext* ctx, void* p) {
    if (!p) return;
    void* evict = nullptr;
    {
        std::lock_guard<std::mutex> lock(ctx->result_mu);
        auto it = ctx->result_live.find(p);
        if (it == ctx->result_live.end()) return;
        ctx->result_cache.push_back({p, it->second});
        ctx->result_live.erase(it);
        if (ctx->result_cache.size() > 32) {  // bound the cache: the oldest block goes back to the driver
            evict = ctx->result_cache.front().p;
            ctx->result_cache.erase(ctx->result_cache.begin());
        }
    }
    if (evict) cudaFree(evict);
}
void obvhs_context_retain(ObvhsContext* ctx) { ctx->refs.fetch_add(1); }
static void context_teardown(ObvhsContext* ctx);
void obvhs_context_release(ObvhsContext* ctx) {
    if (ctx->refs.fetch_sub(1) == 1) context_teardown(ctx);
}
double TraceScope::now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

bool obvhs_is_device_ptr(const void* ptr) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, ptr) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

static thread_local ObvhsContext* tls_ctx = nullptr;

void* obvhs_arena_alloc(size_t bytes) {
    ObvhsContext* ctx = tls_ctx;
    if (!ctx) return nullptr;
    bytes = (bytes + 255) & ~(size_t)255;
    // first block (from the current one on) with room; blocks after the current one are empty
    while (ctx->arena_block < ctx->arena_blocks.size()) {
        ObvhsContext::ArenaBlock& b = ctx->arena_blocks[ctx->arena_block];
        if (ctx->arena_off + bytes <= b.cap) {
            void* p = b.p + ctx->arena_off;
            ctx->arena_off += bytes;
            ctx->arena_used += bytes;
            if (ctx->arena_used > ctx->arena_peak) ctx->arena_peak = ctx->arena_used;
            return p;
        }
        ctx->arena_block++;
        ctx->arena_off = 0;
    }
    size_t total = 0;
    for (auto& b : ctx->arena_blocks) total += b.cap;
    size_t cap = bytes > total ? bytes : total;  // at least double the arena
    if (cap < ((size_t)8 << 20)) cap = (size_t)8 << 20;
    char* p = nullptr;
    if (cudaMalloc((void**)&p, cap) != cudaSuccess) {
        cudaGetLastError();
        cap = bytes;
        if (cudaMalloc((void**)&p, cap) != cudaSuccess) {
            cudaGetLastError();
            return nullptr;
        }
    }
    ctx->arena_blocks.push_back({p, cap});
    ctx->arena_block = ctx->arena_blocks.size() - 1;
    ctx->arena_off = bytes;
    ctx->arena_used += bytes;
    if (ctx->arena_used > ctx->arena_peak) ctx->arena_peak = ctx->arena_used;
    return p;
}

namespace {
struct DeviceScope {  // makes the context's device current for the duration of a call and scopes its scratch
    int prev = -1;
    ObvhsContext* ctx;
    ObvhsContext* prev_ctx;
    size_t mark_block, mark_off, mark_used;
    explicit DeviceScope(ObvhsContext* c) : ctx(c) {
        cudaGetDevice(&prev);
        if (prev != ctx->device) cudaSetDevice(ctx->device);
        else prev = -1;
        prev_ctx = tls_ctx;
        tls_ctx = ctx;
        mark_block = ctx->arena_block;
        mark_off = ctx->arena_off;
        mark_used = ctx->arena_used;
        ctx->api_depth++;
    }
    ~DeviceScope() {
        ctx->api_depth--;
        if (ctx->api_depth == 0 && ctx->staged_host) {
            // a host buffer was copied with cudaMemcpyAsync (truly asynchronous when it is pinned): every entry point behaves
            // synchronously towards host inputs, so the caller may overwrite them right after the call
            cudaStreamSynchronize(ctx->stream);
            ctx->staged_host = false;
        }
        ctx->arena_block = mark_block;
        ctx->arena_off = mark_off;
        ctx->arena_used = mark_used;
        if (ctx->api_depth == 0 && ctx->arena_blocks.size() > 1) {
            // consolidate into one block sized for the peak so the next call bumps through contiguous memory
            cudaStreamSynchronize(ctx->stream);
            size_t total = 0;
            for (auto& b : ctx->arena_blocks) {
                total += b.cap;
                cudaFree(b.p);
            }
            ctx->arena_blocks.clear();
            char* p = nullptr;
            if (cudaMalloc((void**)&p, total) == cudaSuccess) ctx->arena_blocks.push_back({p, total});
            else cudaGetLastError();
            ctx->arena_block = 0;
            ctx->arena_off = 0;
            ctx->arena_used = 0;
        }
        tls_ctx = prev_ctx;
        if (prev >= 0) cudaSetDevice(prev);
    }
};
#define API_ENTER(ctx)                                     \
    if (!(ctx)) return OBVHS_ERR_INVALID_ARG;              \
    DeviceScope _scope(ctx);                               \
    (ctx)->last_error.clear()
#define ARG_CHECK(ctx, cond, msg)                          \
    do {                                                   \
        if (!(cond)) {                                     \
            OBVHS_SET_ERR(ctx, "invalid argument: %s", msg); \
            return OBVHS_ERR_INVALID_ARG;                  \
        }                                                  \
    } while (0)
}  // namespace

static void context_teardown(ObvhsContext* ctx) {
    int prev = -1;
    cudaGetDevice(&prev);
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    comm_destroy(ctx);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    if (ctx->copy_in) cudaStreamDestroy(ctx->copy_in);
    if (ctx->copy_out) cudaStreamDestroy(ctx->copy_out);
    if (ctx->compute_alt) cudaStreamDestroy(ctx->compute_alt);
    for (cudaEvent_t e : ctx->event_pool) cudaEventDestroy(e);
    for (auto& b : ctx->arena_blocks) cudaFree(b.p);
    for (auto& b : ctx->result_cache) cudaFree(b.p);
    for (auto& kv : ctx->result_live) cudaFree(kv.first);
    if (ctx->owns_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    if (prev >= 0 && prev != ctx->device) cudaSetDevice(prev);
    delete ctx;
}

namespace {
// in/out caller arrays -> arena arrays -> splitter -> back (Vec growth becomes "count_out may exceed capacity")
template <class Run>
int split_in_out(ObvhsContext* ctx, ObvhsAabb* aabbs, uint32_t* indices, size_t n, size_t capacity, size_t* count_out, Run run) {
    SplitArrays a;
    a.cap = std::max(capacity, n) + 1;
    a.len = n;
    a.aabbs = static_cast<ObvhsAabb*>(obvhs_arena_alloc(a.cap * sizeof(ObvhsAabb)));
    a.indices = static_cast<u32*>(obvhs_arena_alloc(a.cap * 4));
    if (!a.aabbs || !a.indices) {
        OBVHS_SET_ERR(ctx, "split_aabbs: out of device memory");
        return OBVHS_ERR_CUDA;
    }
    if (n) {
        CU_TRY(ctx, cudaMemcpyAsync(a.aabbs, aabbs, n * sizeof(ObvhsAabb), cudaMemcpyDefault, ctx->stream));
        CU_TRY(ctx, cudaMemcpyAsync(a.indices, indices, n * 4, cudaMemcpyDefault, ctx->stream));
    }
    ST_TRY(run(a));
    *count_out = a.len;
    if (a.len > capacity) {
        OBVHS_SET_ERR(ctx, "split_aabbs: %zu entries after splitting, capacity %zu (call again with more room)", a.len, capacity);
        return OBVHS_ERR_CAPACITY;
    }
    if (a.len > n) {  // entries [0, n) may have been shrunk in place, [n, len) are the appended right halves
        CU_TRY(ctx, cudaMemcpyAsync(aabbs, a.aabbs, a.len * sizeof(ObvhsAabb), cudaMemcpyDefault, ctx->stream));
        CU_TRY(ctx, cudaMemcpyAsync(indices, a.indices, a.len * 4, cudaMemcpyDefault, ctx->stream));
    }
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return OBVHS_OK;
}

// First stage of both triangle builders: PlocBuilder::build over the triangles, or -- with params->pre_split -- over the
// split AABBs with their index map (cwbvh/builder.rs:27-71 == bvh2/builder.rs:24-68). Records ctx->ev0 where the
// reference starts its core_build_time clock.
int ploc_from_tris(ObvhsContext* ctx, const ObvhsTriangle* d_tris, size_t n, const ObvhsBuildParams* params, ObvhsBvh2** out) {
    if (!params->pre_split) {
        CU_TRY(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
        TraceScope ts(ctx, "build_ploc");
        return ploc_build_device(ctx, nullptr, d_tris, nullptr, n, params->ploc_search_distance, params->sort_precision,
                                 (size_t)params->search_depth_threshold, out, nullptr);
    }
    SplitArrays a;
    ST_TRY(presplit_tris_device(ctx, d_tris, n, a, nullptr, ctx->ev0));
    TraceScope ts(ctx, "build_ploc");
    ST_TRY(ploc_build_device(ctx, a.aabbs, nullptr, a.indices, a.len, params->ploc_search_distance, params->sort_precision,
                             (size_t)params->search_depth_threshold, out, nullptr));
    (*out)->uses_spatial_splits = true;  // cwbvh/builder.rs:72
    return OBVHS_OK;
}
}  // namespace

extern "C" {

int obvhs_cuda_create(int device, void* stream, ObvhsContext** out) {
    if (!out) return OBVHS_ERR_INVALID_ARG;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) {
        cudaGetLastError();
        return OBVHS_ERR_CUDA;  // no CPU fallback: without a CUDA device there is no context
    }
    if (cudaSetDevice(device) != cudaSuccess) return OBVHS_ERR_CUDA;
    ObvhsContext* ctx = new ObvhsContext();
    ctx->device = device;
    const char* tr = getenv("OBVHS_TRACE");
    ctx->trace = tr && tr[0] == '1';
    if (const char* tm = getenv("OBVHS_TRAVERSE")) obvhs_cuda_set_option(ctx, "traverse", tm);
    if (const char* tv = getenv("OBVHS_TRAVERSE_VARIANT")) obvhs_cuda_set_option(ctx, "traverse_variant", tv);
    if (stream) {
        ctx->stream = (cudaStream_t)stream;
    } else {
        if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
            delete ctx;
            return OBVHS_ERR_CUDA;
        }
        ctx->owns_stream = true;
    }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) ctx->sm_count = prop.multiProcessorCount;
    if (cudaMallocHost(&ctx->pinned, 4096) != cudaSuccess || cudaEventCreate(&ctx->ev0) != cudaSuccess ||
        cudaEventCreate(&ctx->ev1) != cudaSuccess) {
        obvhs_cuda_destroy(ctx);
        return OBVHS_ERR_CUDA;
    }
    *out = ctx;
    return OBVHS_OK;
}

void obvhs_cuda_destroy(ObvhsContext* ctx) {
    if (!ctx) return;
    int prev = -1;
    cudaGetDevice(&prev);
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    obvhs_context_release(ctx);  // trees built on this context keep it alive until they are freed
    if (prev >= 0) cudaSetDevice(prev);
}

const char* obvhs_cuda_last_error(const ObvhsContext* ctx) { return ctx ? ctx->last_error.c_str() : "null context"; }

int obvhs_cuda_synchronize(ObvhsContext* ctx) {
    API_ENTER(ctx);
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return OBVHS_OK;
}

uint64_t obvhs_cuda_launch_count(const ObvhsContext* ctx) { return ctx ? ctx->launches : 0; }

int obvhs_cuda_set_option(ObvhsContext* ctx, const char* key, const char* value) {
    if (!ctx || !key || !value) return OBVHS_ERR_INVALID_ARG;
    if (strcmp(key, "traverse") == 0) {
        if (strcmp(value, "auto") == 0) ctx->traverse_mode = 2;
        else if (strcmp(value, "static") == 0) ctx->traverse_mode = 0;
        else if (strncmp(value, "persistent", 10) == 0 && (value[10] == 0 || value[10] == ':')) {
            ctx->traverse_mode = 1;
            if (value[10] == ':') {
                int refill = atoi(value + 11);
                if (refill < 1 || refill > 32) {
                    OBVHS_SET_ERR(ctx, "traverse refill threshold must be in 1..32");
                    return OBVHS_ERR_INVALID_ARG;
                }
                ctx->traverse_refill = refill;
                if (const char* c2 = strchr(value + 11, ':')) ctx->traverse_chunk = atoi(c2 + 1) < 32 ? 32 : atoi(c2 + 1);
            }
        } else {
            OBVHS_SET_ERR(ctx, "traverse must be auto, static or persistent[:refill[:chunk]]");
            return OBVHS_ERR_INVALID_ARG;
        }
        return OBVHS_OK;
    }
    if (strcmp(key, "traverse_variant") == 0) {  // persistent-kernel variant (traverse.cu: launch_persistent_t)
        const int id = atoi(value);
        if (id < 0 || id > 4) {
            OBVHS_SET_ERR(ctx, "traverse_variant id must be in 0..4");
            return OBVHS_ERR_INVALID_ARG;
        }
        ctx->traverse_variant = id;
        return OBVHS_OK;
    }
    if (strcmp(key, "host_slice") == 0) {  // rays per pipelined slice of a host batch; 0 = sized for the kernel in use
        long v = atol(value);
        if (v < 0) {
            OBVHS_SET_ERR(ctx, "host_slice must be >= 0");
            return OBVHS_ERR_INVALID_ARG;
        }
        ctx->host_slice = (size_t)v;
        return OBVHS_OK;
    }
    if (strcmp(key, "trace") == 0) {
        ctx->trace = value[0] == '1';
        return OBVHS_OK;
    }
    OBVHS_SET_ERR(ctx, "unknown option %s", key);
    return OBVHS_ERR_INVALID_ARG;
}

int obvhs_cuda_build_params_preset(const char* name, ObvhsBuildParams* out) {
    if (!name || !out) return OBVHS_ERR_INVALID_ARG;
    struct P {
        const char* n;
        ObvhsBuildParams p;
    };
    // src/lib.rs:233-305
    static const P presets[] = {
        {"fastest_build", {0, 1, 0, 0.0f, 0.0f, 64, 1, 1.0f}},
        {"very_fast_build", {0, 1, 0, 0.01f, 0.0f, 64, 8, 3.0f}},
        {"fast_build", {0, 6, 2, 0.02f, 0.0f, 64, 8, 3.0f}},
        {"medium_build", {0, 14, 3, 0.05f, 2.0f, 64, 8, 3.0f}},
        {"slow_build", {1, 24, 2, 0.2f, 2.0f, 128, 8, 3.0f}},
        {"very_slow_build", {1, 14, 1, 1.0f, 1.0f, 128, 8, 3.0f}},
    };
    for (const P& p : presets)
        if (strcmp(p.n, name) == 0) {
            *out = p.p;
            return OBVHS_OK;
        }
    return OBVHS_ERR_INVALID_ARG;
}

// ---- PLOC ---------------------------------------------------------------------------------------------------
int obvhs_cuda_morton_sort(ObvhsContext* ctx, const ObvhsAabb* aabbs, size_t n, uint32_t sort_precision, uint64_t* codes_lo,
                           uint64_t* codes_hi, uint32_t* order, ObvhsAabb* total_aabb) {
    API_ENTER(ctx);
    ARG_CHECK(ctx, n == 0 || aabbs, "aabbs is null");
    DevBuf<ObvhsAabb> st_aabbs, d_total;
    DevBuf<u64> d_lo, d_hi;
    DevBuf<u32> d_order;
    const ObvhsAabb* d_aabbs = nullptr;
    ST_TRY(stage_in(ctx, aabbs, n, st_aabbs, &d_aabbs));
    PlocMortonOut probe;
    if (codes_lo) { CU_TRY(ctx, d_lo.alloc(n, ctx->stream)); probe.codes_lo = d_lo.p; }
    if (codes_hi) { CU_TRY(ctx, d_hi.alloc(n, ctx->stream)); probe.codes_hi = d_hi.p; }
    if (order) { CU_TRY(ctx, d_order.alloc(n, ctx->stream)); probe.order = d_order.p; }
    CU_TRY(ctx, d_total.alloc(1, ctx->stream));
    CU_TRY(ctx, cudaMemsetAsync(d_total.p, 0, sizeof(ObvhsAabb), ctx->stream));
    probe.total = d_total.p;
    ObvhsBvh2* bvh = nullptr;
    // search distance 1 keeps the (unused) PLOC iterations cheap; the probe outputs are captured before them
    ST_TRY(ploc_build_device(ctx, d_aabbs, nullptr, nullptr, n, 1, sort_precision, 0, &bvh, &probe));
    obvhs_cuda_bvh2_free(bvh);
    ST_TRY(copy_out(ctx, codes_lo, (const u64*)d_lo.p, n));
    ST_TRY(copy_out(ctx, codes_hi, (const u64*)d_hi.p, n));
    ST_TRY(copy_out(ctx, order, (const u32*)d_order.p, n));
    ST_TRY(copy_out(ctx, total_aabb, (const ObvhsAabb*)d_total.p, (size_t)1));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return OBVHS_OK;
}

int obvhs_cuda_ploc_build(ObvhsContext* ctx, const ObvhsAabb* aabbs, const uint32_t* indices, size_t n, uint32_t search_distance,
                          uint32_t sort_precision, size_t search_depth_threshold, ObvhsBvh2** out) {
    API_ENTER(ctx);
    ARG_CHECK(ctx, out, "out is null");
    ARG_CHECK(ctx, n == 0 || aabbs, "aabbs is null");
    DevBuf<ObvhsAabb> st_aabbs;
    DevBuf<u32> st_idx;
    const ObvhsAabb* d_aabbs = nullptr;
    const u32* d_idx = nullptr;
    ST_TRY(stage_in(ctx, aabbs, n, st_aabbs, &d_aabbs));
    ST_TRY(stage_in(ctx, indices, n, st_idx, &d_idx));
    return ploc_build_device(ctx, d_aabbs, nullptr, d_idx, n, search_distance, sort_precision, search_depth_threshold, out, nullptr);
}

int obvhs_cuda_ploc_build_tris(ObvhsContext* ctx, const ObvhsTriangle* tris, size_t n, uint32_t search_distance, uint32_t sort_precision,
                               size_t search_depth_threshold, ObvhsBvh2** out) {
    API_ENTER(ctx);
    ARG_CHECK(ctx, out, "out is null");
    ARG_CHECK(ctx, n == 0 || tris, "tris is null");
    DevBuf<ObvhsTriangle> st_tris;
    const ObvhsTriangle* d_tris = nullptr;
    ST_TRY(stage_in(ctx, tris, n, st_tris, &d_tris));
    return ploc_build_device(ctx, nullptr, d_tris, nullptr, n, search_distance, sort_precision, search_depth_threshold, out, nullptr);
}

// ---- Bvh2 ---------------------------------------------------------------------------------------------------
void obvhs_cuda_bvh2_free(ObvhsBvh2* bvh) {
    if (!bvh) return;
    ObvhsContext* ctx = bvh->owner;
    obvhs_result_free(ctx, bvh->nodes);
    obvhs_result_free(ctx, bvh->primitive_indices);
    obvhs_result_free(ctx, bvh->parents);
    obvhs_result_free(ctx, bvh->bvh_tris);
    delete bvh;
    obvhs_context_release(ctx);
}
size_t obvhs_cuda_bvh2_node_count(const ObvhsBvh2* bvh) { return bvh ? bvh->node_count : 0; }
size_t obvhs_cuda_bvh2_prim_count(const ObvhsBvh2* bvh) { return bvh ? bvh->prim_count : 0; }
size_t obvhs_cuda_bvh2_max_depth(const ObvhsBvh2* bvh) { return bvh ? bvh->max_depth : 0; }
size_t obvhs_cuda_bvh2_ploc_iterations(const ObvhsBvh2* bvh) { return bvh ? bvh->ploc_iterations : 0; }
int obvhs_cuda_bvh2_children_ordered_after_parents(const ObvhsBvh2* bvh) { return bvh && bvh->children_are_ordered_after_parents ? 1 : 0; }

int obvhs_cuda_bvh2_download(ObvhsContext* ctx, const ObvhsBvh2* bvh, ObvhsBvh2Node* nodes, uint32_t* primitive_indices, uint32_t* parents) {
    API_ENTER(ctx);
    ARG_CHECK(ctx, bvh, "bvh is null");
    if (nodes && bvh->node_count) {
        if (obvhs_is_device_ptr(nodes)) {
            ST_TRY(bvh2_expand_nodes_device(ctx, bvh->nodes, bvh->node_count, nodes));
        } else {
            DevBuf<ObvhsBvh2Node> tmp;
            CU_TRY(ctx, tmp.alloc(bvh->node_count, ctx->stream));
            ST_TRY(bvh2_expand_nodes_device(ctx, bvh->nodes, bvh->node_count, tmp.p));
            ST_TRY(copy_out(ctx, nodes, (const ObvhsBvh2Node*)tmp.p, bvh->node_count));
        }
    }
    ST_TRY(copy_out(ctx, primitive_indices, (const u32*)bvh->primitive_indices, bvh->prim_count));
    if (parents) {
        ARG_CHECK(ctx, bvh->parents || bvh->node_count == 0, "parents requested but not computed (Bvh2::parents is None)");
        ST_TRY(copy_out(ctx, parents, (const u32*)bvh->parents, bvh->node_count));
    }
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return OBVHS_OK;
}

int obvhs_cuda_bvh2_upload(ObvhsContext* ctx, const ObvhsBvh2Node* nodes, size_t node_count, const uint32_t* primitive_indices,
                           size_t prim_count, size_t max_depth, int children_ordered_after_parents, ObvhsBvh2** out) {
    API_ENTER(ctx);
    ARG_CHECK(ctx, out, "out is null");
    ARG_CHECK(ctx, node_count == 0 || nodes, "nodes is null");
    ARG_CHECK(ctx, prim_count == 0 || primitive_indices, "primitive_indices is null");
    ObvhsBvh2* bvh = new ObvhsBvh2();
    bvh->device = ctx->device;
    bvh->owner = ctx;
    obvhs_context_retain(ctx);
    bvh->node_count = node_count;
    bvh->prim_count = prim_count;
    bvh->max_depth = max_depth ? max_depth : 96;
    bvh->children_are_ordered_after_parents = children_ordered_after_parents != 0;
    struct Guard {
        ObvhsBvh2* b;
        ~Guard() { if (b) obvhs_cuda_bvh2_free(b); }
    } guard{bvh};
    if (node_count) {
        DevBuf<ObvhsBvh2Node> st;
        const ObvhsBvh2Node* d_nodes = nullptr;
        ST_TRY(stage_in(ctx, nodes, node_count, st, &d_nodes));
        CU_TRY(ctx, obvhs_result_alloc(ctx, (void**)&bvh->nodes, node_count * sizeof(Node32)));
        ST_TRY(bvh2_pack_nodes_device(ctx, d_nodes, node_count, bvh->nodes));
    }
    if (prim_count) {
        CU_TRY(ctx, obvhs_result_alloc(ctx, (void**)&bvh->primitive_indices, prim_count * 4));
        CU_TRY(ctx, cudaMemcpyAsync(bvh->primitive_indices, primitive_indices, prim_count * 4, cudaMemcpyDefault, ctx->stream));
    }
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    guard.b = nullptr;
    *out = bvh;
    return OBVHS_OK;
}

int obvhs_cuda_bvh2_compute_parents(ObvhsContext* ctx, ObvhsBvh2* bvh) {
    API_ENTER(ctx);
    ARG_CHECK(ctx, bvh, "bvh is null");
    return bvh2_compute_parents_device(ctx, bvh);
}
int obvhs_cuda_bvh2_refit_all(ObvhsContext* ctx, ObvhsBvh2* bvh) {
    API_ENTER(ctx);
    ARG_CHECK(ctx, bvh, "bvh is null");
    return bvh2_refit_all_device(ctx, bvh);
}
int obvhs_cuda_bvh2_reorder_in_stack_traversal_order(ObvhsContext* ctx, ObvhsBvh2* bvh) {
    API_ENTER(ctx);
    ARG_CHECK(ctx, bvh, "bvh is null");
    // (primitive_indices is untouched, so permuted triangles attached to the handle stay valid)
    return bvh2_reorder_in_stack_traversal_order_device(ctx, bvh);
}
// PlocBuilder::full_rebuild / partial_rebuild / compute_rebuild_path_flags (src/ploc/rebuild.rs)
int obvhs_cuda_ploc_full_rebuild(ObvhsContext* ctx, ObvhsBvh2* bvh, uint32_t search_distance, uint32_t sort_precision,
                                 size_t search_depth_threshold) {
    API_ENTER(ctx);
    ARG_CHECK(ctx, bvh, "bvh is null");
    ST_TRY(ploc_full_rebuild_device(ctx, bvh, search_distance, sort_precision, search_depth_threshold));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return OBVHS_OK;
}
int obvhs_cuda_ploc_partial_rebuild(ObvhsContext* ctx, ObvhsBvh2* bvh, const uint8_t* should_remove, uint32_t search_distance,
                                    uint32_t sort_precision, size_t search_depth_threshold) {
    API_ENTER(ctx);
    ARG_CHECK(ctx, bvh && (bvh->node_count < 2 || should_remove), "null argument");
    DevBuf<u8> st;
    const u8* d_flags = nullptr;
    ST_TRY(stage_in(ctx, should_remove, bvh->node_count, st, &d_flags));
    ST_TRY(ploc_partial_rebuild_device(ctx, bvh, d_flags, search_distance, sort_precision, search_depth_threshold));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return OBVHS_OK;
}
int obvhs_cuda_compute_rebuild_path_flags(ObvhsContext* ctx, const ObvhsBvh2* bvh, const uint32_t* leaves, size_t n_leaves, uint8_t* flags) {
    API_ENTER(ctx);
    ARG_CHECK(ctx, bvh && (n_leaves == 0 || leaves) && (bvh->node_count < 2 || flags), "null argument");
    if (bvh->node_count < 2) return OBVHS_OK;
    DevBuf<u32> st;
    DevBuf<u8> d_out;
    const u32* d_leaves = nullptr;
    ST_TRY(stage_in(ctx, leaves, n_leaves, st, &d_leaves));
    u8* d_flags = flags;
    const bool out_dev = obvhs_is_device_ptr(flags);
    if (!out_dev) {
        CU_TRY(ctx, d_out.alloc(bvh->node_count, ctx->stream));
        d_flags = d_out.p;
    }
    ST_TRY(ploc_compute_rebuild_path_flags_device(ctx, bvh, d_leaves, n_leaves, d_flags));
    if (!out_dev) ST_TRY(copy_out(ctx, flags, (const u8*)d_flags, bvh->node_count));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return OBVHS_OK;
}
// Bvh2Node::set_aabb on a list of nodes (examples/physics.rs:446): the leaves get their new boxes before a partial rebuild
int obvhs_cuda_bvh2_set_node_aabbs(ObvhsContext* ctx, ObvhsBvh2* bvh, const uint32_t* node_ids, const ObvhsAabb* aabbs, size_t n) {
    API_ENTER(ctx);
    ARG_CHECK(ctx, bvh && (n == 0 || (node_ids && aabbs)), "null argument");
    DevBuf<u32> st_ids;
    DevBuf<ObvhsAabb> st_aabbs;
    const u32* d_ids = nullptr;
    const ObvhsAabb* d_aabbs = nullptr;
    ST_TRY(stage_in(ctx, node_ids, n, st_ids, &d_ids));
    ST_TRY(stage_in(ctx, aabbs, n, st_aabbs, &d_aabbs));
    ST_TRY(bvh2_set_node_aabbs_device(ctx, bvh, d_ids, d_aabbs, n));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return OBVHS_OK;
}

int obvhs_cuda_bvh2_set_leaf_aabbs(ObvhsContext* ctx, ObvhsBvh2* bvh, const ObvhsAabb* prim_aabbs, size_t n) {
    API_ENTER(ctx);
    ARG_CHECK(ctx, bvh && (n == 0 || prim_aabbs), "null argument");
    ARG_CHECK(ctx, n >= bvh->prim_count, "fewer AABBs than primitives");
    DevBuf<ObvhsAabb> st;
    const ObvhsAabb* d = nullptr;
    ST_TRY(stage_in(ctx, prim_aabbs, n, st, &d));
    ST_TRY(bvh2_set_leaf_aabbs_device(ctx, bvh, d));
    return bvh2_refit_all_device(ctx, bvh);
}

int obvhs_cuda_reinsertion_run(ObvhsContext* ctx, ObvhsBvh2* bvh, float batch_size_ratio, const float* ratio_sequence, size_t n_sequence,
                               uint64_t* applied_out) {
    API_ENTER(ctx);
    ARG_CHECK(ctx, bvh, "bvh is null");
    ARG_CHECK(ctx, !ratio_sequence || !obvhs_is_device_ptr(ratio_sequence), "ratio_sequence must be host memory");
    u64 applied = 0;
    int rc = reinsertion_run_device(ctx, bvh, batch_size_ratio, ratio_sequence, n_sequence, &applied);
    if (applied_out) *applied_out = applied;
    return rc;
}

int obvhs_cuda_reinsertion_run_with_candidates(ObvhsContext* ctx, ObvhsBvh2* bvh, const uint32_t* node_ids, size_t n, uint32_t iterations,
                                               uint64_t* applied_out) {
    API_ENTER(ctx);
    ARG_CHECK(ctx, bvh, "bvh is null");
    ARG_CHECK(ctx, n == 0 || node_ids, "node_ids is null");
    if (applied_out) *applied_out = 0;
    if (n == 0) return OBVHS_OK;
    // find_reinsertion asserts node_id != 0 and indexes nodes[node_id] (reinsertion.rs:233-240): check instead of panicking
    std::vector<u32> host_ids;
    const u32* h_ids = node_ids;
    if (obvhs_is_device_ptr(node_ids)) {
        host_ids.resize(n);
        CU_TRY(ctx, cudaMemcpyAsync(host_ids.data(), node_ids, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
        CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        h_ids = host_ids.data();
    }
    if (bvh->node_count > 1)
        for (size_t i = 0; i < n; i++)
            if (h_ids[i] == 0 || h_ids[i] >= bvh->node_count) {
                OBVHS_SET_ERR(ctx, "reinsertion candidate %zu = node %u is the root or out of range (%zu nodes)", i, h_ids[i], bvh->node_count);
                return OBVHS_ERR_INVALID_ARG;
            }
    DevBuf<u32> st;
    const u32* d_ids = nullptr;
    ST_TRY(stage_in(ctx, node_ids, n, st, &d_ids));
    u64 applied = 0;
    int rc = reinsertion_run_candidates_device(ctx, bvh, d_ids, n, iterations, &applied);
    if (applied_out) *applied_out = applied;
    return rc;
}

// ---- spatial pre-splits (src/splits.rs) ------------------------------------------------------------------------

int obvhs_cuda_split_aabbs_precise(ObvhsContext* ctx, ObvhsAabb* aabbs, uint32_t* indices, size_t n, size_t capacity,
                                   const ObvhsTriangle* tris, size_t n_tris, float area_thresh_low, float area_thresh_high,
                                   float split_factor_low, float split_factor_high, uint32_t max_iterations, uint32_t split_tests,
                                   size_t* count_out) {
    API_ENTER(ctx);
    ARG_CHECK(ctx, count_out && (n == 0 || (aabbs && indices && tris)) && capacity >= n, "null argument or capacity < n");
    DevBuf<ObvhsTriangle> st_tris;
    const ObvhsTriangle* d_tris = nullptr;
    ST_TRY(stage_in(ctx, tris, n_tris, st_tris, &d_tris));
    return split_in_out(ctx, aabbs, indices, n, capacity, count_out, [&](SplitArrays& a) {
        return split_aabbs_precise_device(ctx, a, d_tris, area_thresh_low, area_thresh_high, split_factor_low, split_factor_high,
                                          max_iterations, split_tests);
    });
}

int obvhs_cuda_split_aabbs_preset(ObvhsContext* ctx, ObvhsAabb* aabbs, uint32_t* indices, size_t n, size_t capacity,
                                  const ObvhsTriangle* tris, size_t n_tris, float avg_half_area, float largest_half_area,
                                  size_t* count_out) {
    // splits.rs:23-33 (host f32 arithmetic; this translation unit is compiled without contraction)
    volatile float lo = avg_half_area * 3.0f, hi_a = avg_half_area * 4.0f, t0 = avg_half_area * 0.9f, t1 = largest_half_area * 0.1f;
    volatile float hi_b = t0 + t1;
    const float hi = fmaxf(hi_a, hi_b);
    return obvhs_cuda_split_aabbs_precise(ctx, aabbs, indices, n, capacity, tris, n_tris, lo, hi, 1.8f, 1.6f, 12, 12, count_out);
}

int obvhs_cuda_presplit_tris(ObvhsContext* ctx, const ObvhsTriangle* tris, size_t n, ObvhsAabb* aabbs_out, uint32_t* indices_out,
                             size_t capacity, size_t* count_out, float* avg_half_area, float* largest_half_area) {
    API_ENTER(ctx);
    ARG_CHECK(ctx, count_out && (n == 0 || tris), "null argument");
    DevBuf<ObvhsTriangle> st_tris;
    const ObvhsTriangle* d_tris = nullptr;
    ST_TRY(stage_in(ctx, tris, n, st_tris, &d_tris));
    SplitArrays a;
    float al[2] = {0.f, 0.f};
    ST_TRY(presplit_tris_device(ctx, d_tris, n, a, al, nullptr));
    if (avg_half_area) *avg_half_area = al[0];
    if (largest_half_area) *largest_half_area = al[1];
    *count_out = a.len;
    if (aabbs_out || indices_out) {
        if (a.len > capacity) {
            OBVHS_SET_ERR(ctx, "presplit_tris: %zu entries, capacity %zu (call again with more room)", a.len, capacity);
            return OBVHS_ERR_CAPACITY;
        }
        ST_TRY(copy_out(ctx, aabbs_out, (const ObvhsAabb*)a.aabbs, a.len));
        ST_TRY(copy_out(ctx, indices_out, (const u32*)a.indices, a.len));
    }
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return OBVHS_OK;
}

// ---- CwBvh --------------------------------------------------------------------------------------------------
int obvhs_cuda_bvh2_to_cwbvh(ObvhsContext* ctx, const ObvhsBvh2* bvh, uint32_t max_prims_per_leaf, int order_children,
                             int include_exact_node_aabbs, ObvhsCwBvh** out) {
    API_ENTER(ctx);
    ARG_CHECK(ctx, bvh && out, "null argument");
    return bvh2_to_cwbvh_device(ctx, bvh, max_prims_per_leaf, order_children != 0, out, include_exact_node_aabbs != 0);
}

int obvhs_cuda_build_cwbvh_from_tris(ObvhsContext* ctx, const ObvhsTriangle* tris, size_t n, const ObvhsBuildParams* params,
                                     double* core_build_seconds, ObvhsCwBvh** out) {
    API_ENTER(ctx);
    ARG_CHECK(ctx, params && out, "null argument");
    ARG_CHECK(ctx, n == 0 || tris, "tris is null");
    DevBuf<ObvhsTriangle> st_tris;
    const ObvhsTriangle* d_tris = nullptr;
    ST_TRY(stage_in(ctx, tris, n, st_tris, &d_tris));
    // core_build_time brackets [splits ->] PLOC -> reinsertion -> collapse (cwbvh/builder.rs:45,62-76), measured on the device
    ObvhsBvh2* bvh2 = nullptr;
    ST_TRY(ploc_from_tris(ctx, d_tris, n, params, &bvh2));
    struct Guard {
        ObvhsBvh2* b;
        ~Guard() { obvhs_cuda_bvh2_free(b); }
    } guard{bvh2};
    ST_TRY(reinsertion_run_device(ctx, bvh2, params->reinsertion_batch_ratio, nullptr, 0, nullptr));
    u32 mp = params->max_prims_per_leaf < 1 ? 1 : (params->max_prims_per_leaf > 3 ? 3 : params->max_prims_per_leaf);  // builder.rs:74
    ObvhsCwBvh* cw = nullptr;
    {
        TraceScope ts(ctx, "bvh2_to_cwbvh");
        ST_TRY(bvh2_to_cwbvh_device(ctx, bvh2, mp, true, &cw));
    }
    struct CwGuard {  // every early return below frees the tree
        ObvhsCwBvh* b;
        ~CwGuard() { if (b) obvhs_cuda_cwbvh_free(b); }
    } cw_guard{cw};
    CU_TRY(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
    ST_TRY(cwbvh_permute_tris_device(ctx, cw, d_tris, n));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    if (core_build_seconds) {
        float ms = 0.f;
        CU_TRY(ctx, cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
        *core_build_seconds += (double)ms * 1e-3;
    }
    cw_guard.b = nullptr;
    *out = cw;
    return OBVHS_OK;
}

// build_cwbvh<T: Boundable> (cwbvh/builder.rs:98-123): PLOC over the AABBs -> reinsertion -> collapse; pre_split is ignored
int obvhs_cuda_build_cwbvh(ObvhsContext* ctx, const ObvhsAabb* aabbs, size_t n, const ObvhsBuildParams* params, double* core_build_seconds,
                           ObvhsCwBvh** out) {
    API_ENTER(ctx);
    ARG_CHECK(ctx, params && out, "null argument");
    ARG_CHECK(ctx, n == 0 || aabbs, "aabbs is null");
    DevBuf<ObvhsAabb> st;
    const ObvhsAabb* d_aabbs = nullptr;
    ST_TRY(stage_in(ctx, aabbs, n, st, &d_aabbs));
    CU_TRY(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
    ObvhsBvh2* bvh2 = nullptr;
    {
        TraceScope ts(ctx, "build_ploc");
        ST_TRY(ploc_build_device(ctx, d_aabbs, nullptr, nullptr, n, params->ploc_search_distance, params->sort_precision,
                                 (size_t)params->search_depth_threshold, &bvh2, nullptr));
    }
    struct Guard {
        ObvhsBvh2* b;
        ~Guard() { obvhs_cuda_bvh2_free(b); }
    } guard{bvh2};
    ST_TRY(reinsertion_run_device(ctx, bvh2, params->reinsertion_batch_ratio, nullptr, 0, nullptr));
    u32 mp = params->max_prims_per_leaf < 1 ? 1 : (params->max_prims_per_leaf > 3 ? 3 : params->max_prims_per_leaf);  // builder.rs:112
    ObvhsCwBvh* cw = nullptr;
    ST_TRY(bvh2_to_cwbvh_device(ctx, bvh2, mp, true, &cw));
    CU_TRY(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
        obvhs_cuda_cwbvh_free(cw);
        OBVHS_SET_ERR(ctx, "build_cwbvh: stream synchronisation failed");
        return OBVHS_ERR_CUDA;
    }
    if (core_build_seconds) {
        float ms = 0.f;
        CU_TRY(ctx, cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
        *core_build_seconds += (double)ms * 1e-3;
    }
    *out = cw;
    return OBVHS_OK;
}

void obvhs_cuda_cwbvh_free(ObvhsCwBvh* bvh) {
    if (!bvh) return;
    ObvhsContext* ctx = bvh->owner;
    obvhs_result_free(ctx, bvh->nodes);
    obvhs_result_free(ctx, bvh->primitive_indices);
    obvhs_result_free(ctx, bvh->bvh_tris);
    obvhs_result_free(ctx, bvh->exact_node_aabbs);
    delete bvh;
    obvhs_context_release(ctx);
}
// CwBvh::exact_node_aabbs (cwbvh/mod.rs:47): *count = number of entries (0 when the tree was converted without them)
int obvhs_cuda_cwbvh_exact_node_aabbs(ObvhsContext* ctx, const ObvhsCwBvh* bvh, ObvhsAabb* out, size_t capacity, size_t* count) {
    API_ENTER(ctx);
    ARG_CHECK(ctx, bvh && count, "null argument");
    *count = bvh->exact_node_aabbs ? bvh->exact_count : 0;
    if (!out || *count == 0) return OBVHS_OK;
    if (capacity < *count) {
        OBVHS_SET_ERR(ctx, "exact_node_aabbs: %zu entries, capacity %zu", *count, capacity);
        return OBVHS_ERR_CAPACITY;
    }
    ST_TRY(copy_out(ctx, out, (const ObvhsAabb*)bvh->exact_node_aabbs, *count));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return OBVHS_OK;
}
int obvhs_cuda_cwbvh_compute_parents(ObvhsContext* ctx, const ObvhsCwBvh* bvh, uint32_t* parents) {
    API_ENTER(ctx);
    ARG_CHECK(ctx, bvh && (parents || bvh->node_count == 0), "null argument");
    if (bvh->node_count == 0) return OBVHS_OK;
    DevBuf<u32> st;
    u32* d = parents;
    if (!obvhs_is_device_ptr(parents)) {
        CU_TRY(ctx, st.alloc(bvh->node_count, ctx->stream));
        d = st.p;
    }
    ST_TRY(cwbvh_compute_parents_device(ctx, bvh, d));
    if (d != parents) ST_TRY(copy_out(ctx, parents, (const u32*)d, bvh->node_count));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return OBVHS_OK;
}
// CwBvh::order_children (src/cwbvh/mod.rs:520-524); the primitives arrive as their AABBs (Boundable::aabb)
int obvhs_cuda_cwbvh_order_children(ObvhsContext* ctx, ObvhsCwBvh* bvh, const ObvhsAabb* prim_aabbs, size_t n, int direct_layout) {
    API_ENTER(ctx);
    ARG_CHECK(ctx, bvh && (n == 0 || prim_aabbs), "null argument");
    if (bvh->node_count == 0) return OBVHS_OK;
    DevBuf<ObvhsAabb> st;
    const ObvhsAabb* d = nullptr;
    ST_TRY(stage_in(ctx, prim_aabbs, n, st, &d));
    return cwbvh_order_children_device(ctx, bvh, d, n, direct_layout != 0);
}
int obvhs_cuda_cwbvh_uses_spatial_splits(const ObvhsCwBvh* bvh) { return bvh && bvh->uses_spatial_splits; }
void obvhs_cuda_cwbvh_set_uses_spatial_splits(ObvhsCwBvh* bvh, int v) { if (bvh) bvh->uses_spatial_splits = v != 0; }
int obvhs_cuda_bvh2_uses_spatial_splits(const ObvhsBvh2* bvh) { return bvh && bvh->uses_spatial_splits; }
void obvhs_cuda_bvh2_set_uses_spatial_splits(ObvhsBvh2* bvh, int v) { if (bvh) bvh->uses_spatial_splits = v != 0; }
size_t obvhs_cuda_cwbvh_node_count(const ObvhsCwBvh* bvh) { return bvh ? bvh->node_count : 0; }
size_t obvhs_cuda_cwbvh_prim_count(const ObvhsCwBvh* bvh) { return bvh ? bvh->prim_count : 0; }

int obvhs_cuda_cwbvh_download(ObvhsContext* ctx, const ObvhsCwBvh* bvh, ObvhsCwBvhNode* nodes, uint32_t* primitive_indices,
                              ObvhsAabb* total_aabb) {
    API_ENTER(ctx);
    ARG_CHECK(ctx, bvh, "bvh is null");
    ST_TRY(copy_out(ctx, nodes, (const ObvhsCwBvhNode*)bvh->nodes, bvh->node_count));
    ST_TRY(copy_out(ctx, primitive_indices, (const u32*)bvh->primitive_indices, bvh->prim_count));
    if (total_aabb) {
        if (obvhs_is_device_ptr(total_aabb)) CU_TRY(ctx, cudaMemcpyAsync(total_aabb, &bvh->total_aabb, 32, cudaMemcpyHostToDevice, ctx->stream));
        else *total_aabb = bvh->total_aabb;
    }
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return OBVHS_OK;
}

int obvhs_cuda_cwbvh_alloc(ObvhsContext* ctx, size_t node_count, size_t prim_count, int with_triangles, const ObvhsAabb* total_aabb,
                           ObvhsCwBvh** out) {
    API_ENTER(ctx);
    ARG_CHECK(ctx, out, "out is null");
    ObvhsCwBvh* cw = new ObvhsCwBvh();
    cw->device = ctx->device;
    cw->owner = ctx;
    obvhs_context_retain(ctx);
    cw->node_count = node_count;
    cw->prim_count = prim_count;
    if (total_aabb) cw->total_aabb = *total_aabb;
    struct Guard {
        ObvhsCwBvh* b;
        ~Guard() { if (b) obvhs_cuda_cwbvh_free(b); }
    } guard{cw};
    if (node_count) CU_TRY(ctx, obvhs_result_alloc(ctx, (void**)&cw->nodes, node_count * sizeof(ObvhsCwBvhNode)));
    if (prim_count) CU_TRY(ctx, obvhs_result_alloc(ctx, (void**)&cw->primitive_indices, prim_count * 4));
    if (prim_count && with_triangles) CU_TRY(ctx, obvhs_result_alloc(ctx, (void**)&cw->bvh_tris, prim_count * OBVHS_RT_TRIANGLE_BYTES));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    guard.b = nullptr;
    *out = cw;
    return OBVHS_OK;
}

int obvhs_cuda_cwbvh_upload(ObvhsContext* ctx, const ObvhsCwBvhNode* nodes, size_t node_count, const uint32_t* primitive_indices,
                            size_t prim_count, const ObvhsAabb* total_aabb, ObvhsCwBvh** out) {
    API_ENTER(ctx);
    ARG_CHECK(ctx, out, "out is null");
    ARG_CHECK(ctx, node_count == 0 || nodes, "nodes is null");
    ARG_CHECK(ctx, prim_count == 0 || primitive_indices, "primitive_indices is null");
    ObvhsCwBvh* cw = nullptr;
    ObvhsAabb total = {};
    if (total_aabb) {
        if (obvhs_is_device_ptr(total_aabb)) CU_TRY(ctx, cudaMemcpy(&total, total_aabb, 32, cudaMemcpyDeviceToHost));
        else total = *total_aabb;
    }
    ST_TRY(obvhs_cuda_cwbvh_alloc(ctx, node_count, prim_count, 0, &total, &cw));
    cudaError_t e = cudaSuccess;
    if (node_count) e = cudaMemcpyAsync(cw->nodes, nodes, node_count * sizeof(ObvhsCwBvhNode), cudaMemcpyDefault, ctx->stream);
    if (e == cudaSuccess && prim_count) e = cudaMemcpyAsync(cw->primitive_indices, primitive_indices, prim_count * 4, cudaMemcpyDefault, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
        OBVHS_SET_ERR(ctx, "cwbvh_upload: %s", cudaGetErrorString(e));
        obvhs_cuda_cwbvh_free(cw);
        return OBVHS_ERR_CUDA;
    }
    *out = cw;
    return OBVHS_OK;
}

int obvhs_cuda_cwbvh_set_triangles(ObvhsContext* ctx, ObvhsCwBvh* bvh, const ObvhsTriangle* tris, size_t n) {
    API_ENTER(ctx);
    ARG_CHECK(ctx, bvh && (n == 0 || tris), "null argument");
    DevBuf<ObvhsTriangle> st;
    const ObvhsTriangle* d = nullptr;
    ST_TRY(stage_in(ctx, tris, n, st, &d));
    ST_TRY(cwbvh_permute_tris_device(ctx, bvh, d, n));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return OBVHS_OK;
}

size_t obvhs_cuda_cwbvh_triangle_bytes(void) { return OBVHS_RT_TRIANGLE_BYTES; }
int obvhs_cuda_cwbvh_device_ptrs(const ObvhsCwBvh* bvh, void** nodes, void** primitive_indices, void** bvh_tris) {
    if (!bvh) return OBVHS_ERR_INVALID_ARG;
    if (nodes) *nodes = bvh->nodes;
    if (primitive_indices) *primitive_indices = bvh->primitive_indices;
    if (bvh_tris) *bvh_tris = bvh->bvh_tris;
    return OBVHS_OK;
}

}  // extern "C"

// Ray::new(origin, direction, min, max) (ray.rs:34-52) for a batch: the 32-byte constructor arguments become the 64-byte Ray the
// traversal kernels read. A host batch then crosses PCIe at 32 B per ray instead of 64 (the link, not the kernel, bounds a
// host-to-host traversal call), and the inverse direction is computed where it is consumed.
__device__ __forceinline__ float safe_inverse_dev(float x) {  // ray.rs:6-12; signum(+-0) = +-1, NaN stays NaN
    return fabsf(x) <= 1.1920929e-07f ? copysignf(1.0f, x) / 1.1920929e-07f : 1.0f / x;
}
__global__ void __launch_bounds__(256) ray_new_kernel(const float4* __restrict__ args, size_t n, float4* __restrict__ rays) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 o = __ldg(args + 2 * i), d = __ldg(args + 2 * i + 1);
    float4* r = rays + 4 * i;
    r[0] = make_float4(o.x, o.y, o.z, 0.0f);
    r[1] = make_float4(d.x, d.y, d.z, 0.0f);
    r[2] = make_float4(safe_inverse_dev(d.x), safe_inverse_dev(d.y), safe_inverse_dev(d.z), 0.0f);
    r[3] = make_float4(o.w, d.w, 0.0f, 0.0f);
}
static int ray_new_device(ObvhsContext* ctx, const ObvhsRayNew* d_args, size_t n, ObvhsRay* d_rays) {
    if (n == 0) return OBVHS_OK;
    ray_new_kernel<<<div_up(n, 256), 256, 0, ctx->stream>>>(reinterpret_cast<const float4*>(d_args), n, reinterpret_cast<float4*>(d_rays));
    KERNEL_CHECK(ctx);
    return OBVHS_OK;
}

// Host-resident ray batches are pipelined in chunks over three streams: H2D of chunk k+1 (copy_in), traversal of chunk k
// (the context's stream) and D2H of chunk k-1 (copy_out) run concurrently, so a batch costs about max(H2D, kernel, D2H)
// instead of their sum (PCIe is full duplex). Letting the kernel read pinned host memory directly (zero-copy) measured
// 1.5x SLOWER on B200/PCIe (425 vs 630 Mrays/s on the kitchen), small PCIe reads being latency-bound.
static int ensure_pipeline(ObvhsContext* ctx, size_t n_events) {
    if (!ctx->copy_in) CU_TRY(ctx, cudaStreamCreateWithFlags(&ctx->copy_in, cudaStreamNonBlocking));
    if (!ctx->copy_out) CU_TRY(ctx, cudaStreamCreateWithFlags(&ctx->copy_out, cudaStreamNonBlocking));
    if (!ctx->compute_alt) CU_TRY(ctx, cudaStreamCreateWithFlags(&ctx->compute_alt, cudaStreamNonBlocking));
    while (ctx->event_pool.size() < n_events) {
        cudaEvent_t e;
        CU_TRY(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        ctx->event_pool.push_back(e);
    }
    return OBVHS_OK;
}

// `launch(d_rays, count, d_out, d_counters)` enqueues the traversal of a device-resident slice on ctx->stream
struct StreamSwap {  // launches go to ctx->stream: point it at another stream for a while, whatever the exit path
    ObvhsContext* ctx;
    cudaStream_t saved;
    explicit StreamSwap(ObvhsContext* c) : ctx(c), saved(c->stream) {}
    ~StreamSwap() { ctx->stream = saved; }
};
// a failed call must not return while the side streams still use arena staging the DeviceScope is about to release
struct PipelineGuard {
    ObvhsContext* ctx;
    bool armed = false;
    ~PipelineGuard() {
        if (!armed) return;
        if (ctx->copy_in) cudaStreamSynchronize(ctx->copy_in);
        if (ctx->copy_out) cudaStreamSynchronize(ctx->copy_out);
        if (ctx->compute_alt) cudaStreamSynchronize(ctx->compute_alt);
        cudaStreamSynchronize(ctx->stream);
    }
};

// RayIn = ObvhsRay (the reference's 64-byte struct) or ObvhsRayNew (the 32-byte arguments of Ray::new; the kernels run the
// constructor themselves when they fetch a ray, traverse.cu: ray_load)
template <class RayIn, class Launch>
static int traverse_common(ObvhsContext* ctx, const void* bvh, size_t prim_count, const RayIn* rays, size_t n, void* out, size_t out_elem,
                           uint64_t* counters, Launch launch) {
    ARG_CHECK(ctx, bvh, "bvh is null");
    ARG_CHECK(ctx, n == 0 || (rays && out), "null rays/out");
    if (n == 0) return OBVHS_OK;
    DevBuf<RayIn> st_in;
    DevBuf<unsigned char> st_out;
    DevBuf<u64> st_cnt;
    const bool rays_dev = obvhs_is_device_ptr(rays);
    const bool out_dev = obvhs_is_device_ptr(out);
    void* d_out = out;
    if (!out_dev) {
        CU_TRY(ctx, st_out.alloc(n * out_elem, ctx->stream));
        d_out = st_out.p;
    }
    u64* d_cnt = nullptr;
    bool cnt_dev = false;
    if (counters) {
        cnt_dev = obvhs_is_device_ptr(counters);
        if (cnt_dev) d_cnt = counters;
        else {
            CU_TRY(ctx, st_cnt.alloc(2, ctx->stream));
            CU_TRY(ctx, cudaMemcpyAsync(st_cnt.p, counters, 16, cudaMemcpyHostToDevice, ctx->stream));
            d_cnt = st_cnt.p;
        }
    }
    bool persistent = false;
    const size_t MIN_CHUNK = traverse_host_chunk_min(ctx, prim_count, &persistent);
    PipelineGuard guard{ctx};
    if (rays_dev || n < MIN_CHUNK + MIN_CHUNK / 2) {
        const RayIn* d_in = nullptr;
        ST_TRY(stage_in(ctx, rays, n, st_in, &d_in));  // (a device pointer passes through)
        ST_TRY(launch(d_in, n, d_out, d_cnt));
        if (!out_dev) CU_TRY(ctx, cudaMemcpyAsync(out, d_out, n * out_elem, cudaMemcpyDeviceToHost, ctx->stream));
    } else {
        CU_TRY(ctx, st_in.alloc(n, ctx->stream));
        unsigned char* d_stage = reinterpret_cast<unsigned char*>(st_in.p);
        // Slice boundaries: ~6 equal slices, but never below what the kernel in use needs to run efficiently
        // (traverse_host_chunk_min). The un-overlapped head (first H2D slice) and tail (last kernel + D2H slice) shrink with the
        // slice, but many small copies in both directions at once cost more on the link than they save: 2 M kitchen rays
        // (66 MB up, 33 MB down; plain cudaMemcpyAsync of both concurrently 1.35-1.45 ms) took 1.75 ms in 16 slices, 1.61 in
        // 4-6, 1.69 in 3 and 1.84 in 2.
        // A batch bound by the link (always the case for the one-ray-per-thread kernel: 8 G rays/s against 1.7 G rays/s of PCIe)
        // ends one kernel + one D2H slice after its last H2D slice, so the slices SHRINK towards the end: 5,5,3,2,1 sixteenths.
        std::vector<size_t> cut(1, 0);
        if (!persistent && !ctx->host_slice && n >= 16 * MIN_CHUNK) {
            const size_t unit = ((n + 15) / 16 + 127) & ~(size_t)127;
            for (size_t parts : {5, 5, 3, 2})
                if (cut.back() + parts * unit < n) cut.push_back(cut.back() + parts * unit);
        } else {
            size_t chunk = (n + 5) / 6;
            if (chunk < MIN_CHUNK) chunk = MIN_CHUNK;
            if (chunk > ((size_t)1 << 21) && MIN_CHUNK <= ((size_t)1 << 21)) chunk = (size_t)1 << 21;
            size_t k = (n + chunk - 1) / chunk;
            if (n - (k - 1) * chunk < chunk / 2 && k > 1) k--;  // no runt slice at the end: spread it over the others
            chunk = ((n + k - 1) / k + 127) & ~(size_t)127;
            while (cut.back() + chunk < n) cut.push_back(cut.back() + chunk);
        }
        if (persistent && !ctx->host_slice && cut.size() >= 2) {
            // the batch ends one kernel + one D2H slice after the last H2D slice: halve that one (the two compute streams hide
            // what the smaller launch loses in efficiency)
            const size_t mid = (cut.back() + (n - cut.back()) / 2 + 127) & ~(size_t)127;
            if (mid > cut.back() && mid < n) cut.push_back(mid);
        }
        cut.push_back(n);
        const size_t n_chunks = cut.size() - 1;
        ST_TRY(ensure_pipeline(ctx, 2 * n_chunks + 2));
        cudaEvent_t* ev = ctx->event_pool.data();
        // the staging areas come from the arena, whose reuse is ordered on ctx->stream: the copy streams start after it
        CU_TRY(ctx, cudaEventRecord(ev[0], ctx->stream));
        guard.armed = true;
        CU_TRY(ctx, cudaStreamWaitEvent(ctx->copy_in, ev[0], 0));
        CU_TRY(ctx, cudaStreamWaitEvent(ctx->copy_out, ev[0], 0));
        // Persistent-kernel slices alternate between two compute streams: the CTAs of slice k+1 move in as those of slice k
        // run out of rays, so the tail of a slice (a few long rays on an otherwise idle GPU) is hidden behind the next one.
        const bool two_streams = persistent && n_chunks > 1;
        if (two_streams) CU_TRY(ctx, cudaStreamWaitEvent(ctx->compute_alt, ev[0], 0));
        StreamSwap swap(ctx);
        cudaEvent_t last_alt = nullptr;
        const unsigned char* h_rays = reinterpret_cast<const unsigned char*>(rays);
        // every H2D slice is queued before the first launch: the copy engine never waits for the host to get through the
        // launches of the previous slice
        for (size_t c = 0; c < n_chunks; c++) {
            const size_t off = cut[c], cnt = cut[c + 1] - off;
            CU_TRY(ctx, cudaMemcpyAsync(d_stage + off * sizeof(RayIn), h_rays + off * sizeof(RayIn), cnt * sizeof(RayIn), cudaMemcpyHostToDevice, ctx->copy_in));
            CU_TRY(ctx, cudaEventRecord(ev[2 + 2 * c], ctx->copy_in));
        }
        for (size_t c = 0; c < n_chunks; c++) {
            const size_t off = cut[c], cnt = cut[c + 1] - off;
            cudaEvent_t e_in = ev[2 + 2 * c], e_k = ev[3 + 2 * c];
            ctx->stream = (two_streams && (c & 1)) ? ctx->compute_alt : swap.saved;
            CU_TRY(ctx, cudaStreamWaitEvent(ctx->stream, e_in, 0));
            ST_TRY(launch(st_in.p + off, cnt, (unsigned char*)d_out + off * out_elem, d_cnt));
            if (!out_dev || ctx->stream != swap.saved) CU_TRY(ctx, cudaEventRecord(e_k, ctx->stream));
            if (ctx->stream != swap.saved) last_alt = e_k;
            if (!out_dev) {
                CU_TRY(ctx, cudaStreamWaitEvent(ctx->copy_out, e_k, 0));
                CU_TRY(ctx, cudaMemcpyAsync((unsigned char*)out + off * out_elem, (unsigned char*)d_out + off * out_elem, cnt * out_elem,
                                            cudaMemcpyDeviceToHost, ctx->copy_out));
            }
        }
        ctx->stream = swap.saved;
        if (last_alt) CU_TRY(ctx, cudaStreamWaitEvent(ctx->stream, last_alt, 0));  // join the second compute stream
        if (!out_dev) {  // join the copy-out stream back into the context's stream
            CU_TRY(ctx, cudaEventRecord(ev[1], ctx->copy_out));
            CU_TRY(ctx, cudaStreamWaitEvent(ctx->stream, ev[1], 0));
        }
    }
    if (counters && !cnt_dev) CU_TRY(ctx, cudaMemcpyAsync(counters, d_cnt, 16, cudaMemcpyDeviceToHost, ctx->stream));
    // host-visible results (or host-staged inputs whose arena slot is released on return): synchronous like the reference
    if (!out_dev || !rays_dev || (counters && !cnt_dev)) CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    guard.armed = false;
    return OBVHS_OK;
}

template <class RayIn>
static RayFormat ray_format_of(float tmin = 0.f, float tmax = 0.f, bool hit8 = false) {
    return RayFormat{std::is_same<RayIn, ObvhsRayOd>::value ? 2u : std::is_same<RayIn, ObvhsRayNew>::value ? 1u : 0u, tmin, tmax, hit8 ? 1u : 0u};
}
template <class RayIn>
static int cw_traverse(ObvhsContext* ctx, const ObvhsCwBvh* bvh, const RayIn* rays, size_t n, int mode, void* out, size_t out_elem,
                       uint64_t* counters, float tmin = 0.f, float tmax = 0.f) {
    const RayFormat fmt = ray_format_of<RayIn>(tmin, tmax, mode == 0 && out_elem == sizeof(ObvhsRayHit8));
    return traverse_common(ctx, bvh, bvh ? bvh->prim_count : 0, rays, n, out, out_elem, counters, [=](const RayIn* d_rays, size_t cnt, void* d_out, u64* d_cnt) {
        return cwbvh_traverse_device(ctx, bvh, d_rays, fmt, cnt, mode, d_out, d_cnt);
    });
}
template <class RayIn>
static int b2_traverse(ObvhsContext* ctx, const ObvhsBvh2* bvh, const RayIn* rays, size_t n, int mode, void* out, size_t out_elem,
                       uint64_t* counters, float tmin = 0.f, float tmax = 0.f) {
    const RayFormat fmt = ray_format_of<RayIn>(tmin, tmax);
    return traverse_common(ctx, bvh, bvh ? bvh->prim_count : 0, rays, n, out, out_elem, counters, [=](const RayIn* d_rays, size_t cnt, void* d_out, u64* d_cnt) {
        return bvh2_traverse_device(ctx, bvh, d_rays, fmt, cnt, mode, d_out, d_cnt);
    });
}
extern "C" {

int obvhs_cuda_cwbvh_ray_traverse_batch(ObvhsContext* ctx, const ObvhsCwBvh* bvh, const ObvhsRay* rays, size_t n, ObvhsRayHit* hits) {
    API_ENTER(ctx);
    return cw_traverse(ctx, bvh, rays, n, 0, hits, sizeof(ObvhsRayHit), nullptr);
}
int obvhs_cuda_cwbvh_ray_traverse_miss_batch(ObvhsContext* ctx, const ObvhsCwBvh* bvh, const ObvhsRay* rays, size_t n, uint8_t* miss) {
    API_ENTER(ctx);
    return cw_traverse(ctx, bvh, rays, n, 1, miss, 1, nullptr);
}
int obvhs_cuda_cwbvh_ray_traverse_anyhit_count_batch(ObvhsContext* ctx, const ObvhsCwBvh* bvh, const ObvhsRay* rays, size_t n,
                                                     uint32_t* counts) {
    API_ENTER(ctx);
    return cw_traverse(ctx, bvh, rays, n, 2, counts, 4, nullptr);
}
int obvhs_cuda_cwbvh_ray_traverse_batch_counted(ObvhsContext* ctx, const ObvhsCwBvh* bvh, const ObvhsRay* rays, size_t n, ObvhsRayHit* hits,
                                                uint64_t* counters) {
    API_ENTER(ctx);
    ARG_CHECK(ctx, counters, "counters is null");
    return cw_traverse(ctx, bvh, rays, n, 0, hits, sizeof(ObvhsRayHit), counters);
}

int obvhs_cuda_cwbvh_ray_new_traverse_batch_counted(ObvhsContext* ctx, const ObvhsCwBvh* bvh, const ObvhsRayNew* args, size_t n, ObvhsRayHit* hits,
                                                    uint64_t* counters) {
    API_ENTER(ctx);
    ARG_CHECK(ctx, counters, "counters is null");
    return cw_traverse(ctx, bvh, args, n, 0, hits, sizeof(ObvhsRayHit), counters);
}

// ---- multi-GPU (comm.cu) -----------------------------------------------------------------------------------------
int obvhs_cuda_nccl_unique_id(uint8_t id[OBVHS_NCCL_UNIQUE_ID_BYTES]) {
    if (!id) return OBVHS_ERR_INVALID_ARG;
    return comm_unique_id(id);
}
int obvhs_cuda_comm_init(ObvhsContext* ctx, const uint8_t id[OBVHS_NCCL_UNIQUE_ID_BYTES], int rank, int world) {
    API_ENTER(ctx);
    ARG_CHECK(ctx, id && world >= 1 && rank >= 0 && rank < world, "bad id / rank / world");
    return comm_init(ctx, id, rank, world);
}
int obvhs_cuda_cwbvh_broadcast(ObvhsContext* ctx, ObvhsCwBvh** bvh, int root) {
    API_ENTER(ctx);
    ARG_CHECK(ctx, bvh && root >= 0 && root < ctx->comm_world, "bad handle pointer / root");
    return comm_broadcast_cwbvh(ctx, bvh, root);
}

// ---- the same traversals over Ray::new arguments (32 B per ray) -------------------------------------------------------
int obvhs_cuda_ray_new_batch(ObvhsContext* ctx, const ObvhsRayNew* args, size_t n, ObvhsRay* rays) {
    API_ENTER(ctx);
    ARG_CHECK(ctx, n == 0 || (args && rays), "null args/rays");
    if (n == 0) return OBVHS_OK;
    DevBuf<ObvhsRayNew> st_in;
    DevBuf<ObvhsRay> st_out;
    const ObvhsRayNew* d_args = nullptr;
    ST_TRY(stage_in(ctx, args, n, st_in, &d_args));
    ObvhsRay* d_rays = rays;
    if (!obvhs_is_device_ptr(rays)) {
        CU_TRY(ctx, st_out.alloc(n, ctx->stream));
        d_rays = st_out.p;
    }
    ST_TRY(ray_new_device(ctx, d_args, n, d_rays));
    if (d_rays != rays) ST_TRY(copy_out(ctx, rays, (const ObvhsRay*)d_rays, n));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return OBVHS_OK;
}
int obvhs_cuda_cwbvh_ray_new_traverse_batch(ObvhsContext* ctx, const ObvhsCwBvh* bvh, const ObvhsRayNew* args, size_t n, ObvhsRayHit* hits) {
    API_ENTER(ctx);
    return cw_traverse(ctx, bvh, args, n, 0, hits, sizeof(ObvhsRayHit), nullptr);
}
int obvhs_cuda_cwbvh_ray_new_traverse_miss_batch(ObvhsContext* ctx, const ObvhsCwBvh* bvh, const ObvhsRayNew* args, size_t n, uint8_t* miss) {
    API_ENTER(ctx);
    return cw_traverse(ctx, bvh, args, n, 1, miss, 1, nullptr);
}
int obvhs_cuda_cwbvh_ray_new_traverse_anyhit_count_batch(ObvhsContext* ctx, const ObvhsCwBvh* bvh, const ObvhsRayNew* args, size_t n,
                                                         uint32_t* counts) {
    API_ENTER(ctx);
    return cw_traverse(ctx, bvh, args, n, 2, counts, 4, nullptr);
}

// ---- Bvh2 ray traversal, collapse, builder -----------------------------------------------
// rays[i] = Ray::new(od[i].origin, od[i].direction, tmin, tmax): 24 bytes per ray, one pair of bounds per batch
int obvhs_cuda_cwbvh_ray_od_traverse_batch(ObvhsContext* ctx, const ObvhsCwBvh* bvh, const ObvhsRayOd* od, size_t n, float tmin, float tmax,
                                           ObvhsRayHit* hits) {
    API_ENTER(ctx);
    return cw_traverse(ctx, bvh, od, n, 0, hits, sizeof(ObvhsRayHit), nullptr, tmin, tmax);
}
int obvhs_cuda_cwbvh_ray_od_traverse_hit8_batch(ObvhsContext* ctx, const ObvhsCwBvh* bvh, const ObvhsRayOd* od, size_t n, float tmin, float tmax,
                                                ObvhsRayHit8* hits) {
    API_ENTER(ctx);
    return cw_traverse(ctx, bvh, od, n, 0, hits, sizeof(ObvhsRayHit8), nullptr, tmin, tmax);
}
int obvhs_cuda_cwbvh_ray_od_traverse_miss_batch(ObvhsContext* ctx, const ObvhsCwBvh* bvh, const ObvhsRayOd* od, size_t n, float tmin,
                                                float tmax, uint8_t* miss) {
    API_ENTER(ctx);
    return cw_traverse(ctx, bvh, od, n, 1, miss, 1, nullptr, tmin, tmax);
}
int obvhs_cuda_bvh2_ray_od_traverse_batch(ObvhsContext* ctx, const ObvhsBvh2* bvh, const ObvhsRayOd* od, size_t n, float tmin, float tmax,
                                          ObvhsRayHit* hits) {
    API_ENTER(ctx);
    return b2_traverse(ctx, bvh, od, n, 0, hits, sizeof(ObvhsRayHit), nullptr, tmin, tmax);
}
int obvhs_cuda_bvh2_ray_traverse_batch(ObvhsContext* ctx, const ObvhsBvh2* bvh, const ObvhsRay* rays, size_t n, ObvhsRayHit* hits) {
    API_ENTER(ctx);
    return b2_traverse(ctx, bvh, rays, n, 0, hits, sizeof(ObvhsRayHit), nullptr);
}
int obvhs_cuda_bvh2_ray_traverse_miss_batch(ObvhsContext* ctx, const ObvhsBvh2* bvh, const ObvhsRay* rays, size_t n, uint8_t* miss) {
    API_ENTER(ctx);
    return b2_traverse(ctx, bvh, rays, n, 1, miss, 1, nullptr);
}
int obvhs_cuda_bvh2_ray_traverse_anyhit_count_batch(ObvhsContext* ctx, const ObvhsBvh2* bvh, const ObvhsRay* rays, size_t n, uint32_t* counts) {
    API_ENTER(ctx);
    return b2_traverse(ctx, bvh, rays, n, 2, counts, 4, nullptr);
}
int obvhs_cuda_bvh2_ray_traverse_batch_counted(ObvhsContext* ctx, const ObvhsBvh2* bvh, const ObvhsRay* rays, size_t n, ObvhsRayHit* hits,
                                               uint64_t* counters) {
    API_ENTER(ctx);
    ARG_CHECK(ctx, counters, "counters is null");
    return b2_traverse(ctx, bvh, rays, n, 0, hits, sizeof(ObvhsRayHit), counters);
}
int obvhs_cuda_bvh2_ray_new_traverse_batch(ObvhsContext* ctx, const ObvhsBvh2* bvh, const ObvhsRayNew* args, size_t n, ObvhsRayHit* hits) {
    API_ENTER(ctx);
    return b2_traverse(ctx, bvh, args, n, 0, hits, sizeof(ObvhsRayHit), nullptr);
}
int obvhs_cuda_bvh2_ray_new_traverse_miss_batch(ObvhsContext* ctx, const ObvhsBvh2* bvh, const ObvhsRayNew* args, size_t n, uint8_t* miss) {
    API_ENTER(ctx);
    return b2_traverse(ctx, bvh, args, n, 1, miss, 1, nullptr);
}
int obvhs_cuda_bvh2_set_triangles(ObvhsContext* ctx, ObvhsBvh2* bvh, const ObvhsTriangle* tris, size_t n) {
    API_ENTER(ctx);
    ARG_CHECK(ctx, bvh && (n == 0 || tris), "null argument");
    DevBuf<ObvhsTriangle> st;
    const ObvhsTriangle* d = nullptr;
    ST_TRY(stage_in(ctx, tris, n, st, &d));
    ST_TRY(bvh2_permute_tris_device(ctx, bvh, d, n));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return OBVHS_OK;
}
int obvhs_cuda_bvh2_collapse(ObvhsContext* ctx, ObvhsBvh2* bvh, uint32_t max_prims, float traversal_cost) {
    API_ENTER(ctx);
    ARG_CHECK(ctx, bvh, "bvh is null");
    if (bvh->bvh_tris) {  // primitive_indices is about to change: the permuted triangles would be stale
        obvhs_result_free(bvh->owner, bvh->bvh_tris);
        bvh->bvh_tris = nullptr;
    }
    return bvh2_collapse_device(ctx, bvh, max_prims, traversal_cost);
}
// build_bvh2<T: Boundable> (bvh2/builder.rs:103-140): PLOC -> reinsertion -> collapse -> reinsertion; pre_split is ignored
int obvhs_cuda_build_bvh2(ObvhsContext* ctx, const ObvhsAabb* aabbs, size_t n, const ObvhsBuildParams* params, double* core_build_seconds,
                          ObvhsBvh2** out) {
    API_ENTER(ctx);
    ARG_CHECK(ctx, params && out, "null argument");
    ARG_CHECK(ctx, n == 0 || aabbs, "aabbs is null");
    DevBuf<ObvhsAabb> st;
    const ObvhsAabb* d_aabbs = nullptr;
    ST_TRY(stage_in(ctx, aabbs, n, st, &d_aabbs));
    CU_TRY(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
    ObvhsBvh2* bvh2 = nullptr;
    ST_TRY(ploc_build_device(ctx, d_aabbs, nullptr, nullptr, n, params->ploc_search_distance, params->sort_precision,
                             (size_t)params->search_depth_threshold, &bvh2, nullptr));
    struct Guard {
        ObvhsBvh2* b;
        ~Guard() { if (b) obvhs_cuda_bvh2_free(b); }
    } guard{bvh2};
    ST_TRY(reinsertion_run_device(ctx, bvh2, params->reinsertion_batch_ratio, nullptr, 0, nullptr));
    u32 mp = params->max_prims_per_leaf < 1 ? 1 : (params->max_prims_per_leaf > 255 ? 255 : params->max_prims_per_leaf);  // builder.rs:121
    ST_TRY(bvh2_collapse_device(ctx, bvh2, mp, params->collapse_traversal_cost));
    ST_TRY(reinsertion_run_device(ctx, bvh2, params->reinsertion_batch_ratio * params->post_collapse_reinsertion_batch_ratio_multiplier,
                                  nullptr, 0, nullptr));
    CU_TRY(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    if (core_build_seconds) {
        float ms = 0.f;
        CU_TRY(ctx, cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
        *core_build_seconds += (double)ms * 1e-3;
    }
    guard.b = nullptr;
    *out = bvh2;
    return OBVHS_OK;
}

int obvhs_cuda_build_bvh2_from_tris(ObvhsContext* ctx, const ObvhsTriangle* tris, size_t n, const ObvhsBuildParams* params,
                                    double* core_build_seconds, ObvhsBvh2** out) {
    API_ENTER(ctx);
    ARG_CHECK(ctx, params && out, "null argument");
    ARG_CHECK(ctx, n == 0 || tris, "tris is null");
    DevBuf<ObvhsTriangle> st_tris;
    const ObvhsTriangle* d_tris = nullptr;
    ST_TRY(stage_in(ctx, tris, n, st_tris, &d_tris));
    ObvhsBvh2* bvh2 = nullptr;
    ST_TRY(ploc_from_tris(ctx, d_tris, n, params, &bvh2));  // core_build_time: bvh2/builder.rs:41,60 .. 83
    struct Guard {
        ObvhsBvh2* b;
        ~Guard() { if (b) obvhs_cuda_bvh2_free(b); }
    } guard{bvh2};
    ST_TRY(reinsertion_run_device(ctx, bvh2, params->reinsertion_batch_ratio, nullptr, 0, nullptr));
    u32 mp = params->max_prims_per_leaf < 1 ? 1 : (params->max_prims_per_leaf > 255 ? 255 : params->max_prims_per_leaf);  // builder.rs:75
    ST_TRY(bvh2_collapse_device(ctx, bvh2, mp, params->collapse_traversal_cost));
    ST_TRY(reinsertion_run_device(ctx, bvh2, params->reinsertion_batch_ratio * params->post_collapse_reinsertion_batch_ratio_multiplier, nullptr, 0,
                                  nullptr));
    CU_TRY(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
    ST_TRY(bvh2_permute_tris_device(ctx, bvh2, d_tris, n));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    if (core_build_seconds) {
        float ms = 0.f;
        CU_TRY(ctx, cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
        *core_build_seconds += (double)ms * 1e-3;
    }
    guard.b = nullptr;
    *out = bvh2;
    return OBVHS_OK;
}

}  // extern "C"

// ---- broad-phase queries (query.cu) -----------------------------------------------------------------------------
namespace {
template <class Run>
int query_entry(ObvhsContext* ctx, const float* queries, size_t n, int floats_per_query, Run run) {
    DevBuf<float> st;
    const float* d = nullptr;
    ST_TRY(stage_in(ctx, queries, n * floats_per_query, st, &d));
    ST_TRY(run(reinterpret_cast<const float4*>(d)));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return OBVHS_OK;
}
}  // namespace
extern "C" {
int obvhs_cuda_bvh2_aabb_traverse_batch(ObvhsContext* ctx, const ObvhsBvh2* bvh, const ObvhsAabb* queries, size_t n, uint32_t* counts,
                                        uint32_t* leaf_ids, size_t capacity, size_t* total) {
    API_ENTER(ctx);
    ARG_CHECK(ctx, bvh && (n == 0 || queries), "null argument");
    return query_entry(ctx, reinterpret_cast<const float*>(queries), n, 8,
                       [&](const float4* d) { return bvh2_query_device(ctx, bvh, 0, d, n, counts, leaf_ids, capacity, total); });
}
int obvhs_cuda_bvh2_point_traverse_batch(ObvhsContext* ctx, const ObvhsBvh2* bvh, const float* points, size_t n, uint32_t* counts,
                                         uint32_t* leaf_ids, size_t capacity, size_t* total) {
    API_ENTER(ctx);
    ARG_CHECK(ctx, bvh && (n == 0 || points), "null argument");
    return query_entry(ctx, points, n, 4, [&](const float4* d) { return bvh2_query_device(ctx, bvh, 1, d, n, counts, leaf_ids, capacity, total); });
}
int obvhs_cuda_cwbvh_aabb_traverse_batch(ObvhsContext* ctx, const ObvhsCwBvh* bvh, const ObvhsAabb* queries, size_t n,
                                         const float* traversal_direction, uint32_t* counts, uint32_t* primitive_ids, size_t capacity,
                                         size_t* total) {
    API_ENTER(ctx);
    ARG_CHECK(ctx, bvh && (n == 0 || queries), "null argument");
    ARG_CHECK(ctx, !traversal_direction || !obvhs_is_device_ptr(traversal_direction), "traversal_direction must be host memory");
    return query_entry(ctx, reinterpret_cast<const float*>(queries), n, 8, [&](const float4* d) {
        return cwbvh_query_device(ctx, bvh, 0, d, n, traversal_direction, counts, primitive_ids, capacity, total);
    });
}
int obvhs_cuda_cwbvh_point_traverse_batch(ObvhsContext* ctx, const ObvhsCwBvh* bvh, const float* points, size_t n,
                                          const float* traversal_direction, uint32_t* counts, uint32_t* primitive_ids, size_t capacity,
                                          size_t* total) {
    API_ENTER(ctx);
    ARG_CHECK(ctx, bvh && (n == 0 || points), "null argument");
    ARG_CHECK(ctx, !traversal_direction || !obvhs_is_device_ptr(traversal_direction), "traversal_direction must be host memory");
    return query_entry(ctx, points, n, 4, [&](const float4* d) {
        return cwbvh_query_device(ctx, bvh, 1, d, n, traversal_direction, counts, primitive_ids, capacity, total);
    });
}
}  // extern "C"

extern "C" int obvhs_cuda_make_rays(ObvhsContext* ctx, const float* origin_dir, size_t n, float tmin, float tmax, ObvhsRay* rays) {
    API_ENTER(ctx);
    ARG_CHECK(ctx, n == 0 || (origin_dir && rays), "null argument");
    if (n == 0) return OBVHS_OK;
    DevBuf<float> st_od;
    DevBuf<ObvhsRay> st_rays;
    const float* d_od = nullptr;
    ST_TRY(stage_in(ctx, origin_dir, n * 6, st_od, &d_od));
    const bool out_dev = obvhs_is_device_ptr(rays);
    ObvhsRay* d_rays = rays;
    if (!out_dev) {
        CU_TRY(ctx, st_rays.alloc(n, ctx->stream));
        d_rays = st_rays.p;
    }
    ST_TRY(make_rays_device(ctx, d_od, n, tmin, tmax, d_rays));
    if (!out_dev) ST_TRY(copy_out(ctx, rays, (const ObvhsRay*)d_rays, n));
    return OBVHS_OK;
}


