// comm.cu -- multi-GPU plumbing below the C ABI: replicate a finished CwBvh over NCCL (NVLink 5 / NVSwitch).
//
// The reference has no multi-device path: a CwBvh is three Vecs + an Aabb, cloned freely (src/cwbvh/mod.rs:43-55), and every
// ray_traverse call only reads &self (:169), so rays shard trivially. The build runs on ONE GPU (PLOC iterations are globally
// ordered); this file sends the result to the other ranks: one 64-byte header broadcast (sizes, scene box, flags), then ONE
// grouped NCCL launch for nodes + primitive_indices + permuted triangles, straight out of / into the handles' device buffers
// on the context's stream. There is no collective on the traversal path.
//
// NCCL is bound at run time (dlopen "libnccl.so.2"): the static library keeps no link-time dependency, and a process that
// already holds NCCL (torch.distributed, or the Rust host's own linkage) shares that copy instead of loading a second one.
#include <dlfcn.h>
#include <nccl.h>
#include <string.h>

#include "common.cuh"

namespace {
struct NcclApi {
    void* lib = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclBroadcast) Broadcast = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    std::string error;
};
NcclApi* nccl_api() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* names[] = {getenv("OBVHS_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
        for (const char* n : names) {
            if (!n || !*n) continue;
            api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (api.lib) break;
        }
        if (!api.lib) {
            api.error = "libnccl.so.2 not found (set OBVHS_NCCL_LIB)";
            return;
        }
#define OBVHS_SYM(field, name)                                          \
    api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.lib, name)); \
    if (!api.field) api.error = std::string("missing NCCL symbol ") + name;
        OBVHS_SYM(GetUniqueId, "ncclGetUniqueId")
        OBVHS_SYM(CommInitRank, "ncclCommInitRank")
        OBVHS_SYM(CommDestroy, "ncclCommDestroy")
        OBVHS_SYM(Broadcast, "ncclBroadcast")
        OBVHS_SYM(GroupStart, "ncclGroupStart")
        OBVHS_SYM(GroupEnd, "ncclGroupEnd")
        OBVHS_SYM(GetErrorString, "ncclGetErrorString")
#undef OBVHS_SYM
    });
    return &api;
}
#define NCCL_TRY(ctx, api, expr)                                                                            \
    do {                                                                                                    \
        ncclResult_t _r = (expr);                                                                           \
        if (_r != ncclSuccess) {                                                                            \
            OBVHS_SET_ERR(ctx, "%s:%d %s -> %s", __FILE__, __LINE__, #expr, (api)->GetErrorString(_r));      \
            return OBVHS_ERR_NCCL;                                                                          \
        }                                                                                                   \
    } while (0)

// what the receivers need before they can allocate: 64 bytes, sent first
struct BroadcastHeader {
    u64 node_count, prim_count;
    u32 has_tris, uses_spatial_splits;
    u32 magic, _pad;
    ObvhsAabb total_aabb;
};
static_assert(sizeof(BroadcastHeader) == 64, "BroadcastHeader");
constexpr u32 HEADER_MAGIC = 0x0b5c3b11u;
}  // namespace

int comm_unique_id(uint8_t* id) {
    NcclApi* api = nccl_api();
    if (!api->error.empty()) return OBVHS_ERR_NCCL;
    static_assert(sizeof(ncclUniqueId) == OBVHS_NCCL_UNIQUE_ID_BYTES, "ncclUniqueId size");
    ncclUniqueId uid;
    if (api->GetUniqueId(&uid) != ncclSuccess) return OBVHS_ERR_NCCL;
    memcpy(id, &uid, sizeof(uid));
    return OBVHS_OK;
}

int comm_init(ObvhsContext* ctx, const uint8_t* id, int rank, int world) {
    NcclApi* api = nccl_api();
    if (!api->error.empty()) {
        OBVHS_SET_ERR(ctx, "NCCL unavailable: %s", api->error.c_str());
        return OBVHS_ERR_NCCL;
    }
    if (ctx->comm) {
        OBVHS_SET_ERR(ctx, "comm_init: the context already has a communicator");
        return OBVHS_ERR_INVALID_ARG;
    }
    ncclUniqueId uid;
    memcpy(&uid, id, sizeof(uid));
    ncclComm_t comm = nullptr;
    NCCL_TRY(ctx, api, api->CommInitRank(&comm, world, uid, rank));
    ctx->comm = comm;
    ctx->comm_rank = rank;
    ctx->comm_world = world;
    if (!ctx->comm_header) CU_TRY(ctx, cudaMalloc(&ctx->comm_header, sizeof(BroadcastHeader)));
    return OBVHS_OK;
}

void comm_destroy(ObvhsContext* ctx) {
    if (ctx->comm) {
        NcclApi* api = nccl_api();
        if (api->CommDestroy) api->CommDestroy(static_cast<ncclComm_t>(ctx->comm));
        ctx->comm = nullptr;
    }
    if (ctx->comm_header) {
        cudaFree(ctx->comm_header);
        ctx->comm_header = nullptr;
    }
}

// *bvh: the finished tree on `root`; on the other ranks the handle to (re)fill -- NULL, or a handle from an earlier broadcast
// whose buffers are reused when the sizes match (a per-frame rebroadcast then allocates nothing). Everything is enqueued on
// ctx->stream; the receivers make ONE host round trip (the 64-byte header), the root none.
int comm_broadcast_cwbvh(ObvhsContext* ctx, ObvhsCwBvh** bvh, int root) {
    NcclApi* api = nccl_api();
    if (!ctx->comm) {
        OBVHS_SET_ERR(ctx, "broadcast: call obvhs_cuda_comm_init first");
        return OBVHS_ERR_INVALID_ARG;
    }
    ncclComm_t comm = static_cast<ncclComm_t>(ctx->comm);
    const bool is_root = ctx->comm_rank == root;
    BroadcastHeader* h = static_cast<BroadcastHeader*>(ctx->pinned);
    if (is_root) {
        const ObvhsCwBvh* b = *bvh;
        if (!b) {
            OBVHS_SET_ERR(ctx, "broadcast: the root rank has no tree");
            return OBVHS_ERR_INVALID_ARG;
        }
        *h = BroadcastHeader{b->node_count, b->prim_count, b->bvh_tris ? 1u : 0u, b->uses_spatial_splits ? 1u : 0u, HEADER_MAGIC, 0u, b->total_aabb};
        CU_TRY(ctx, cudaMemcpyAsync(ctx->comm_header, h, sizeof(*h), cudaMemcpyHostToDevice, ctx->stream));
    }
    NCCL_TRY(ctx, api, api->Broadcast(ctx->comm_header, ctx->comm_header, sizeof(BroadcastHeader), ncclChar, root, comm, ctx->stream));
    ctx->launches++;
    BroadcastHeader hdr = *h;
    if (!is_root) {
        CU_TRY(ctx, cudaMemcpyAsync(h, ctx->comm_header, sizeof(*h), cudaMemcpyDeviceToHost, ctx->stream));
        CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        hdr = *h;
        if (hdr.magic != HEADER_MAGIC) {
            OBVHS_SET_ERR(ctx, "broadcast: corrupt header");
            return OBVHS_ERR_NCCL;
        }
        ObvhsCwBvh* b = *bvh;
        const bool reuse = b && b->owner == ctx && b->node_count == hdr.node_count && b->prim_count == hdr.prim_count &&
                           (b->bvh_tris != nullptr) == (hdr.has_tris != 0) && !b->exact_node_aabbs;
        if (!reuse) {
            if (b) obvhs_cuda_cwbvh_free(b);
            *bvh = nullptr;
            ST_TRY(obvhs_cuda_cwbvh_alloc(ctx, hdr.node_count, hdr.prim_count, (int)hdr.has_tris, &hdr.total_aabb, &b));
            *bvh = b;
        }
        b->total_aabb = hdr.total_aabb;
        b->uses_spatial_splits = hdr.uses_spatial_splits != 0;
    }
    ObvhsCwBvh* b = *bvh;
    NCCL_TRY(ctx, api, api->GroupStart());
    ncclResult_t r = ncclSuccess;
    if (hdr.node_count) r = api->Broadcast(b->nodes, b->nodes, hdr.node_count * sizeof(ObvhsCwBvhNode), ncclChar, root, comm, ctx->stream);
    if (r == ncclSuccess && hdr.prim_count)
        r = api->Broadcast(b->primitive_indices, b->primitive_indices, hdr.prim_count * 4, ncclChar, root, comm, ctx->stream);
    if (r == ncclSuccess && hdr.prim_count && hdr.has_tris)
        r = api->Broadcast(b->bvh_tris, b->bvh_tris, hdr.prim_count * (size_t)OBVHS_RT_TRIANGLE_BYTES, ncclChar, root, comm, ctx->stream);
    ncclResult_t r2 = api->GroupEnd();
    NCCL_TRY(ctx, api, r);
    NCCL_TRY(ctx, api, r2);
    ctx->launches++;
    return OBVHS_OK;
}
