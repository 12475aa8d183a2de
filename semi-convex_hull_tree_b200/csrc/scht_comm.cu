// Sharded-dataset mode with one PROCESS per GPU (include/scht.h, scht_comm_*): the exchange steps run inside the library
// over NCCL (NVLink / NVSwitch), so a host program only creates a shard handle and a communicator per rank.
//
// The reference has one device (Code/GPU_Convex_Tree.h:443-446); BASELINE north_star: "NCCL over NVLink is used only for
// the optional sharded-dataset mode, which merges per-GPU top-K lists".  Per batch of queries (identical on every rank):
//   1. seed scan on the rank whose shard holds the query's approximate leaf -> K-th squared distance per query
//   2. ncclAllReduce(min) of the bounds (4 B per query)
//   3. main scan of the rank's shard with the bound -> K sorted keys per query
//   4. ONE grouped ncclSend / ncclRecv exchange: rank g receives every rank's lists for its contiguous query slice
//   5. k_merge_partial merges the world_size lists of the slice; results for the slice stay on the rank's device
// NCCL is loaded at run time (dlopen of libnccl.so.2: the copy a torch process has already loaded, or the system one), so
// the library and the single-GPU / single-process paths do not depend on it.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "scht_handle.h"

using namespace scht;

namespace {

struct NcclApi {
    void* lib = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclAllReduce) AllReduce = nullptr;
    decltype(&ncclSend) Send = nullptr;
    decltype(&ncclRecv) Recv = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    std::string error;
};

NcclApi* nccl_api() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, []() {
        for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
            api.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (api.lib) break;
        }
        if (!api.lib) {
            api.error = std::string("cannot load libnccl.so.2: ") + (dlerror() ? dlerror() : "not found");
            return;
        }
#define SCHT_NCCL_SYM(field, sym)                                                  \
    api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.lib, #sym));       \
    if (!api.field && api.error.empty()) api.error = "libnccl lacks " #sym;
        SCHT_NCCL_SYM(GetUniqueId, ncclGetUniqueId)
        SCHT_NCCL_SYM(CommInitRank, ncclCommInitRank)
        SCHT_NCCL_SYM(CommDestroy, ncclCommDestroy)
        SCHT_NCCL_SYM(AllReduce, ncclAllReduce)
        SCHT_NCCL_SYM(Send, ncclSend)
        SCHT_NCCL_SYM(Recv, ncclRecv)
        SCHT_NCCL_SYM(GroupStart, ncclGroupStart)
        SCHT_NCCL_SYM(GroupEnd, ncclGroupEnd)
        SCHT_NCCL_SYM(GetErrorString, ncclGetErrorString)
#undef SCHT_NCCL_SYM
    });
    return &api;
}

int nccl_ready(NcclApi*& api) {
    api = nccl_api();
    if (!api->error.empty()) return scht::fail(SCHT_ERR_NO_DEVICE, "NCCL unavailable: " + api->error);
    return SCHT_OK;
}

#define NCCL_TRY(api, expr)                                                                                   \
    do {                                                                                                      \
        ncclResult_t r__ = (expr);                                                                            \
        if (r__ != ncclSuccess) return scht::fail(SCHT_ERR_CUDA, std::string(#expr) + ": " + (api)->GetErrorString(r__)); \
    } while (0)

}  // namespace

struct scht_comm {
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1, device = 0;
    DevBuf gather;  // [world][slice][K] keys received for this rank's query slice
};

static_assert(SCHT_COMM_ID_BYTES == sizeof(ncclUniqueId), "scht.h: SCHT_COMM_ID_BYTES must equal sizeof(ncclUniqueId)");

extern "C" int scht_comm_unique_id(void* id_out) {
    if (!id_out) return scht::fail(SCHT_ERR_INVALID_ARG, "scht_comm_unique_id: NULL");
    NcclApi* api;
    int rc = nccl_ready(api);
    if (rc) return rc;
    ncclUniqueId id;
    NCCL_TRY(api, api->GetUniqueId(&id));
    memcpy(id_out, &id, sizeof(id));
    return SCHT_OK;
}

extern "C" int scht_comm_create(const void* id, int32_t rank, int32_t world, int32_t device, scht_comm** out) {
    if (!out) return scht::fail(SCHT_ERR_INVALID_ARG, "scht_comm_create: out is NULL");
    *out = nullptr;
    if (!id || world < 1 || world > kMaxShards || rank < 0 || rank >= world) return scht::fail(SCHT_ERR_INVALID_ARG, "scht_comm_create: bad rank / world size (<= 16)");
    NcclApi* api;
    int rc = nccl_ready(api);
    if (rc) return rc;
    CU_TRY(cudaSetDevice(device));
    ncclUniqueId uid;
    memcpy(&uid, id, sizeof(uid));
    scht_comm* c = new scht_comm();
    c->rank = rank;
    c->world = world;
    c->device = device;
    ncclResult_t r = api->CommInitRank(&c->comm, world, uid, rank);
    if (r != ncclSuccess) {
        delete c;
        return scht::fail(SCHT_ERR_CUDA, std::string("ncclCommInitRank: ") + api->GetErrorString(r));
    }
    *out = c;
    return SCHT_OK;
}

extern "C" void scht_comm_destroy(scht_comm* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    c->gather.release();
    NcclApi* api = nccl_api();
    if (c->comm && api->CommDestroy) api->CommDestroy(c->comm);
    delete c;
}

extern "C" int scht_comm_query_slice(const scht_comm* c, int64_t n_queries, int64_t* begin, int64_t* end) {
    if (!c || n_queries < 0) return scht::fail(SCHT_ERR_INVALID_ARG, "scht_comm_query_slice: bad arguments");
    const int64_t base = n_queries / c->world, rem = n_queries % c->world;
    const int64_t b = c->rank * base + std::min<int64_t>(c->rank, rem);
    if (begin) *begin = b;
    if (end) *end = b + base + (c->rank < rem ? 1 : 0);
    return SCHT_OK;
}

extern "C" int scht_query_sharded_device(scht_handle* h, scht_comm* c, const float* d_queries, int64_t n_queries, int32_t dim, int32_t K,
                                         int32_t* d_idx_out, float* d_dist2_out, float* d_dist_out, void* cuda_stream, scht_stats* stats) {
    if (!c) return scht::fail(SCHT_ERR_INVALID_ARG, "scht_query_sharded_device: communicator is NULL");
    int rc = check_query_args(h, d_queries, n_queries, dim, K, d_idx_out, d_dist2_out);
    if (rc) return rc;
    if (!h->subs.empty()) return scht::fail(SCHT_ERR_INVALID_ARG, "scht_query_sharded_device needs a single-GPU (shard) handle");
    if (h->device != c->device) return scht::fail(SCHT_ERR_INVALID_ARG, "handle and communicator live on different devices");
    if (n_queries > 0x7fffffffLL / std::max(K, dim)) return scht::fail(SCHT_ERR_INVALID_ARG, "n_queries * max(K, dim) must stay below 2^31");
    if (stats) memset(stats, 0, sizeof(*stats));
    if (n_queries == 0) return SCHT_OK;
    NcclApi* api;
    if ((rc = nccl_ready(api))) return rc;
    const double t0 = now_ms();
    CU_TRY(cudaSetDevice(h->device));
    cudaStream_t s = cuda_stream ? (cudaStream_t)cuda_stream : h->stream;
    const int nq = (int)n_queries, W = c->world;
    std::vector<int64_t> lo(W + 1);
    {
        const int64_t base = n_queries / W, rem = n_queries % W;
        lo[0] = 0;
        for (int g = 0; g < W; g++) lo[g + 1] = lo[g] + base + (g < rem ? 1 : 0);
    }
    const int my_n = (int)(lo[c->rank + 1] - lo[c->rank]);
    CU_TRY(h->ext_bound.reserve((size_t)nq * 4));
    CU_TRY(h->tmp_keys.reserve((size_t)nq * K * 8));
    CU_TRY(c->gather.reserve((size_t)W * std::max(my_n, 1) * K * 8));
    float* bound = h->ext_bound.as<float>();
    uint64_t* keys = h->tmp_keys.as<uint64_t>();
    // 1. seed scan of the resident leaves, 2. bounds = min over the ranks
    rc = batch_seed(h, d_queries, nq, K, kBatchPartial, h->leaf_lo, h->leaf_hi, s, stats, stats != nullptr);
    if (rc) return rc;
    if ((rc = batch_bounds(h, nq, K, bound, s))) return rc;
    NCCL_TRY(api, api->AllReduce(bound, bound, (size_t)nq, ncclFloat, ncclMin, c->comm, s));
    // 3. main scan of the shard
    QueryOut out;
    out.keys = keys;
    if ((rc = batch_main(h, d_queries, nq, K, kBatchPartial, h->leaf_lo, h->leaf_hi, bound, out, s, stats, stats != nullptr))) return rc;
    // 4. every rank's lists of slice g go to rank g
    NCCL_TRY(api, api->GroupStart());
    for (int g = 0; g < W; g++) {
        const size_t n_send = (size_t)(lo[g + 1] - lo[g]) * K;
        if (n_send) NCCL_TRY(api, api->Send(keys + (size_t)lo[g] * K, n_send, ncclUint64, g, c->comm, s));
        if (my_n) NCCL_TRY(api, api->Recv(c->gather.as<uint64_t>() + (size_t)g * my_n * K, (size_t)my_n * K, ncclUint64, g, c->comm, s));
    }
    NCCL_TRY(api, api->GroupEnd());
    // 5. merge this rank's slice
    if (my_n) {
        if ((rc = merge_lists(h, c->gather.as<uint64_t>(), W, my_n, K, false, d_idx_out, d_dist2_out, d_dist_out, s))) return rc;
    }
    CU_TRY(cudaStreamSynchronize(s));
    if (stats) stats->ms_total = now_ms() - t0;
    return SCHT_OK;
}
