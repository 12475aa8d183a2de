// Internal definition of scht_handle (include/scht.h), shared by the create path (scht_create.cu), the query path
// (scht_device.cu) and the multi-GPU layer (scht_multi.cu).  One plain handle = one GPU, one stream; a MULTI handle
// (scht_create_multi) owns one plain handle per device and no device state of its own.
#pragma once
#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "../../include/scht.h"
#include "scht_error.h"
#include "scht_types.cuh"

#define CU_TRY(expr)                                                                                              \
    do {                                                                                                          \
        cudaError_t e__ = (expr);                                                                                 \
        if (e__ != cudaSuccess)                                                                                   \
            return scht::fail(e__ == cudaErrorMemoryAllocation ? SCHT_ERR_OOM : SCHT_ERR_CUDA,                   \
                              std::string(#expr) + ": " + cudaGetErrorString(e__));                             \
    } while (0)

#ifndef SCHT_RESCORER_SLEEP_NS_DEFAULT
#define SCHT_RESCORER_SLEEP_NS_DEFAULT 4000
#endif

namespace scht {

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e == cudaSuccess) cap = bytes;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <class T>
    T* as() const { return reinterpret_cast<T*>(p); }
};

// pinned host staging buffer (pageable callers: results / queries pass through it so that copies run asynchronously)
struct PinBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
        cudaError_t e = cudaMallocHost(&p, bytes);
        if (e == cudaSuccess) cap = bytes;
        return e;
    }
    void release() {
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
    }
};

// Schedule knobs (scht_set_option; the environment variables of DESIGN.md 7d only provide the defaults).
// None of them changes a result bit.
struct Options {
    int scan_variant = -1;      // k_scan shape: -1 auto, 0..3 (kVariantWarps)
    int scan_stages = 0;        // slice ring depth of k_scan: 0 auto, 2 or 4
    int traverse_sampling = 1;  // sample the traversal / page selection before running it for every tile
    int sample_reuse = 7;       // batches that reuse two consecutive "list every page" verdicts of one batch shape
    int tc_mode = 2;            // tensor-core prefilter: 0 off, 1 always, 2 auto (per batch, from the seed bounds)
    int tc_split = 1;           // page segments per prefilter tile: 0 never, 1 auto, n > 1 forced
    int tc_issuers = 0;         // MMA issuer warps of k_scan_tc: 0 auto, 2 or 4
    int tc_n256 = -1;           // one N=256 MMA chain per page: -1 auto, 0, 1
    int seed_points = kSeedPoints;  // size of the seed neighbourhood
    int groups = -1;            // scan groups of the scan-order image: -1 auto (resident points / 2048, <= 8192), 1 = tree order
    int page_select = 1;        // ball test of scan pages before the prefilter scan (0: every page is listed)
    int brute_tc = 1;           // brute-force entry through the prefilter when eligible (0: exact FP32 scan only)
    int pipeline = 1;           // overlap H2D / kernels / D2H of consecutive batches of one scht_query call
    int rescorer_sleep_ns = SCHT_RESCORER_SLEEP_NS_DEFAULT;  // k_scan_tc: an idle rescorer warp's pause between polls (1500 and 8000 measured equal on uniform data)
    int first_pass = -1;        // pruned prefilter batches: 0 = bound test only, 1 = queries with a loose seed bound scan their nearest pages first, 2 = all queries do, -1 = auto
    int tc_ring = 0;            // k_scan_tc: survivor ring entries per epilogue warp: 0 auto (64 for K <= 32, else 256) or a power of two in 32..1024
    int ring_wait_ns = 1000;    // k_scan_tc: pause of an epilogue warp between polls while its survivor ring is full (40 -> 1000 ns: K=100 main scan 33.6 -> 32.3 ms, its polls share an issue port with a rescorer)
    int tc_recheck = 1;         // k_scan_tc, wide points: ring entries carry the prefilter value, stale ones are dropped before the exact evaluation
    int lazy_seed = 1;          // seed scan by fast-path distances, exact evaluation of the K entries that are left (batches whose main scan is the prefilter scan)
    int pipeline_split = 0;     // split a single-batch scht_query call into this many sub-batches so that its copies overlap (0/1: off)
};

}  // namespace scht

struct scht_handle {
    int device = 0;
    int max_smem = 0;
    int n_sm = 148;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;  // H2D of the next batch (scht_query pipeline)
    cudaStream_t out_stream = nullptr;   // D2H of the previous batch
    cudaEvent_t ev[6] = {};
    cudaEvent_t ev_in[2] = {}, ev_done[2] = {}, ev_out[2] = {};  // per pipeline slot: queries uploaded, kernels done, results downloaded
    scht::DevTree T{};
    scht::ScanImage I{};
    int depth = 0;
    int leaf_lo = 0, leaf_hi = 0;      // leaves resident on this device (whole tree: 0, n_leaves)
    std::vector<void*> owned;          // device allocations of the tree
    size_t tree_bytes = 0;
    int64_t batch_size = 1000000;      // Code/main.cc:670
    scht::Options opt;
    int sample_streak = 0, sample_skip = 0;
    long long sample_sig = -1;
    // the same for the eligibility verdict of the prefilter (tc = 2): two equal verdicts in a row are reused for sample_reuse batches
    int elig_streak = 0, elig_skip = 0, elig_last = -1;
    long long elig_sig = -1;
    int64_t tc_batches = 0;
    // state of the batch between batch_seed and batch_main
    struct { int variant = 0, tq = 0, launches = 0, scans = 0, kind = 0, nq = 0, K = 0, lazy_seed = 0; } bs;
    std::vector<int> h_leaf_start, h_leaf_count, h_parent;
    std::vector<scht::NodeRec> h_rec;
    scht::DevBuf q[2], idx[2], d2[2], dist[2];  // per pipeline slot
    scht::DevBuf q_leaf, hist, q_order, state, ranges, n_ranges, counters, seg_keys, ext_bound, bound_own, sel_bound, tmp_keys;
    scht::PinBuf pin_q[2], pin_out[2];
    // ---- multi-GPU handle (scht_create_multi): the plain handles it drives, one per device -----------------
    std::vector<scht_handle*> subs;
    int multi_mode = 0;  // SCHT_SHARD_QUERIES / SCHT_SHARD_DATASET
};

namespace scht {

// scht_create.cu
// leaf_means: optional host table [n_leaves][dim] (computed when NULL and needed); opt NULL = defaults + environment
int create_handle(const scht_tree_desc* t, int device, int leaf_begin, int leaf_end, const Options* opt, const float* leaf_means, scht_handle** out);
void compute_leaf_means(const scht_tree_desc* t, std::vector<float>& out);
int rebuild_seed_ranges(scht_handle* h);
void options_from_env(Options& o);
// scht_device.cu
struct QueryOut {
    int32_t* idx = nullptr;    // device
    float* d2 = nullptr;       // device
    float* dist = nullptr;     // device, optional
    uint64_t* keys = nullptr;  // device, key output (partial mode); with key_tab: unused
    // sharded-dataset scatter: keys of query qid go to key_tab[g] + ((size_t)list * (slice_hi[g]-slice_lo[g]) + qid - slice_lo[g]) * K
    // with g the slice containing qid (P2P stores into the owner's gather buffer)
    int n_slices = 0;
    int list = 0;
    uint64_t* key_tab[kMaxShards] = {};
    int slice_lo[kMaxShards + 1] = {};
};
enum BatchKind : int { kBatchTree = 0, kBatchBrute = 1, kBatchPartial = 2 };
// phase 1: descent, bucketing, seed scan (lists -> h->state); phase 2: (eligibility, page selection / traversal, main scan)
int batch_seed(scht_handle* h, const float* d_q, int nq, int K, int kind, int leaf_lo, int leaf_hi, cudaStream_t s, scht_stats* st, bool timed,
               bool lazy = false);
int batch_bounds(scht_handle* h, int nq, int K, float* d_bound_out, cudaStream_t s);  // bound[qid] = K-th seed distance^2 (FLT_MAX: list not full)
int batch_main(scht_handle* h, const float* d_q, int nq, int K, int kind, int leaf_lo, int leaf_hi, const float* d_ext_bound, const QueryOut& out,
               cudaStream_t s, scht_stats* st, bool timed);
int check_query_args(const scht_handle* h, const void* q, int64_t nq, int32_t dim, int32_t K, const void* idx, const void* d2);
int host_query(scht_handle* h, const float* queries, int64_t n_queries, int32_t dim, int32_t K, int32_t* idx_out, float* dist2_out,
               float* dist_out, scht_stats* stats, bool brute);
int merge_lists(scht_handle* h, const uint64_t* d_keys, int n_lists, int nq, int K, bool brute_keys, int32_t* d_idx, float* d_d2, float* d_dist,
                cudaStream_t s);
int min_bounds(scht_handle* h, float* const* d_bounds, int n, int nq, float* d_out, cudaStream_t s);
double now_ms();
// scht_multi.cu
int multi_query(scht_handle* h, const float* queries, int64_t n_queries, int32_t dim, int32_t K, int32_t* idx_out, float* dist2_out,
                float* dist_out, scht_stats* stats, bool brute);
void multi_destroy(scht_handle* h);

}  // namespace scht
