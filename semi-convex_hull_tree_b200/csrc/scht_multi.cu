// Multi-GPU layer of the C ABI (include/scht.h, scht_create_multi): ONE process, one host thread per GPU, no NCCL.
//
// The reference has a single device (Code/GPU_Convex_Tree.h:443-446); SURVEY 8e defines two ways to use several:
//   SCHT_SHARD_QUERIES  the tree is replicated, every GPU answers a contiguous slice of the queries of a call; no GPU ever
//                       talks to another (queries are independent, Code/kernels/do_kNN_gpu.cl:322-326)
//   SCHT_SHARD_DATASET  every GPU holds a contiguous range of LEAVES (1/G of the point matrix and of the prefilter image,
//                       the tree top is replicated) and scans it for ALL queries of a batch.  Per batch:
//                         1. seed scan on the GPU that holds the query's approximate leaf -> K-th distance as a bound
//                         2. every GPU takes the minimum of all GPUs' bounds (peer reads over NVLink, 4 B per query)
//                         3. main scan of the shard with that bound; the scan kernel's epilogue stores each query's K keys
//                            straight into the gather buffer of the GPU that owns the query's slice (peer stores over
//                            NVLink): the exchange is fused into the scan, there is no separate all-to-all
//                         4. the owner merges the G sorted lists of its slice (k_merge_partial) and copies the result out
//                       Keys (d2, approximate-leaf bit, sorted position) are global, so the merge of the per-shard
//                       K-smallest sets is the global K-smallest set: results are bit-identical to one GPU.
#include <cuda_runtime.h>

#include <algorithm>
#include <condition_variable>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

#include "scht_handle.h"

using namespace scht;

namespace {

struct Barrier {
    std::mutex m;
    std::condition_variable cv;
    int n, waiting = 0;
    unsigned gen = 0;
    explicit Barrier(int n_) : n(n_) {}
    void wait() {
        std::unique_lock<std::mutex> lk(m);
        const unsigned g = gen;
        if (++waiting == n) {
            waiting = 0;
            gen++;
            cv.notify_all();
        } else {
            cv.wait(lk, [&] { return gen != g; });
        }
    }
};

void add_stats(scht_stats& a, const scht_stats& b) {
    a.n_queries += b.n_queries;
    a.dist_computations += b.dist_computations;
    a.exact_reevaluations += b.exact_reevaluations;
    a.list_insertions += b.list_insertions;
    a.hyperplane_tests += b.hyperplane_tests;
    a.descent_dists += b.descent_dists;
    a.batches = std::max(a.batches, b.batches);
    a.ms_device = std::max(a.ms_device, b.ms_device);
    a.ms_scan = std::max(a.ms_scan, b.ms_scan);
    a.scan_launches += b.scan_launches;
    a.kernel_launches += b.kernel_launches;
    a.pages_listed += b.pages_listed;
    a.prefilter_batches = std::max(a.prefilter_batches, b.prefilter_batches);
    a.ms_seed_scan = std::max(a.ms_seed_scan, b.ms_seed_scan);
    a.ms_traverse = std::max(a.ms_traverse, b.ms_traverse);
    a.ms_main_scan = std::max(a.ms_main_scan, b.ms_main_scan);
}

void split(int64_t n, int parts, std::vector<int64_t>& lo) {
    lo.resize(parts + 1);
    const int64_t base = n / parts, rem = n % parts;
    lo[0] = 0;
    for (int g = 0; g < parts; g++) lo[g + 1] = lo[g] + base + (g < rem ? 1 : 0);
}

// leaves [0, L) in `parts` contiguous ranges holding about the same number of points
void leaf_shards(const scht_tree_desc* t, int parts, std::vector<int>& cut) {
    cut.assign(parts + 1, 0);
    cut[parts] = t->n_leaves;
    int leaf = 0;
    for (int g = 1; g < parts; g++) {
        const int64_t target = t->n_points * g / parts;
        while (leaf < t->n_leaves && t->leaf_start[leaf] < target) leaf++;
        cut[g] = std::min(std::max(leaf, cut[g - 1] + 1), t->n_leaves - (parts - g));
    }
}

}  // namespace

namespace scht {

void multi_destroy(scht_handle* h) {
    for (scht_handle* s : h->subs) scht_destroy(s);
    h->subs.clear();
}

// ---- SCHT_SHARD_QUERIES: contiguous query slices, one independent single-GPU call per device ----------------------
static int multi_query_sharded_queries(scht_handle* h, const float* queries, int64_t n_queries, int32_t dim, int32_t K, int32_t* idx_out,
                                       float* dist2_out, float* dist_out, scht_stats* stats, bool brute) {
    const int G = (int)h->subs.size();
    std::vector<int64_t> lo;
    split(n_queries, G, lo);
    std::vector<int> rcs(G, SCHT_OK);
    std::vector<std::string> errs(G);
    std::vector<scht_stats> sts(G);
    std::vector<std::thread> th;
    for (int g = 0; g < G; g++) {
        th.emplace_back([&, g]() {
            const int64_t n = lo[g + 1] - lo[g];
            memset(&sts[g], 0, sizeof(scht_stats));
            if (n == 0) return;
            rcs[g] = host_query(h->subs[g], queries + (size_t)lo[g] * dim, n, dim, K, idx_out + (size_t)lo[g] * K, dist2_out + (size_t)lo[g] * K,
                                dist_out ? dist_out + (size_t)lo[g] * K : nullptr, stats ? &sts[g] : nullptr, brute);
            if (rcs[g]) errs[g] = scht_last_error();
        });
    }
    for (auto& x : th) x.join();
    for (int g = 0; g < G; g++)
        if (rcs[g]) return scht::fail(rcs[g], "device " + std::to_string(h->subs[g]->device) + ": " + errs[g]);
    if (stats)
        for (int g = 0; g < G; g++) add_stats(*stats, sts[g]);
    return SCHT_OK;
}

// ---- SCHT_SHARD_DATASET ---------------------------------------------------------------------------------------------
static int multi_query_sharded_dataset(scht_handle* h, const float* queries, int64_t n_queries, int32_t dim, int32_t K, int32_t* idx_out,
                                       float* dist2_out, float* dist_out, scht_stats* stats, bool brute) {
    const int G = (int)h->subs.size();
    const int64_t bs = std::min<int64_t>(h->batch_size, 0x7fffffffLL / std::max(K, dim));
    const int64_t n_batches = (n_queries + bs - 1) / bs;
    Barrier bar(G);
    std::vector<int> rcs(G, SCHT_OK);
    std::vector<std::string> errs(G);
    std::vector<scht_stats> sts(G);
    std::vector<float*> bound_own(G, nullptr);   // per device: its seed bounds of the current batch
    std::vector<uint64_t*> gather(G, nullptr);   // per device: [G lists][its slice][K] keys, written by every device's scan
    volatile bool failed = false;
    const int kind = brute ? kBatchBrute : kBatchPartial;
    std::vector<std::thread> th;
    for (int d = 0; d < G; d++) {
        th.emplace_back([&, d]() {
            scht_handle* x = h->subs[d];
            memset(&sts[d], 0, sizeof(scht_stats));
            auto step = [&](int rc) {
                if (rc && !rcs[d]) {
                    rcs[d] = rc;
                    errs[d] = scht_last_error();
                    failed = true;
                }
            };
            auto cu = [&](cudaError_t e, const char* what) {
                if (e != cudaSuccess) step(scht::fail(e == cudaErrorMemoryAllocation ? SCHT_ERR_OOM : SCHT_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e)));
                return e == cudaSuccess;
            };
            cu(cudaSetDevice(x->device), "cudaSetDevice");
            cudaStream_t s = x->stream;
            for (int64_t b = 0; b < n_batches; b++) {
                const int64_t b0 = b * bs;
                const int nq = (int)std::min<int64_t>(bs, n_queries - b0);
                std::vector<int64_t> sl;
                split(nq, G, sl);
                const int my_n = (int)(sl[d + 1] - sl[d]);
                // buffers
                if (!failed) {
                    cu(x->q[0].reserve((size_t)nq * dim * 4), "queries") && cu(x->ext_bound.reserve((size_t)nq * 4), "bounds") &&
                        cu(x->bound_own.reserve((size_t)nq * 4), "bounds") && cu(x->tmp_keys.reserve((size_t)G * std::max(my_n, 1) * K * 8), "gather buffer") &&
                        cu(x->idx[0].reserve((size_t)std::max(my_n, 1) * K * 4), "idx") && cu(x->d2[0].reserve((size_t)std::max(my_n, 1) * K * 4), "d2") &&
                        (!dist_out || cu(x->dist[0].reserve((size_t)std::max(my_n, 1) * K * 4), "dist"));
                }
                bound_own[d] = x->bound_own.as<float>();
                gather[d] = x->tmp_keys.as<uint64_t>();
                // phase 1: all queries of the batch to this device, seed scan where the approximate leaf is resident, bounds
                if (!failed) cu(cudaMemcpyAsync(x->q[0].p, queries + (size_t)b0 * dim, (size_t)nq * dim * 4, cudaMemcpyHostToDevice, s), "H2D queries");
                if (!failed) step(batch_seed(x, x->q[0].as<float>(), nq, K, kind, x->leaf_lo, x->leaf_hi, s, stats ? &sts[d] : nullptr, stats != nullptr));
                if (!failed) step(batch_bounds(x, nq, K, x->bound_own.as<float>(), s));
                if (!failed) cu(cudaStreamSynchronize(s), "seed phase");
                bar.wait();  // A: every device's bounds and buffer addresses are published
                // phase 2: bound = min over the shards, main scan of the shard, keys scattered to the slice owners
                if (!failed) step(min_bounds(x, bound_own.data(), G, nq, x->ext_bound.as<float>(), s));
                QueryOut out;
                out.n_slices = G;
                out.list = d;
                for (int g = 0; g < G; g++) {
                    out.key_tab[g] = gather[g];
                    out.slice_lo[g] = (int)sl[g];
                }
                out.slice_lo[G] = (int)sl[G];
                if (!failed) step(batch_main(x, x->q[0].as<float>(), nq, K, kind, x->leaf_lo, x->leaf_hi, x->ext_bound.as<float>(), out, s, stats ? &sts[d] : nullptr, stats != nullptr));
                if (!failed) cu(cudaStreamSynchronize(s), "scan phase");
                bar.wait();  // B: every device has stored its lists into the owners' gather buffers
                // phase 3: merge this device's query slice and copy it out
                if (!failed && my_n > 0) {
                    step(merge_lists(x, x->tmp_keys.as<uint64_t>(), G, my_n, K, brute, x->idx[0].as<int32_t>(), x->d2[0].as<float>(),
                                     dist_out ? x->dist[0].as<float>() : nullptr, s));
                    const size_t off = (size_t)(b0 + sl[d]) * K, n = (size_t)my_n * K * 4;
                    if (!failed) {
                        cu(cudaMemcpyAsync(idx_out + off, x->idx[0].p, n, cudaMemcpyDeviceToHost, s), "D2H idx") &&
                            cu(cudaMemcpyAsync(dist2_out + off, x->d2[0].p, n, cudaMemcpyDeviceToHost, s), "D2H d2") &&
                            (!dist_out || cu(cudaMemcpyAsync(dist_out + off, x->dist[0].p, n, cudaMemcpyDeviceToHost, s), "D2H dist")) &&
                            cu(cudaStreamSynchronize(s), "merge phase");
                    }
                }
                bar.wait();  // C: nobody starts the next batch (and overwrites bounds / gather buffers) before every merge is done
            }
        });
    }
    for (auto& x : th) x.join();
    for (int d = 0; d < G; d++)
        if (rcs[d]) return scht::fail(rcs[d], "device " + std::to_string(h->subs[d]->device) + ": " + errs[d]);
    if (stats) {
        for (int d = 0; d < G; d++) add_stats(*stats, sts[d]);
        stats->n_queries = n_queries;  // every device saw every query
        stats->batches = n_batches;
    }
    return SCHT_OK;
}

int multi_query(scht_handle* h, const float* queries, int64_t n_queries, int32_t dim, int32_t K, int32_t* idx_out, float* dist2_out,
                float* dist_out, scht_stats* stats, bool brute) {
    if (n_queries == 0) return SCHT_OK;
    const double t0 = now_ms();
    int rc = (h->multi_mode == SCHT_SHARD_DATASET)
                 ? multi_query_sharded_dataset(h, queries, n_queries, dim, K, idx_out, dist2_out, dist_out, stats, brute)
                 : multi_query_sharded_queries(h, queries, n_queries, dim, K, idx_out, dist2_out, dist_out, stats, brute);
    if (stats) stats->ms_total = now_ms() - t0;
    return rc;
}

}  // namespace scht

extern "C" int scht_create_multi(const scht_tree_desc* t, const int32_t* devices, int32_t n_devices, int32_t mode, scht_handle** out) {
    if (!out) return scht::fail(SCHT_ERR_INVALID_ARG, "scht_create_multi: out is NULL");
    *out = nullptr;
    if (mode != SCHT_SHARD_QUERIES && mode != SCHT_SHARD_DATASET) return scht::fail(SCHT_ERR_INVALID_ARG, "scht_create_multi: unknown mode");
    const int ndev = scht_device_count();
    if (ndev == 0) return scht::fail(SCHT_ERR_NO_DEVICE, "no CUDA device visible (this library has no CPU fallback)");
    std::vector<int> devs;
    if (!devices || n_devices <= 0) {
        for (int i = 0; i < ndev; i++) devs.push_back(i);
    } else {
        devs.assign(devices, devices + n_devices);
    }
    const int G = (int)devs.size();
    if (G > kMaxShards) return scht::fail(SCHT_ERR_INVALID_ARG, "scht_create_multi: at most 16 devices");
    for (int i = 0; i < G; i++) {
        if (devs[i] < 0 || devs[i] >= ndev) return scht::fail(SCHT_ERR_INVALID_ARG, "scht_create_multi: bad device ordinal");
        for (int j = 0; j < i; j++)
            if (devs[j] == devs[i]) return scht::fail(SCHT_ERR_INVALID_ARG, "scht_create_multi: a device is listed twice");
    }
    if (!t || t->n_leaves < 1) return scht::fail(SCHT_ERR_INVALID_ARG, "scht_create_multi: incomplete tree description");
    if (mode == SCHT_SHARD_DATASET && t->n_leaves < G) return scht::fail(SCHT_ERR_INVALID_ARG, "scht_create_multi: fewer leaves than devices");
    Options opt;
    options_from_env(opt);
    std::vector<float> leaf_means;
    if (opt.tc_mode != 0 && opt.groups != 1 && t->n_points >= 8192 && t->sorted_points && t->leaf_start && t->leaf_count) compute_leaf_means(t, leaf_means);
    std::vector<int> cut;
    if (mode == SCHT_SHARD_DATASET) leaf_shards(t, G, cut);
    std::vector<scht_handle*> subs(G, nullptr);
    std::vector<int> rcs(G, SCHT_OK);
    std::vector<std::string> errs(G);
    std::vector<std::thread> th;
    for (int g = 0; g < G; g++) {
        th.emplace_back([&, g]() {
            const int lb = mode == SCHT_SHARD_DATASET ? cut[g] : 0, le = mode == SCHT_SHARD_DATASET ? cut[g + 1] : t->n_leaves;
            rcs[g] = create_handle(t, devs[g], lb, le, &opt, leaf_means.empty() ? nullptr : leaf_means.data(), &subs[g]);
            if (rcs[g]) errs[g] = scht_last_error();
        });
    }
    for (auto& x : th) x.join();
    int rc = SCHT_OK;
    std::string err;
    for (int g = 0; g < G; g++)
        if (rcs[g] && !rc) {
            rc = rcs[g];
            err = "device " + std::to_string(devs[g]) + ": " + errs[g];
        }
    if (!rc && mode == SCHT_SHARD_DATASET && G > 1) {
        // every device reads the others' bounds and stores keys into the others' gather buffers: peer access both ways
        for (int a = 0; a < G && !rc; a++) {
            cudaSetDevice(devs[a]);
            for (int b = 0; b < G && !rc; b++) {
                if (a == b) continue;
                int can = 0;
                cudaDeviceCanAccessPeer(&can, devs[a], devs[b]);
                if (!can) {
                    rc = SCHT_ERR_CUDA;
                    err = "device " + std::to_string(devs[a]) + " cannot access device " + std::to_string(devs[b]) + " (sharded-dataset mode needs peer access)";
                    break;
                }
                cudaError_t e = cudaDeviceEnablePeerAccess(devs[b], 0);
                if (e == cudaErrorPeerAccessAlreadyEnabled) {
                    cudaGetLastError();
                } else if (e != cudaSuccess) {
                    rc = SCHT_ERR_CUDA;
                    err = std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e);
                }
            }
        }
    }
    if (rc) {
        for (scht_handle* s : subs)
            if (s) scht_destroy(s);
        return scht::fail(rc, err);
    }
    scht_handle* h = new scht_handle();
    h->subs = subs;
    h->multi_mode = mode;
    h->device = devs[0];
    h->opt = opt;
    h->batch_size = subs[0]->batch_size;
    *out = h;
    return SCHT_OK;
}
