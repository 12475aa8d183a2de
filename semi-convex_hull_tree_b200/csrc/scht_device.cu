// CUDA side of the C ABI (include/scht.h): device tree upload and the batched exact k-NN query.
//
// Replaces GPU_Convex_Tree<T>'s OpenCL host code (reference: Code/GPU_Convex_Tree.h:150-163 ctor tail,
// :522-748 init_openCL_sorted_data_set, :208-323 do_kNN, :328-427 do_brute_force_kNN).  One handle = one GPU,
// one stream.  There is no CPU fallback: without an sm_100 device every entry point fails.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/scht.h"
#include "scht_error.h"
#include "scht_kernels.cuh"
#include "scht_tc.cuh"

using namespace scht;

#define CU_TRY(expr)                                                                                              \
    do {                                                                                                          \
        cudaError_t e__ = (expr);                                                                                 \
        if (e__ != cudaSuccess)                                                                                   \
            return scht::fail(e__ == cudaErrorMemoryAllocation ? SCHT_ERR_OOM : SCHT_ERR_CUDA,                   \
                              std::string(#expr) + ": " + cudaGetErrorString(e__));                             \
    } while (0)

namespace {

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

}  // namespace

struct scht_handle {
    int device = 0;
    int sm_count = 0;
    int max_smem = 0;
    int n_sm = 148;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[6] = {};
    DevTree T{};
    int depth = 0;
    std::vector<void*> owned;  // device allocations of the tree
    size_t tree_bytes = 0;
    int64_t batch_size = 1000000;  // Code/main.cc:670
    int scan_variant = -1;        // k_scan shape (SCHT_SCAN_VARIANT): -1 auto, 0..3 see kVariantWarps
    int scan_stages = 0;          // slice ring depth of k_scan's default variant: 0 auto, 2 or 4 (SCHT_SCAN_STAGES)
    int traverse_sampling = 1;    // sample the traversal before running it for every tile (SCHT_TRAVERSE_SAMPLING=0 disables)
    int sample_reuse = 7;         // after two consecutive batches whose sample said "list every page", the next n batches of the same
    int sample_streak = 0;        //   shape skip the sample (SCHT_SAMPLE_REUSE, 0 = sample every batch); results never depend on it
    int sample_skip = 0;
    long long sample_sig = -1;
    int tc_mode = 2;              // tensor-core prefilter: 0 off, 1 always, 2 auto (per batch, from the seed bounds)
    int tc_split = 1;             // page segments per prefilter tile: 0 never split, 1 auto, n > 1 forced (SCHT_TC_SPLIT)
    int tc_issuers = 0;           // MMA issuer warps of k_scan_tc: 0 auto (2 for one K slice per page, else 4), or 2 / 4 (SCHT_TC_ISSUERS)
    int tc_n256 = -1;             // one N=256 MMA chain per page (two issuers) instead of two N=128 chains: -1 auto (pages of several K slices), 0, 1 (SCHT_TC_N256)
    int kp = 0;                   // padded K of the tensor-core page image (0: not built)
    const unsigned char* tcpages = nullptr;
    float tc_scale = 1.f, pabs_sum = 0.f, norm_unit = 1.f;
    const float* d_mean = nullptr;
    const float* d_pmax = nullptr;
    float pnorm_max = 0.f;
    int64_t tc_batches = 0;
    std::vector<int> h_leaf_start, h_leaf_count;
    DevBuf q, q_leaf, hist, q_order, state, ranges, n_ranges, idx, d2, dist, counters, seg_keys;
};

namespace {

template <class T>
int upload(scht_handle* h, const T* src, size_t count, const T** dst) {
    void* d = nullptr;
    size_t bytes = std::max<size_t>(count, 1) * sizeof(T);
    CU_TRY(cudaMalloc(&d, bytes));
    h->owned.push_back(d);
    h->tree_bytes += bytes;
    if (count) CU_TRY(cudaMemcpyAsync(d, src, count * sizeof(T), cudaMemcpyHostToDevice, h->stream));
    *dst = reinterpret_cast<const T*>(d);
    return SCHT_OK;
}

// Leaf-scan kernel variants: consumer warps per CTA x resident CTAs per SM the register budget is sized for.
//   0: 8 warps (64 queries/tile), 1 CTA/SM   1: 7 warps (56 queries/tile, 256 threads with the producer), 2 CTAs/SM
//   2: 4 warps (32 queries/tile), 3 CTAs/SM   3: the same with 96 registers and a 2-stage ring, 4 CTAs/SM
// auto: 2, and 3 for the SEED scan of narrow points with short lists (1.50 -> 1.41 ms on 1Mx16; no gain for wide points or
// K = 100; the FP32-bound main scan keeps its 128 registers)
constexpr int kVariantWarps[4] = {8, 7, 4, 4};
int pick_variant(const scht_handle* h, int K) {
    int v = h->scan_variant;
    if (v < 0) v = 2;  // auto (the seed scan of narrow points runs as variant 3, see run_batch)
    int nw = kVariantWarps[v];
    if ((int)scan_smem_bytes(nw, h->T.dpad, K) <= h->max_smem) return v;
    if ((int)scan_smem_bytes(4, h->T.dpad, K) <= h->max_smem) return 2;
    return -1;
}

template <int NW, int MINB, bool FULL, int STAGES = kStages>
int launch_scan_t(scht_handle* h, const ScanParams& P, cudaStream_t s) {
    size_t smem = scan_smem_bytes(NW, P.T.dpad, P.K, STAGES);
    CU_TRY(cudaFuncSetAttribute(k_scan<NW, MINB, FULL, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int tq = NW * kQPerWarp;
    int n_tiles = (P.nq + tq - 1) / tq;
    k_scan<NW, MINB, FULL, STAGES><<<n_tiles, (NW + 1) * 32, smem, s>>>(P);
    CU_TRY(cudaGetLastError());
    return SCHT_OK;
}
int launch_scan(scht_handle* h, int variant, const ScanParams& P, cudaStream_t s) {
    const bool full = (P.T.dpad % kRowsPerStage) == 0;
    switch (variant) {
        case 0: return full ? launch_scan_t<8, 1, true>(h, P, s) : launch_scan_t<8, 1, false>(h, P, s);
        case 1: return full ? launch_scan_t<7, 2, true>(h, P, s) : launch_scan_t<7, 2, false>(h, P, s);
        case 3:  // 4 CTAs/SM (96 registers, 2-stage ring)
            return full ? launch_scan_t<4, 4, true, 2>(h, P, s) : launch_scan_t<4, 4, false, 2>(h, P, s);
        default: {
            // three resident CTAs per SM are what hides the exact path's FP64 chains: when the query block of wide points makes a
            // 4-stage CTA too big for that (D >= 64), a 2-stage ring brings the third CTA back
            const size_t per_sm = 233472 - 3 * 1024;
            const bool shallow = 3 * scan_smem_bytes(4, P.T.dpad, P.K) > per_sm && 3 * scan_smem_bytes(4, P.T.dpad, P.K, 2) <= per_sm && h->scan_stages != 4;
            if (shallow || h->scan_stages == 2) return full ? launch_scan_t<4, 3, true, 2>(h, P, s) : launch_scan_t<4, 3, false, 2>(h, P, s);
            return full ? launch_scan_t<4, 3, true>(h, P, s) : launch_scan_t<4, 3, false>(h, P, s);
        }
    }
}

int launch_traverse(scht_handle* h, const TraverseParams& P, cudaStream_t s) {
    int tq = P.tq;
    int n_tiles = (P.nq + tq - 1) / tq;
    int warps = 4;
    const int slots = (tq > 64) ? 128 : (tq > 32) ? 64 : 32;
    while (warps > 1 && (int)traverse_smem_bytes(warps, slots, P.T.dpad) > h->max_smem) warps >>= 1;
    size_t smem = traverse_smem_bytes(warps, slots, P.T.dpad);
    if ((int)smem > h->max_smem) return scht::fail(SCHT_ERR_INVALID_ARG, "dimension too large for the traversal kernel");
    long long jobs = (long long)selection_size(n_tiles, P.sel_mode, P.sel_step) * P.T.n_cuts;
    if (jobs == 0) return SCHT_OK;
    int blocks = (int)((jobs + warps - 1) / warps);
    if (tq > 64) {
        CU_TRY(cudaFuncSetAttribute(k_traverse<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_traverse<4><<<blocks, warps * 32, smem, s>>>(P);
    } else if (tq > 32) {
        CU_TRY(cudaFuncSetAttribute(k_traverse<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_traverse<2><<<blocks, warps * 32, smem, s>>>(P);
    } else {
        CU_TRY(cudaFuncSetAttribute(k_traverse<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_traverse<1><<<blocks, warps * 32, smem, s>>>(P);
    }
    CU_TRY(cudaGetLastError());
    return SCHT_OK;
}

struct QueryOut {
    int32_t* idx = nullptr;   // device
    float* d2 = nullptr;      // device
    float* dist = nullptr;    // device, optional
    uint64_t* keys = nullptr; // device, partial mode
};

// One batch, everything on the device.  brute: exhaustive scan keyed by original index.  partial: only leaves
// [leaf_lo,leaf_hi) and keys out.
int run_batch(scht_handle* h, const float* d_q, int nq, int K, bool brute, bool partial, int leaf_lo, int leaf_hi, const QueryOut& out,
              cudaStream_t s, scht_stats* st, bool timed) {
    const int variant = pick_variant(h, K);
    if (variant < 0) return scht::fail(SCHT_ERR_INVALID_ARG, "dimension * K too large for the leaf-scan kernel's shared memory");
    const int tq = kVariantWarps[variant] * kQPerWarp;
    const int n_tiles = (nq + tq - 1) / tq;
    const DevTree& T = h->T;

    CU_TRY(h->q_leaf.reserve((size_t)nq * 4));
    CU_TRY(h->q_order.reserve((size_t)nq * 4));
    CU_TRY(h->hist.reserve((size_t)(T.n_leaves + 1) * 4));
    CU_TRY(h->counters.reserve(16 * 8));
    CU_TRY(cudaMemsetAsync(h->counters.p, 0, 128, s));
    unsigned long long* ctr = h->counters.as<unsigned long long>();

    ScanParams P{};
    P.T = T;
    P.q = d_q;
    P.q_order = h->q_order.as<int>();
    P.q_leaf = h->q_leaf.as<int>();
    P.nq = nq;
    P.K = K;
    P.pos_lo = 0;
    P.pos_hi = T.n_points;
    P.idx_out = out.idx;
    P.d2_out = out.d2;
    P.dist_out = out.dist;
    P.keys_out = out.keys;
    P.counters = ctr;
    int launches = 0, scans = 0;
    bool used_tc = false;

    if (timed) CU_TRY(cudaEventRecord(h->ev[0], s));
    if (brute) {
        k_iota<<<(nq + 255) / 256, 256, 0, s>>>(h->q_order.as<int>(), nq);
        launches++;
        P.mode = kScanAll;
        P.write_mode = kWriteFinal;
        if (timed) CU_TRY(cudaEventRecord(h->ev[1], s));
        int rc = launch_scan(h, variant, P, s);
        if (rc) return rc;
        launches++;
        scans++;
        if (timed) {
            CU_TRY(cudaEventRecord(h->ev[2], s));
            CU_TRY(cudaEventRecord(h->ev[3], s));
            CU_TRY(cudaEventRecord(h->ev[4], s));
        }
    } else {
        // K1 descend + leaf histogram, K3 bucketing
        CU_TRY(cudaMemsetAsync(h->hist.p, 0, (size_t)(T.n_leaves + 1) * 4, s));
        k_descend<<<(nq + 127) / 128, 128, 0, s>>>(T, d_q, nq, h->q_leaf.as<int>(), h->hist.as<int>(), ctr);
        k_leaf_prefix_sum<<<1, 1024, 0, s>>>(h->hist.as<int>(), T.n_leaves);
        k_scatter<<<(nq + 255) / 256, 256, 0, s>>>(h->q_leaf.as<int>(), nq, h->hist.as<int>(), h->q_order.as<int>());
        CU_TRY(cudaGetLastError());
        launches += 3;
        if (timed) CU_TRY(cudaEventRecord(h->ev[1], s));

        CU_TRY(h->state.reserve((size_t)nq * K * 8));
        CU_TRY(h->ranges.reserve((size_t)n_tiles * kRangeCap * sizeof(int2)));
        CU_TRY(h->n_ranges.reserve((size_t)n_tiles * kMaxCuts * 4));
        P.state = h->state.as<uint64_t>();
        P.ranges = h->ranges.as<int2>();
        P.n_ranges = h->n_ranges.as<int>();
        if (partial) {
            // positions of the shard: leaves [leaf_lo, leaf_hi) are contiguous in the sorted matrix
            P.pos_lo = (leaf_lo < T.n_leaves) ? h->h_leaf_start[leaf_lo] : T.n_points;
            P.pos_hi = (leaf_hi > leaf_lo) ? h->h_leaf_start[leaf_hi - 1] + h->h_leaf_count[leaf_hi - 1] : P.pos_lo;
        }
        // K4a seed scan of the approximate leaves
        P.mode = kScanSeed;
        P.write_mode = kWriteState;
        const int seed_variant = (h->scan_variant < 0 && variant == 2 && T.dpad <= kRowsPerStage && K <= 32) ? 3 : variant;  // same 32-query tiles
        int rc = launch_scan(h, seed_variant, P, s);
        if (rc) return rc;
        launches++;
        scans++;
        if (timed) CU_TRY(cudaEventRecord(h->ev[2], s));
        // tensor-core prefilter or exact FP32 scan for the main pass?
        bool use_tc = false;
        int tc_stages = kTcMaxStages;
        while (tc_stages > 3 && (int)tc_smem_bytes(T.dpad, h->kp, K, tc_stages) > h->max_smem) tc_stages--;
        if (h->kp > 0 && h->tc_mode != 0 && (int)tc_smem_bytes(T.dpad, h->kp, K, tc_stages) <= h->max_smem) {
            if (h->tc_mode == 1) {
                use_tc = true;
            } else {
                k_tc_eligibility<<<(nq + 127) / 128, 128, 0, s>>>(T, d_q, P.q_order, nq, K, P.state, h->d_mean, h->d_pmax, h->pnorm_max, h->pabs_sum,
                                                                  h->tc_scale, ctr);
                launches++;
                unsigned long long ok = 0;
                CU_TRY(cudaMemcpyAsync(&ok, ctr + 6, 8, cudaMemcpyDeviceToHost, s));
                CU_TRY(cudaStreamSynchronize(s));
                use_tc = ok * 10 >= (unsigned long long)nq * 9;  // >= 90 % of the queries have a tight margin
            }
        }
        const int tq_main = use_tc ? kTcQ : tq;
        const int n_tiles_main = (nq + tq_main - 1) / tq_main;
        CU_TRY(h->ranges.reserve((size_t)std::max(n_tiles, n_tiles_main) * kRangeCap * sizeof(int2)));
        P.ranges = h->ranges.as<int2>();
        // K2 traversal with the seed bounds -> page ranges per tile
        TraverseParams TP{};
        TP.T = T;
        TP.q = d_q;
        TP.q_order = P.q_order;
        TP.q_leaf = P.q_leaf;
        TP.nq = nq;
        TP.K = K;
        TP.tq = tq_main;
        TP.state = P.state;
        TP.ranges = h->ranges.as<int2>();
        TP.n_ranges = h->n_ranges.as<int>();
        TP.leaf_lo = leaf_lo;
        TP.leaf_hi = leaf_hi;
        TP.counters = ctr;
        TP.ball_only = use_tc ? 1 : 0;
        // Sampled first: when the bounds keep >= 90 % of the pages for every 8th tile, they are not evaluated for the
        // other tiles (all their pages are listed); otherwise the remaining tiles are traversed too.
        // (about 24 sampled tiles, at least every 64th and at most every 8th tile)
        const int kSample = std::max(8, std::min(64, n_tiles_main / 24));
        const int page_lo = P.pos_lo / kPage, page_hi = (P.pos_hi + kPage - 1) / kPage;
        const long long sig = (long long)use_tc | ((long long)K << 1) | ((long long)leaf_lo << 12) | ((long long)leaf_hi << 36);
        if (n_tiles_main >= 64 && h->traverse_sampling && h->sample_skip > 0 && sig == h->sample_sig) {
            // the samples of the last batches of this shape all kept (nearly) every page: list them without asking again
            h->sample_skip--;
            k_list_all<<<(n_tiles_main + 127) / 128, 128, 0, s>>>(h->ranges.as<int2>(), h->n_ranges.as<int>(), n_tiles_main, T.n_cuts, 0, 1, page_lo, page_hi);
            CU_TRY(cudaGetLastError());
            launches++;
        } else if (n_tiles_main >= 64 && h->traverse_sampling) {
            unsigned long long before = 0, after = 0;
            TP.sel_mode = 1;
            TP.sel_step = kSample;
            rc = launch_traverse(h, TP, s);
            if (rc) return rc;
            launches++;
            CU_TRY(cudaMemcpyAsync(&after, ctr + 5, 8, cudaMemcpyDeviceToHost, s));
            CU_TRY(cudaStreamSynchronize(s));
            long long sampled_q = 0;
            for (int t = 0; t < n_tiles_main; t += kSample) sampled_q += std::min(tq_main, nq - t * tq_main);
            const double kept = (double)(after - before) / std::max(1.0, (double)sampled_q * std::max(1, page_hi - page_lo));
            TP.sel_mode = 2;
            // a prefilter page costs ~1/10 of an exact page, so pruning must remove much more before it pays for itself
            if (kept >= (use_tc ? 0.60 : 0.90)) {
                int rest = selection_size(n_tiles_main, 2, kSample);
                k_list_all<<<(rest + 127) / 128, 128, 0, s>>>(h->ranges.as<int2>(), h->n_ranges.as<int>(), n_tiles_main, T.n_cuts, 2, kSample, page_lo,
                                                             page_hi);
                CU_TRY(cudaGetLastError());
                if (sig != h->sample_sig) h->sample_streak = 0;
                if (++h->sample_streak >= 2) h->sample_skip = h->sample_reuse;
            } else {
                rc = launch_traverse(h, TP, s);
                if (rc) return rc;
                h->sample_streak = 0;
            }
            h->sample_sig = sig;
            launches++;
        } else {
            TP.sel_mode = 0;
            TP.sel_step = 1;
            rc = launch_traverse(h, TP, s);
            if (rc) return rc;
            launches++;
        }
        if (timed) CU_TRY(cudaEventRecord(h->ev[3], s));
        // K4b main scan
        P.mode = kScanList;
        P.write_mode = partial ? kWriteKeys : kWriteFinal;
        if (use_tc) {
            TcScanParams TC{};
            TC.S = P;
            TC.tcpages = h->tcpages;
            TC.mean = h->d_mean;
            TC.pmax = h->d_pmax;
            TC.pnorm_max = h->pnorm_max;
            TC.pabs_sum = h->pabs_sum;
            TC.scale = h->tc_scale;
            TC.norm_unit = h->norm_unit;
            TC.kp = h->kp;
            TC.n_stages = tc_stages;
#ifdef SCHT_TC_PROF
            static long long* d_trace = nullptr;
            if (!d_trace) cudaMalloc(&d_trace, 64 * 8 * 8);
            cudaMemsetAsync(d_trace, 0, 64 * 8 * 8, s);
            TC.trace = d_trace;
            if (const char* ev = getenv("SCHT_TC_DEBUG")) TC.debug = atoi(ev);
#endif
            // Tail effect: n_tiles CTAs of equal length on n_sm SMs take ceil(n_tiles/n_sm) rounds (782 tiles on 148 SMs: 6 rounds
            // for 5.28 rounds of work).  Splitting every tile's page list into S segments (own CTA each, partial lists merged by
            // k_merge_partial) makes the rounds S times shorter: pick the S <= 8 with the smallest ceil(S*n_tiles/n_sm)/S, charging
            // 5 % per extra segment for the repeated set-up and the looser early bounds (measured).  Also what keeps a SMALL batch from
            // running on a handful of SMs.
            int n_seg = 1;
            if (!partial && h->tc_split) {
                const int page_lo = P.pos_lo / kPage, page_hi = (P.pos_hi + kPage - 1) / kPage;
                double best = 1e30;
                for (int S = 1; S <= 8; S++) {
                    if (S > 1 && (page_hi - page_lo) / S < 64) break;
                    const double rounds = std::ceil((double)S * n_tiles_main / h->n_sm) / S * (1.0 + 0.05 * (S - 1));
                    if (rounds < best - 1e-9) {
                        best = rounds;
                        n_seg = S;
                    }
                }
                if (h->tc_split > 1) n_seg = std::min(8, h->tc_split);  // SCHT_TC_SPLIT=n forces n segments
                if (n_seg > 1) {
                    CU_TRY(h->seg_keys.reserve((size_t)n_seg * nq * K * sizeof(uint64_t)));
                    TC.S.keys_out = h->seg_keys.as<uint64_t>();
                    TC.n_seg = n_seg;
                    TC.seg_page0 = page_lo;
                    TC.seg_pages = (page_hi - page_lo + n_seg - 1) / n_seg;
                }
            }
            if (n_seg == 1) TC.n_seg = 1;
            size_t smem = tc_smem_bytes(T.dpad, h->kp, K, tc_stages);
            // One K slice per page (narrow points): two issuers, one per page parity, each issuing the page's two N=128 half-page
            // chains (+3 % over four issuers; whole-page N=256 chains cost 2.4 % there).  Several slices per page (wide points): the
            // MMA time counts, and one N=256 chain per page (128 instead of 2 x 92 cycles per K16 step) beats four issuers of
            // half-page chains by 2 % (D=128) to 5 % (D=32).
            const bool wide = h->kp > kTcKSlice;
            bool want_n256 = (h->tc_n256 < 0) ? wide : (h->tc_n256 != 0);
            int n_iss = (wide && !want_n256) ? 4 : 2;
            if (h->tc_issuers == 2 || h->tc_issuers == 4) n_iss = h->tc_issuers;  // SCHT_TC_ISSUERS
            auto launch_tc = [&](auto kernel, int threads) {
                cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                if (e == cudaSuccess) kernel<<<n_tiles_main * n_seg, threads, smem, s>>>(TC);
                return e;
            };
            const bool rows = T.rows != nullptr;  // wide points: exact re-evaluation from the row-major copy
            const bool n256 = n_iss == 2 && want_n256;
            if (n256) CU_TRY(rows ? launch_tc(k_scan_tc<2, true, true>, tc_threads(2)) : launch_tc(k_scan_tc<2, false, true>, tc_threads(2)));
            else if (n_iss == 2) CU_TRY(rows ? launch_tc(k_scan_tc<2, true, false>, tc_threads(2)) : launch_tc(k_scan_tc<2, false, false>, tc_threads(2)));
            else CU_TRY(rows ? launch_tc(k_scan_tc<4, true, false>, tc_threads(4)) : launch_tc(k_scan_tc<4, false, false>, tc_threads(4)));
            CU_TRY(cudaGetLastError());
            if (n_seg > 1) {
                k_merge_partial<<<(nq + 127) / 128, 128, 0, s>>>(T, h->seg_keys.as<uint64_t>(), n_seg, nq, K, P.idx_out, P.d2_out, P.dist_out);
                CU_TRY(cudaGetLastError());
                launches++;
            }
            h->tc_batches++;
#ifdef SCHT_TC_PROF
            {
                long long tr[64 * 8];
                cudaMemcpyAsync(tr, TC.trace, sizeof(tr), cudaMemcpyDeviceToHost, s);
                cudaStreamSynchronize(s);
                long long t0 = tr[2];
                for (int i = 0; i < 40; i++)
                    fprintf(stderr, "[trace pg %2d] prod: slot@%6lld sent@%6lld | mma: tempty@%6lld full@%6lld issued@%6lld | epi: tfull@%6lld released@%6lld done@%6lld\n", i,
                            tr[i*8+0]-t0, tr[i*8+1]-t0, tr[i*8+2]-t0, tr[i*8+3]-t0, tr[i*8+4]-t0, tr[i*8+5]-t0, tr[i*8+6]-t0, tr[i*8+7]-t0);
            }
#endif
            used_tc = true;
        } else {
            rc = launch_scan(h, variant, P, s);
            if (rc) return rc;
        }
        launches++;
        scans++;
        if (timed) CU_TRY(cudaEventRecord(h->ev[4], s));
    }
    if (st) {
        unsigned long long c[16];
        CU_TRY(cudaMemcpyAsync(c, ctr, 128, cudaMemcpyDeviceToHost, s));
        CU_TRY(cudaStreamSynchronize(s));
        st->n_queries += nq;
        st->dist_computations += (int64_t)c[0];
        st->exact_reevaluations += (int64_t)c[1];
        st->list_insertions += (int64_t)c[2];
        st->hyperplane_tests += (int64_t)c[4];
        st->descent_dists += (int64_t)c[3];
        st->batches += 1;
        st->scan_launches += scans;
        st->kernel_launches += launches;
        st->pages_listed += (int64_t)c[5];
        st->prefilter_batches += used_tc ? 1 : 0;
        if (timed) {
            float a = 0, b = 0, c2 = 0;
            CU_TRY(cudaEventElapsedTime(&a, h->ev[0], h->ev[4]));
            CU_TRY(cudaEventElapsedTime(&b, h->ev[1], h->ev[2]));
            CU_TRY(cudaEventElapsedTime(&c2, h->ev[3], h->ev[4]));
            st->ms_device += a;
            st->ms_scan += b + c2;
            st->ms_seed_scan += brute ? 0.f : b;
            st->ms_main_scan += brute ? b : c2;
            float tr = 0;
            CU_TRY(cudaEventElapsedTime(&tr, h->ev[2], h->ev[3]));
            st->ms_traverse += tr;
        }
    }
    return SCHT_OK;
}

int check_query_args(const scht_handle* h, const void* q, int64_t nq, int32_t dim, int32_t K, const void* idx, const void* d2) {
    if (!h) return scht::fail(SCHT_ERR_INVALID_ARG, "handle is NULL");
    if (nq < 0 || nq > 0x7fffffffLL) return scht::fail(SCHT_ERR_INVALID_ARG, "n_queries out of range");
    if (K < 1 || K > SCHT_MAX_K) return scht::fail(SCHT_ERR_INVALID_ARG, "K must be in [1, SCHT_MAX_K]");
    if (dim != h->T.dim) {
        // Code/Convex_Tree.h:458-461 prints "dimension of query points is not the same as the data" and returns
        return scht::fail(SCHT_ERR_DIM_MISMATCH, "query dimension " + std::to_string(dim) + " != tree dimension " + std::to_string(h->T.dim));
    }
    if (nq > 0 && (!q || !idx || !d2)) return scht::fail(SCHT_ERR_INVALID_ARG, "NULL query or output buffer");
    return SCHT_OK;
}

double now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

int host_query(scht_handle* h, const float* queries, int64_t n_queries, int32_t dim, int32_t K, int32_t* idx_out, float* dist2_out,
               float* dist_out, scht_stats* stats, bool brute) {
    int rc = check_query_args(h, queries, n_queries, dim, K, idx_out, dist2_out);
    if (rc) return rc;
    if (stats) memset(stats, 0, sizeof(*stats));
    double t0 = now_ms();
    CU_TRY(cudaSetDevice(h->device));
    int64_t bs = std::min<int64_t>(h->batch_size, 0x7fffffffLL / std::max(K, dim));
    for (int64_t b0 = 0; b0 < n_queries; b0 += bs) {
        int nq = (int)std::min<int64_t>(bs, n_queries - b0);
        CU_TRY(h->q.reserve((size_t)nq * dim * 4));
        CU_TRY(h->idx.reserve((size_t)nq * K * 4));
        CU_TRY(h->d2.reserve((size_t)nq * K * 4));
        if (dist_out) CU_TRY(h->dist.reserve((size_t)nq * K * 4));
        CU_TRY(cudaMemcpyAsync(h->q.p, queries + (size_t)b0 * dim, (size_t)nq * dim * 4, cudaMemcpyHostToDevice, h->stream));
        QueryOut out;
        out.idx = h->idx.as<int32_t>();
        out.d2 = h->d2.as<float>();
        out.dist = dist_out ? h->dist.as<float>() : nullptr;
        rc = run_batch(h, h->q.as<float>(), nq, K, brute, false, 0, h->T.n_leaves, out, h->stream, stats, stats != nullptr);
        if (rc) return rc;
        CU_TRY(cudaMemcpyAsync(idx_out + (size_t)b0 * K, out.idx, (size_t)nq * K * 4, cudaMemcpyDeviceToHost, h->stream));
        CU_TRY(cudaMemcpyAsync(dist2_out + (size_t)b0 * K, out.d2, (size_t)nq * K * 4, cudaMemcpyDeviceToHost, h->stream));
        if (dist_out) CU_TRY(cudaMemcpyAsync(dist_out + (size_t)b0 * K, out.dist, (size_t)nq * K * 4, cudaMemcpyDeviceToHost, h->stream));
        CU_TRY(cudaStreamSynchronize(h->stream));
    }
    if (stats) stats->ms_total = now_ms() - t0;
    return SCHT_OK;
}

}  // namespace

extern "C" int scht_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

extern "C" int scht_create(const scht_tree_desc* t, int device, scht_handle** out) {
    if (!out) return scht::fail(SCHT_ERR_INVALID_ARG, "scht_create: out is NULL");
    *out = nullptr;
    if (!t || t->dim < 1 || t->n_points < 1 || t->n_points > 0x7fffff00LL || t->n_nodes < 1 || t->n_leaves < 1 || !t->sorted_points ||
        !t->sorted_to_orig || !t->leaf_start || !t->leaf_count || !t->leaf_node_id || !t->node_left || !t->node_right || !t->node_parent ||
        !t->node_leaf_index || !t->node_left_center || !t->node_right_center || !t->node_normal || !t->node_beta_off || !t->node_beta)
        return scht::fail(SCHT_ERR_INVALID_ARG, "scht_create: incomplete tree description");
    int ndev = scht_device_count();
    if (ndev == 0) return scht::fail(SCHT_ERR_NO_DEVICE, "no CUDA device visible (this library has no CPU fallback)");
    if (device < 0 || device >= ndev) return scht::fail(SCHT_ERR_INVALID_ARG, "scht_create: bad device ordinal");
    cudaDeviceProp prop;
    CU_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) return scht::fail(SCHT_ERR_NO_DEVICE, std::string("device is sm_") + std::to_string(prop.major * 10 + prop.minor) + ", this library is built for sm_100a only");

    const int nn = t->n_nodes, L = t->n_leaves, dim = t->dim;
    const int64_t N = t->n_points;
    // ---- validate the structure and derive per-node records (host) -------------------------------------------
    std::vector<NodeRec> rec((size_t)nn);
    for (int i = 0; i < nn; i++) {
        bool leaf = t->node_leaf_index[i] >= 0;
        if (leaf != (t->node_left[i] < 0) || leaf != (t->node_right[i] < 0)) return scht::fail(SCHT_ERR_INVALID_ARG, "scht_create: node child links inconsistent");
        if (!leaf && (t->node_left[i] != i + 1 || t->node_right[i] <= i + 1 || t->node_right[i] >= nn))
            return scht::fail(SCHT_ERR_INVALID_ARG, "scht_create: nodes must be numbered in DFS pre-order (left child = id+1)");
        if (leaf && t->node_leaf_index[i] >= L) return scht::fail(SCHT_ERR_INVALID_ARG, "scht_create: leaf index out of range");
        rec[i].depth = t->node_beta_off[i + 1] - t->node_beta_off[i];
        rec[i].beta_off = t->node_beta_off[i];
        rec[i].leaf = leaf ? t->node_leaf_index[i] : -1;
        rec[i].right = leaf ? -1 : t->node_right[i];
        rec[i].pad = 0;
    }
    int depth = 0;
    for (int i = nn - 1; i >= 0; i--) {
        if (rec[i].leaf >= 0) {
            int li = rec[i].leaf;
            rec[i].size = 1;
            rec[i].pos0 = t->leaf_start[li];
            rec[i].pos1 = t->leaf_start[li] + t->leaf_count[li];
            if (rec[i].pos0 < 0 || rec[i].pos1 > N || t->leaf_count[li] < 0) return scht::fail(SCHT_ERR_INVALID_ARG, "scht_create: leaf range out of bounds");
        } else {
            int l = i + 1, r = rec[i].right;
            rec[i].size = 1 + rec[l].size + rec[r].size;
            if (l + rec[l].size != r) return scht::fail(SCHT_ERR_INVALID_ARG, "scht_create: nodes are not in DFS pre-order");
            if (rec[l].pos1 != rec[r].pos0) return scht::fail(SCHT_ERR_INVALID_ARG, "scht_create: leaves are not contiguous in DFS order");
            rec[i].pos0 = rec[l].pos0;
            rec[i].pos1 = rec[r].pos1;
            if (rec[l].depth != rec[i].depth + 1 || rec[r].depth != rec[i].depth + 1)
                return scht::fail(SCHT_ERR_INVALID_ARG, "scht_create: node_beta_off does not match node depth");
        }
        depth = std::max(depth, rec[i].depth);
    }
    if (rec[0].size != nn || rec[0].pos0 != 0 || rec[0].pos1 != N || rec[0].depth != 0)
        return scht::fail(SCHT_ERR_INVALID_ARG, "scht_create: tree does not cover the data set");

    scht_handle* h = new scht_handle();
    h->device = device;
    h->sm_count = prop.multiProcessorCount;
    h->max_smem = (int)prop.sharedMemPerBlockOptin;
    h->n_sm = std::max(1, prop.multiProcessorCount);
    h->depth = depth;
    h->h_leaf_start.assign(t->leaf_start, t->leaf_start + L);
    h->h_leaf_count.assign(t->leaf_count, t->leaf_count + L);
    if (const char* ev = getenv("SCHT_SCAN_VARIANT")) h->scan_variant = std::max(-1, std::min(3, atoi(ev)));
    auto fail_free = [&](int rc) {
        scht_destroy(h);
        return rc;
    };
#define TRY_H(x)                   \
    do {                           \
        int rc__ = (x);            \
        if (rc__) return fail_free(rc__); \
    } while (0)
#define CU_H(expr)                                                                                           \
    do {                                                                                                     \
        cudaError_t e__ = (expr);                                                                            \
        if (e__ != cudaSuccess)                                                                              \
            return fail_free(scht::fail(e__ == cudaErrorMemoryAllocation ? SCHT_ERR_OOM : SCHT_ERR_CUDA,    \
                                        std::string(#expr) + ": " + cudaGetErrorString(e__)));              \
    } while (0)
    CU_H(cudaSetDevice(device));
    CU_H(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    for (auto& e : h->ev) CU_H(cudaEventCreate(&e));

    DevTree& T = h->T;
    T.dim = dim;
    T.dpad = (dim + 3) & ~3;
    T.n_points = (int)N;
    T.n_pages = (int)((N + kPage - 1) / kPage);
    T.n_nodes = nn;
    T.n_leaves = L;
    const int dpad = T.dpad;

    // padded copies of the per-node vectors
    std::vector<float> nrm((size_t)nn * dpad, 0.f), lc((size_t)nn * dpad, 0.f), rc((size_t)nn * dpad, 0.f);
    for (int i = 0; i < nn; i++) {
        memcpy(&nrm[(size_t)i * dpad], t->node_normal + (size_t)i * dim, sizeof(float) * dim);
        memcpy(&lc[(size_t)i * dpad], t->node_left_center + (size_t)i * dim, sizeof(float) * dim);
        memcpy(&rc[(size_t)i * dpad], t->node_right_center + (size_t)i * dim, sizeof(float) * dim);
    }
    {
        // subtree jobs of the traversal kernel: the deepest level whose cut (nodes at that depth + shallower leaves) has <= kMaxCuts nodes
        int d0 = 0;
        auto count_cuts = [&](int d) {
            int c = 0;
            for (int i = 0; i < nn; i++) c += (rec[i].depth == d) || (rec[i].leaf >= 0 && rec[i].depth < d);
            return c;
        };
        while (d0 < depth && count_cuts(d0 + 1) <= kMaxCuts) d0++;
        std::vector<int> cuts;
        for (int i = 0; i < nn; i++)
            if (rec[i].depth == d0 || (rec[i].leaf >= 0 && rec[i].depth < d0)) cuts.push_back(i);
        std::vector<int> paths(cuts.size() * (size_t)(d0 + 1), -1);
        for (size_t c = 0; c < cuts.size(); c++) {
            int n = cuts[c], len = rec[n].depth + 1;
            for (int i = len - 1; i >= 0; i--) {
                paths[c * (size_t)(d0 + 1) + i] = n;
                n = t->node_parent[n];
            }
        }
        T.n_cuts = (int)cuts.size();
        T.cut_depth = d0;
        TRY_H(upload(h, paths.data(), paths.size(), &T.cut_path));
    }
    TRY_H(upload(h, rec.data(), rec.size(), &T.nodes));
    TRY_H(upload(h, t->node_beta, (size_t)t->node_beta_off[nn], &T.node_beta));
    TRY_H(upload(h, nrm.data(), nrm.size(), &T.node_normal));
    TRY_H(upload(h, lc.data(), lc.size(), &T.node_lc));
    TRY_H(upload(h, rc.data(), rc.size(), &T.node_rc));
    TRY_H(upload(h, t->leaf_start, (size_t)L, &T.leaf_start));
    TRY_H(upload(h, t->leaf_count, (size_t)L, &T.leaf_count));
    {
        // seed neighbourhood of every leaf: the smallest enclosing subtree with >= kSeedPoints points
        std::vector<int> ss(L), se(L);
        // (512 points, about one tree level above a 500-point leaf.  Against 1024: 1Mx16 step 11.41 -> 11.29 ms, Gaussian clusters
        // 2Mx32 47.3 -> 46.2 ms, SIFT-like 1Mx128 46.7 -> 44.0 ms; 256: 11.21 ms on 1Mx16, 2048: 11.74 ms)
        int seed_points = kSeedPoints;
        if (const char* ev = getenv("SCHT_SEED_POINTS")) seed_points = std::max(1, atoi(ev));
        for (int i = 0; i < nn; i++) {
            if (rec[i].leaf < 0) continue;
            int a = i;
            while (rec[a].pos1 - rec[a].pos0 < seed_points && t->node_parent[a] >= 0) a = t->node_parent[a];
            ss[rec[i].leaf] = rec[a].pos0;
            se[rec[i].leaf] = rec[a].pos1;
        }
        TRY_H(upload(h, ss.data(), ss.size(), &T.seed_start));
        TRY_H(upload(h, se.data(), se.size(), &T.seed_end));
    }
    {
        std::vector<int32_t> oi((size_t)T.n_pages * kPage, -1);
        memcpy(oi.data(), t->sorted_to_orig, sizeof(int32_t) * (size_t)N);
        TRY_H(upload(h, oi.data(), oi.size(), &T.orig_idx));
        CU_H(cudaStreamSynchronize(h->stream));
    }
    // max ||x||_2 (double, rounded up) for the pruning margin
    {
        double mx = 0;
        for (int64_t i = 0; i < N; i++) {
            double s = 0;
            const float* p = t->sorted_points + (size_t)i * dim;
            for (int j = 0; j < dim; j++) s += (double)p[j] * p[j];
            mx = std::max(mx, s);
        }
        T.xnorm_max = (float)(std::sqrt(mx) * 1.0001);
        if (!(T.xnorm_max < 3.0e38f)) T.xnorm_max = 3.0e38f;
    }
    // pages: upload row-major slabs, transpose on the device (and keep a padded row-major copy)
    {
        float* pages = nullptr;
        size_t pbytes = (size_t)T.n_pages * dpad * kPage * 4;
        CU_H(cudaMalloc(&pages, pbytes));
        h->owned.push_back(pages);
        h->tree_bytes += pbytes;
        T.pages = pages;
        // wide points (several 16 KB slices per page): the same padded values row-major, for the exact re-evaluation of single
        // points.  Narrow points are re-evaluated from the stage in shared memory (k_scan) or the L2-resident pages (k_scan_tc)
        float* rows = nullptr;
        if (dpad > kRowsPerStage) {
            CU_H(cudaMalloc(&rows, pbytes));
            h->owned.push_back(rows);
            h->tree_bytes += pbytes;
        }
        T.rows = rows;
        const int64_t slab = 4 << 20;  // rows per slab
        DevBuf tmp;
        CU_H(tmp.reserve((size_t)std::min<int64_t>(slab, (int64_t)T.n_pages * kPage) * dim * 4));
        const int64_t total_rows = (int64_t)T.n_pages * kPage;
        for (int64_t r0 = 0; r0 < total_rows; r0 += slab) {
            int64_t nr = std::min<int64_t>(slab, total_rows - r0);
            int64_t real = std::max<int64_t>(0, std::min<int64_t>(nr, N - r0));
            if (real > 0) CU_H(cudaMemcpyAsync(tmp.p, t->sorted_points + (size_t)r0 * dim, (size_t)real * dim * 4, cudaMemcpyHostToDevice, h->stream));
            int64_t work = nr * dpad;
            k_build_pages<<<(unsigned)((work + 255) / 256), 256, 0, h->stream>>>(tmp.as<float>(), r0, nr, dim, dpad, N, pages, rows);
            CU_H(cudaGetLastError());
            CU_H(cudaStreamSynchronize(h->stream));
        }
        tmp.release();
    }
    // tensor-core page image (centred, TF32-rounded coordinates + |p'|^2 hi/lo), see scht_tc.cuh
    if (const char* ev = getenv("SCHT_TRAVERSE_SAMPLING")) h->traverse_sampling = atoi(ev) != 0;
    if (const char* ev = getenv("SCHT_SAMPLE_REUSE")) h->sample_reuse = std::max(0, atoi(ev));
    if (const char* ev = getenv("SCHT_TC")) h->tc_mode = std::max(0, std::min(2, atoi(ev)));
    if (const char* ev = getenv("SCHT_TC_SPLIT")) h->tc_split = std::max(0, atoi(ev));
    if (const char* ev = getenv("SCHT_TC_ISSUERS")) h->tc_issuers = atoi(ev);
    if (const char* ev = getenv("SCHT_SCAN_STAGES")) h->scan_stages = atoi(ev);
    if (const char* ev = getenv("SCHT_TC_N256")) h->tc_n256 = atoi(ev);
    {
        const int kp = (dim + 3 + 15) & ~15;
        const size_t tbytes = (size_t)T.n_pages * tc_page_bytes(kp);
        size_t free_b = 0, total_b = 0;
        cudaMemGetInfo(&free_b, &total_b);
        if (h->tc_mode != 0 && kp <= 1024 && tbytes < free_b / 2) {
            std::vector<double> sum(dim, 0.0);
            for (int64_t i = 0; i < N; i++) {
                const float* p = t->sorted_points + (size_t)i * dim;
                for (int j = 0; j < dim; j++) sum[j] += p[j];
            }
            std::vector<float> mean(dpad, 0.f), pmax(dpad, 0.f);
            for (int j = 0; j < dim; j++) mean[j] = (float)(sum[j] / (double)N);
            double n2max = 0;
            std::vector<double> amax(dim, 0.0);
            for (int64_t i = 0; i < N; i++) {
                const float* p = t->sorted_points + (size_t)i * dim;
                double s2 = 0;
                for (int j = 0; j < dim; j++) {
                    double c = (double)p[j] - (double)mean[j];
                    amax[j] = std::max(amax[j], std::fabs(c));
                    s2 += c * c;
                }
                n2max = std::max(n2max, s2);
            }
            double amax_all = 0, asum = 0;
            for (int j = 0; j < dim; j++) {
                pmax[j] = (float)(amax[j] * 1.0001);
                amax_all = std::max(amax_all, amax[j]);
                asum += amax[j];
            }
            h->pnorm_max = (float)(std::sqrt(n2max) * 1.0001);
            h->pabs_sum = (float)(asum * 1.0001);
            // power-of-two scale that puts the largest centred coordinate into (128, 256]: well inside FP16's range
            h->tc_scale = (amax_all > 0 && std::isfinite(amax_all)) ? (float)std::exp2(std::floor(std::log2(256.0 / amax_all))) : 1.f;
            const double sn = (double)h->tc_scale * h->tc_scale * n2max;
            // norm slots hold s^2|p'|^2 / u with u a power of two that keeps them below 16384 (FP16 range, room for 60000 on padding)
            h->norm_unit = (sn > 16384.0) ? (float)std::exp2(std::ceil(std::log2(sn / 16384.0))) : 1.f;
            if (std::isfinite(h->pnorm_max) && std::isfinite(h->tc_scale) && h->tc_scale > 1e-30f && h->tc_scale < 1e30f && sn < 1e37) {
                TRY_H(upload(h, mean.data(), mean.size(), &h->d_mean));
                TRY_H(upload(h, pmax.data(), pmax.size(), &h->d_pmax));
                unsigned char* tcp = nullptr;
                CU_H(cudaMalloc(&tcp, tbytes));
                h->owned.push_back(tcp);
                h->tree_bytes += tbytes;
                const int64_t npos = (int64_t)T.n_pages * kPage;
                k_build_tcpages<<<(unsigned)((npos + 255) / 256), 256, 0, h->stream>>>(T, h->d_mean, h->tc_scale, h->norm_unit, kp, tcp);
                CU_H(cudaGetLastError());
                h->tcpages = tcp;
                h->kp = kp;
            }
        }
    }
    CU_H(cudaStreamSynchronize(h->stream));
#undef TRY_H
#undef CU_H
    *out = h;
    return SCHT_OK;
}

extern "C" void scht_destroy(scht_handle* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    for (void* p : h->owned) cudaFree(p);
    for (DevBuf* b : {&h->q, &h->q_leaf, &h->hist, &h->q_order, &h->state, &h->ranges, &h->n_ranges, &h->idx, &h->d2, &h->dist, &h->counters, &h->seg_keys})
        b->release();
    for (auto& e : h->ev)
        if (e) cudaEventDestroy(e);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

extern "C" int scht_set_batch_size(scht_handle* h, int64_t batch_size) {
    if (!h || batch_size < 1) return scht::fail(SCHT_ERR_INVALID_ARG, "scht_set_batch_size: batch_size must be >= 1");
    h->batch_size = batch_size;
    return SCHT_OK;
}

extern "C" int scht_query(scht_handle* h, const float* queries, int64_t n_queries, int32_t dim, int32_t K, int32_t* idx_out,
                          float* dist2_out, float* dist_out, scht_stats* stats) {
    return host_query(h, queries, n_queries, dim, K, idx_out, dist2_out, dist_out, stats, false);
}

extern "C" int scht_brute_force(scht_handle* h, const float* queries, int64_t n_queries, int32_t dim, int32_t K, int32_t* idx_out,
                                float* dist2_out, float* dist_out, scht_stats* stats) {
    return host_query(h, queries, n_queries, dim, K, idx_out, dist2_out, dist_out, stats, true);
}

extern "C" int scht_query_device(scht_handle* h, const float* d_queries, int64_t n_queries, int32_t dim, int32_t K, int32_t* d_idx_out,
                                 float* d_dist2_out, float* d_dist_out, void* cuda_stream, scht_stats* stats) {
    int rc = check_query_args(h, d_queries, n_queries, dim, K, d_idx_out, d_dist2_out);
    if (rc) return rc;
    if (stats) memset(stats, 0, sizeof(*stats));
    double t0 = now_ms();
    CU_TRY(cudaSetDevice(h->device));
    cudaStream_t s = cuda_stream ? (cudaStream_t)cuda_stream : h->stream;
    int64_t bs = std::min<int64_t>(h->batch_size, 0x7fffffffLL / std::max(K, dim));
    for (int64_t b0 = 0; b0 < n_queries; b0 += bs) {
        int nq = (int)std::min<int64_t>(bs, n_queries - b0);
        QueryOut out;
        out.idx = d_idx_out + (size_t)b0 * K;
        out.d2 = d_dist2_out + (size_t)b0 * K;
        out.dist = d_dist_out ? d_dist_out + (size_t)b0 * K : nullptr;
        rc = run_batch(h, d_queries + (size_t)b0 * dim, nq, K, false, false, 0, h->T.n_leaves, out, s, stats, stats != nullptr);
        if (rc) return rc;
    }
    CU_TRY(cudaStreamSynchronize(s));
    if (stats) stats->ms_total = now_ms() - t0;
    return SCHT_OK;
}

extern "C" int scht_find_approximate_leaves(scht_handle* h, const float* queries, int64_t n_queries, int32_t dim, int32_t* leaf_out) {
    if (!h) return scht::fail(SCHT_ERR_INVALID_ARG, "handle is NULL");
    if (dim != h->T.dim) return scht::fail(SCHT_ERR_DIM_MISMATCH, "query dimension != tree dimension");
    if (n_queries < 0 || n_queries > 0x7fffffffLL / std::max(1, dim) || (n_queries > 0 && (!queries || !leaf_out)))
        return scht::fail(SCHT_ERR_INVALID_ARG, "scht_find_approximate_leaves: bad arguments");
    if (n_queries == 0) return SCHT_OK;
    CU_TRY(cudaSetDevice(h->device));
    int nq = (int)n_queries;
    CU_TRY(h->q.reserve((size_t)nq * dim * 4));
    CU_TRY(h->q_leaf.reserve((size_t)nq * 4));
    CU_TRY(cudaMemcpyAsync(h->q.p, queries, (size_t)nq * dim * 4, cudaMemcpyHostToDevice, h->stream));
    k_descend<<<(nq + 127) / 128, 128, 0, h->stream>>>(h->T, h->q.as<float>(), nq, h->q_leaf.as<int>(), nullptr, nullptr);
    CU_TRY(cudaGetLastError());
    CU_TRY(cudaMemcpyAsync(leaf_out, h->q_leaf.p, (size_t)nq * 4, cudaMemcpyDeviceToHost, h->stream));
    CU_TRY(cudaStreamSynchronize(h->stream));
    return SCHT_OK;
}

extern "C" int scht_query_partial_device(scht_handle* h, const float* d_queries, int64_t n_queries, int32_t dim, int32_t K,
                                         int32_t leaf_begin, int32_t leaf_end, uint64_t* d_keys_out, void* cuda_stream, scht_stats* stats) {
    int rc = check_query_args(h, d_queries, n_queries, dim, K, d_keys_out, d_keys_out);
    if (rc) return rc;
    if (leaf_begin < 0 || leaf_end > h->T.n_leaves || leaf_begin > leaf_end) return scht::fail(SCHT_ERR_INVALID_ARG, "bad leaf range");
    if (stats) memset(stats, 0, sizeof(*stats));
    if (n_queries == 0) return SCHT_OK;
    double t0 = now_ms();
    CU_TRY(cudaSetDevice(h->device));
    cudaStream_t s = cuda_stream ? (cudaStream_t)cuda_stream : h->stream;
    int64_t bs = std::min<int64_t>(h->batch_size, 0x7fffffffLL / std::max(K, dim));
    for (int64_t b0 = 0; b0 < n_queries; b0 += bs) {
        int nq = (int)std::min<int64_t>(bs, n_queries - b0);
        QueryOut out;
        out.keys = d_keys_out + (size_t)b0 * K;
        rc = run_batch(h, d_queries + (size_t)b0 * dim, nq, K, false, true, leaf_begin, leaf_end, out, s, stats, stats != nullptr);
        if (rc) return rc;
    }
    CU_TRY(cudaStreamSynchronize(s));
    if (stats) stats->ms_total = now_ms() - t0;
    return SCHT_OK;
}

extern "C" int scht_merge_partial_device(scht_handle* h, const uint64_t* d_keys, int32_t n_lists, const float* d_queries, int64_t n_queries,
                                         int32_t dim, int32_t K, int32_t* d_idx_out, float* d_dist2_out, float* d_dist_out, void* cuda_stream) {
    (void)d_queries;
    int rc = check_query_args(h, d_keys, n_queries, dim, K, d_idx_out, d_dist2_out);
    if (rc) return rc;
    if (n_lists < 1 || n_lists > 16) return scht::fail(SCHT_ERR_INVALID_ARG, "n_lists must be in [1,16]");
    if (n_queries == 0) return SCHT_OK;
    CU_TRY(cudaSetDevice(h->device));
    cudaStream_t s = cuda_stream ? (cudaStream_t)cuda_stream : h->stream;
    int nq = (int)n_queries;
    k_merge_partial<<<(nq + 127) / 128, 128, 0, s>>>(h->T, d_keys, n_lists, nq, K, d_idx_out, d_dist2_out, d_dist_out);
    CU_TRY(cudaGetLastError());
    CU_TRY(cudaStreamSynchronize(s));
    return SCHT_OK;
}

extern "C" int scht_handle_info(const scht_handle* h, int32_t* dim, int64_t* n_points, int32_t* n_nodes, int32_t* n_leaves, int32_t* device,
                                int64_t* device_bytes) {
    if (!h) return scht::fail(SCHT_ERR_INVALID_ARG, "handle is NULL");
    if (dim) *dim = h->T.dim;
    if (n_points) *n_points = h->T.n_points;
    if (n_nodes) *n_nodes = h->T.n_nodes;
    if (n_leaves) *n_leaves = h->T.n_leaves;
    if (device) *device = h->device;
    if (device_bytes) *device_bytes = (int64_t)h->tree_bytes;
    return SCHT_OK;
}
