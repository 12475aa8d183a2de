// Query path of the C ABI (include/scht.h): the batched exact k-NN query on one GPU.
//
// Replaces GPU_Convex_Tree<T>'s OpenCL host code (reference: Code/GPU_Convex_Tree.h:208-323 do_kNN, :328-427
// do_brute_force_kNN, :168-193 do_find_approximate_nodes).  One plain handle = one GPU; the create path lives in
// scht_create.cu, the multi-GPU layer in scht_multi.cu.  There is no CPU fallback: without an sm_100 device every entry
// point fails.
//
// One batch = two phases:
//   batch_seed  K1 descent + leaf histogram, K3 bucketing, K4a exact seed scan of the approximate leaves (lists -> state)
//   batch_main  eligibility of the tensor-core prefilter, page selection (scan-page balls) or tree traversal
//               (hyper-plane bounds), K4b main scan (k_scan_tc or k_scan), merge of page segments
// Between the phases the sharded-dataset mode exchanges the seed bounds of the shards, and the brute-force entry
// turns the seed lists into a bound (its keys carry original indices, so the lists start empty).
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "scht_handle.h"
#include "scht_kernels.cuh"
#include "scht_select.cuh"
#include "scht_tc.cuh"

using namespace scht;

namespace {

// Leaf-scan kernel variants: consumer warps per CTA x resident CTAs per SM the register budget is sized for.
//   0: 8 warps (64 queries/tile), 1 CTA/SM   1: 7 warps (56 queries/tile, 256 threads with the producer), 2 CTAs/SM
//   2: 4 warps (32 queries/tile), 3 CTAs/SM   3: the same with 96 registers and a 2-stage ring, 4 CTAs/SM
// auto: 2, and 3 for the SEED scan of narrow points with short lists (1.50 -> 1.41 ms on 1Mx16; no gain for wide points or
// K = 100; the FP32-bound main scan keeps its 128 registers)
constexpr int kVariantWarps[4] = {8, 7, 4, 4};
int pick_variant(const scht_handle* h, int K) {
    int v = h->opt.scan_variant;
    if (v < 0) v = 2;  // auto (the seed scan of narrow points runs as variant 3, see batch_seed)
    int nw = kVariantWarps[v];
    if ((int)scan_smem_bytes(nw, h->T.dpad, K) <= h->max_smem) return v;
    if ((int)scan_smem_bytes(4, h->T.dpad, K) <= h->max_smem) return 2;
    return -1;
}

template <int NW, int MINB, bool FULL, int STAGES, bool LAZY>
int launch_scan_l(scht_handle* h, const ScanParams& P, cudaStream_t s) {
    size_t smem = scan_smem_bytes(NW, P.T.dpad, P.K, STAGES);
    CU_TRY(cudaFuncSetAttribute(k_scan<NW, MINB, FULL, STAGES, LAZY>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int tq = NW * kQPerWarp;
    int n_tiles = (P.nq + tq - 1) / tq;
    k_scan<NW, MINB, FULL, STAGES, LAZY><<<n_tiles, (NW + 1) * 32, smem, s>>>(P);
    CU_TRY(cudaGetLastError());
    return SCHT_OK;
}
template <int NW, int MINB, bool FULL, int STAGES = kStages>
int launch_scan_t(scht_handle* h, const ScanParams& P, cudaStream_t s) {
    if (P.lazy_seed && P.mode == kScanSeed) return launch_scan_l<NW, MINB, FULL, STAGES, true>(h, P, s);
    return launch_scan_l<NW, MINB, FULL, STAGES, false>(h, P, s);
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
            const bool shallow = 3 * scan_smem_bytes(4, P.T.dpad, P.K) > per_sm && 3 * scan_smem_bytes(4, P.T.dpad, P.K, 2) <= per_sm && h->opt.scan_stages != 4;
            if (shallow || h->opt.scan_stages == 2) return full ? launch_scan_t<4, 3, true, 2>(h, P, s) : launch_scan_t<4, 3, false, 2>(h, P, s);
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

// page selection by bounding balls (scht_select.cuh); the query lives in registers up to 128 dimensions
int select_regs(int dpad) { return dpad <= 16 ? 16 : dpad <= 32 ? 32 : dpad <= 64 ? 64 : dpad <= 128 ? 128 : 0; }
bool select_fits(const scht_handle* h) { return (int)select_smem_bytes(select_regs(h->T.dpad), h->T.dpad, h->I.n_supers) <= h->max_smem; }
int launch_select(scht_handle* h, const SelectParams& P, cudaStream_t s) {
    const int n_tiles = (P.nq + kSelQ - 1) / kSelQ;
    const int blocks = selection_size(n_tiles, P.sel_mode, P.sel_step);
    if (blocks == 0) return SCHT_OK;
    const int dp = select_regs(P.T.dpad);
    const size_t smem = select_smem_bytes(dp, P.T.dpad, P.I.n_supers);
    auto go = [&](auto kernel) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) kernel<<<blocks, kSelQ, smem, s>>>(P);
        return e;
    };
    if (dp == 16) CU_TRY(go(k_select_pages<16>));
    else if (dp == 32) CU_TRY(go(k_select_pages<32>));
    else if (dp == 64) CU_TRY(go(k_select_pages<64>));
    else if (dp == 128) CU_TRY(go(k_select_pages<128>));
    else CU_TRY(go(k_select_pages<0>));
    CU_TRY(cudaGetLastError());
    return SCHT_OK;
}

void fill_scan_params(scht_handle* h, ScanParams& P, const float* d_q, int nq, int K) {
    P = ScanParams{};
    P.T = h->T;
    P.q = d_q;
    P.q_order = h->q_order.as<int>();
    P.q_leaf = h->q_leaf.as<int>();
    P.nq = nq;
    P.K = K;
    P.pos_lo = h->T.pos_lo;
    P.pos_hi = h->T.pos_hi;
    P.lazy_seed = h->bs.lazy_seed;
    P.range_groups = h->T.n_cuts;
    P.state = h->state.as<uint64_t>();
    P.ranges = h->ranges.as<int2>();
    P.n_ranges = h->n_ranges.as<int>();
    P.counters = h->counters.as<unsigned long long>();
}

void leaf_range_positions(const scht_handle* h, int leaf_lo, int leaf_hi, int& pos_lo, int& pos_hi) {
    pos_lo = (leaf_lo < h->T.n_leaves) ? h->h_leaf_start[leaf_lo] : h->T.n_points;
    pos_hi = (leaf_hi > leaf_lo) ? h->h_leaf_start[leaf_hi - 1] + h->h_leaf_count[leaf_hi - 1] : pos_lo;
    pos_lo = std::max(pos_lo, h->T.pos_lo);
    pos_hi = std::max(pos_lo, std::min(pos_hi, h->T.pos_hi));
}

}  // namespace

namespace scht {

double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

// Phase 1 of a batch.  kind: kBatchTree / kBatchPartial / kBatchBrute (the brute-force entry seeds through the tree too: the
// seed lists only provide its bound).  The seed scan covers the handle's resident positions.
int batch_seed(scht_handle* h, const float* d_q, int nq, int K, int kind, int leaf_lo, int leaf_hi, cudaStream_t s, scht_stats* st, bool timed,
               bool lazy) {
    (void)leaf_lo;
    (void)leaf_hi;
    (void)st;
    const int variant = pick_variant(h, K);
    if (variant < 0) return scht::fail(SCHT_ERR_INVALID_ARG, "dimension * K too large for the leaf-scan kernel's shared memory");
    const DevTree& T = h->T;
    h->bs.variant = variant;
    h->bs.tq = kVariantWarps[variant] * kQPerWarp;
    h->bs.launches = 0;
    h->bs.scans = 0;
    h->bs.kind = kind;
    h->bs.nq = nq;
    h->bs.K = K;
    h->bs.lazy_seed = lazy ? 1 : 0;
    const int n_tiles = (nq + h->bs.tq - 1) / h->bs.tq;
    const int n_tiles_tc = (nq + kTcQ - 1) / kTcQ;
    CU_TRY(h->q_leaf.reserve((size_t)nq * 4));
    CU_TRY(h->q_order.reserve((size_t)nq * 4));
    CU_TRY(h->hist.reserve((size_t)(T.n_leaves + 1) * 4));
    CU_TRY(h->counters.reserve(16 * 8));
    CU_TRY(h->state.reserve((size_t)nq * K * 8));
    CU_TRY(h->ranges.reserve((size_t)std::max(n_tiles, n_tiles_tc) * kRangeCap * sizeof(int2)));
    CU_TRY(h->n_ranges.reserve((size_t)std::max(n_tiles, n_tiles_tc) * kMaxCuts * 4));
    CU_TRY(cudaMemsetAsync(h->counters.p, 0, 128, s));
    unsigned long long* ctr = h->counters.as<unsigned long long>();
    if (timed) CU_TRY(cudaEventRecord(h->ev[0], s));
    // K1 descend + leaf histogram, K3 bucketing
    CU_TRY(cudaMemsetAsync(h->hist.p, 0, (size_t)(T.n_leaves + 1) * 4, s));
    k_descend<<<(nq + 127) / 128, 128, 0, s>>>(T, d_q, nq, h->q_leaf.as<int>(), h->hist.as<int>(), ctr);
    k_leaf_prefix_sum<<<1, 1024, 0, s>>>(h->hist.as<int>(), T.n_leaves);
    k_scatter<<<(nq + 255) / 256, 256, 0, s>>>(h->q_leaf.as<int>(), T.leaf_rank, nq, h->hist.as<int>(), h->q_order.as<int>());
    CU_TRY(cudaGetLastError());
    h->bs.launches += 3;
    if (timed) CU_TRY(cudaEventRecord(h->ev[1], s));
    // K4a seed scan of the approximate leaves
    ScanParams P;
    fill_scan_params(h, P, d_q, nq, K);
    P.mode = kScanSeed;
    P.write_mode = kWriteState;
    const int seed_variant = (h->opt.scan_variant < 0 && variant == 2 && T.dpad <= kRowsPerStage && K <= 32) ? 3 : variant;  // same 32-query tiles
    int rc = launch_scan(h, seed_variant, P, s);
    if (rc) return rc;
    h->bs.launches++;
    h->bs.scans++;
    if (timed) CU_TRY(cudaEventRecord(h->ev[2], s));
    return SCHT_OK;
}

int batch_bounds(scht_handle* h, int nq, int K, float* d_bound_out, cudaStream_t s) {
    k_extract_bounds<<<(nq + 255) / 256, 256, 0, s>>>(h->state.as<uint64_t>(), h->q_order.as<int>(), nq, K, d_bound_out);
    CU_TRY(cudaGetLastError());
    h->bs.launches++;
    return SCHT_OK;
}

int min_bounds(scht_handle* h, float* const* d_bounds, int n, int nq, float* d_out, cudaStream_t s) {
    (void)h;
    BoundPtrs B{};
    for (int i = 0; i < n; i++) B.p[i] = d_bounds[i];
    k_min_bounds<<<(nq + 255) / 256, 256, 0, s>>>(B, n, nq, d_out);
    CU_TRY(cudaGetLastError());
    return SCHT_OK;
}

int merge_lists(scht_handle* h, const uint64_t* d_keys, int n_lists, int nq, int K, bool brute_keys, int32_t* d_idx, float* d_d2, float* d_dist,
                cudaStream_t s) {
    k_merge_partial<<<(nq + 127) / 128, 128, 0, s>>>(h->T, d_keys, n_lists, nq, K, brute_keys ? 1 : 0, d_idx, d_d2, d_dist);
    CU_TRY(cudaGetLastError());
    return SCHT_OK;
}

// Phase 2 of a batch.  d_ext_bound (by query id, may be null): upper bound of the K-th distance^2 known from elsewhere.
//   kBatchTree     lists resume from the seed state; results -> out.idx / d2 / dist
//   kBatchPartial  lists resume from the seed state clipped to the leaves [leaf_lo, leaf_hi); keys -> out.keys / out.key_tab
//   kBatchBrute    lists start empty, keys carry the original index; results -> out.idx ... (or keys when out.keys / key_tab is set)
int batch_main(scht_handle* h, const float* d_q, int nq, int K, int kind, int leaf_lo, int leaf_hi, const float* d_ext_bound, const QueryOut& out,
               cudaStream_t s, scht_stats* st, bool timed) {
    const DevTree& T = h->T;
    const int variant = h->bs.variant, tq = h->bs.tq;
    int& launches = h->bs.launches;
    int& scans = h->bs.scans;
    unsigned long long* ctr = h->counters.as<unsigned long long>();
    const bool brute = kind == kBatchBrute, partial = kind == kBatchPartial;
    const bool keys_wanted = out.keys != nullptr || out.n_slices > 0;
    bool used_tc = false;

    ScanParams P;
    fill_scan_params(h, P, d_q, nq, K);
    P.idx_out = out.idx;
    P.d2_out = out.d2;
    P.dist_out = out.dist;
    P.keys_out = out.keys;
    P.n_slices = out.n_slices;
    P.list = out.list;
    for (int i = 0; i < kMaxShards; i++) P.key_tab[i] = out.key_tab[i];
    for (int i = 0; i <= kMaxShards; i++) P.slice_lo[i] = out.slice_lo[i];
    P.ext_bound = d_ext_bound;
    if (partial) {
        leaf_range_positions(h, leaf_lo, leaf_hi, P.pos_lo, P.pos_hi);
        if (P.pos_lo > T.pos_lo || P.pos_hi < T.pos_hi) {  // seed lists may hold positions outside the leaf range
            k_clip_state<<<(nq + 255) / 256, 256, 0, s>>>(h->state.as<uint64_t>(), nq, K, P.pos_lo, P.pos_hi);
            launches++;
        }
    } else if (brute) {
        k_clip_state<<<(nq + 255) / 256, 256, 0, s>>>(h->state.as<uint64_t>(), nq, K, 0, 0);  // empty lists: brute-force keys differ
        launches++;
    }
    CU_TRY(cudaGetLastError());

    // tensor-core prefilter or exact FP32 scan for the main pass?
    const ScanImage& I = h->I;
    bool use_tc = false;
    int tc_stages = kTcMaxStages;
    while (tc_stages > 3 && (int)tc_smem_bytes(T.dpad, I.kp, K, tc_stages) > h->max_smem) tc_stages--;
    if (I.kp > 0 && h->opt.tc_mode != 0 && (!brute || h->opt.brute_tc) && (int)tc_smem_bytes(T.dpad, I.kp, K, tc_stages) <= h->max_smem) {
        const long long esig = ((long long)K << 1) | ((long long)kind << 9) | ((long long)(d_ext_bound ? 1 : 0) << 12) | ((long long)std::min(nq >> 10, 1 << 20) << 16);
        if (h->opt.tc_mode == 1) {
            use_tc = true;
        } else if (h->elig_skip > 0 && esig == h->elig_sig) {
            // the last batches of this shape all got the same verdict: no kernel, no host round trip (a wrong guess only costs
            // time: queries the prefilter cannot handle pass everything to the exact path)
            h->elig_skip--;
            use_tc = h->elig_last != 0;
        } else {
            k_tc_eligibility<<<(nq + 127) / 128, 128, 0, s>>>(T, I, d_q, P.q_order, nq, K, P.state, d_ext_bound, ctr);
            launches++;
            unsigned long long ok = 0;
            CU_TRY(cudaMemcpyAsync(&ok, ctr + 6, 8, cudaMemcpyDeviceToHost, s));
            CU_TRY(cudaStreamSynchronize(s));
            use_tc = ok * 10 >= (unsigned long long)nq * 9;  // >= 90 % of the queries have a tight margin
            if (esig != h->elig_sig || (int)use_tc != h->elig_last) h->elig_streak = 0;
            if (++h->elig_streak >= 2) h->elig_skip = h->opt.sample_reuse;
            h->elig_sig = esig;
            h->elig_last = use_tc ? 1 : 0;
        }
    }
    const int tq_main = use_tc ? kTcQ : tq;
    const int n_tiles_main = (nq + tq_main - 1) / tq_main;
    const int page_lo = use_tc ? 0 : P.pos_lo / kPage;
    const int page_hi = use_tc ? I.n_spages : (P.pos_hi + kPage - 1) / kPage;
    const int n_list_pages = std::max(1, page_hi - page_lo);
    bool pruned = false;  // the page lists are (much) shorter than "every page"
    const float* sel_bound = nullptr;  // per query: the bound k_select_pages tested with (prefilter scan with page selection)

    // ---- the prefilter scan kernel over the current page lists (used for the first pass of a pruned batch and for the main scan)
    auto launch_tc_scan = [&](int write_mode, bool allow_split) -> int {
        TcScanParams TC{};
        TC.S = P;
        if (sel_bound) TC.S.ext_bound = sel_bound;
        TC.S.mode = kScanList;
        TC.S.write_mode = write_mode;
        TC.S.range_groups = kMaxCuts;
        TC.I = I;
        TC.brute = brute ? 1 : 0;
        TC.n_stages = tc_stages;
        TC.rescorer_sleep_ns = h->opt.rescorer_sleep_ns;
        TC.ring_wait_ns = h->opt.ring_wait_ns;
        // Tail effect: n_tiles CTAs of equal length on n_sm SMs take ceil(n_tiles/n_sm) rounds (782 tiles on 148 SMs: 6 rounds
        // for 5.28 rounds of work).  Splitting every tile's page list into S segments (own CTA each, partial lists merged by
        // k_merge_partial) makes the rounds S times shorter: pick the S <= 8 with the smallest ceil(S*n_tiles/n_sm)/S, charging
        // 5 % per extra segment for the repeated set-up and the looser early bounds (measured).  Also what keeps a SMALL batch from
        // running on a handful of SMs.  Not with pruned page lists: their pages are not spread over the page index range.
        int n_seg = 1;
        if (allow_split && write_mode == kWriteFinal && h->opt.tc_split && !(pruned && h->opt.tc_split == 1 && n_tiles_main >= h->n_sm)) {
            double best = 1e30;
            for (int S = 1; S <= 8; S++) {
                if (S > 1 && (page_hi - page_lo) / S < 64) break;
                const double rounds = std::ceil((double)S * n_tiles_main / h->n_sm) / S * (1.0 + 0.05 * (S - 1));
                if (rounds < best - 1e-9) {
                    best = rounds;
                    n_seg = S;
                }
            }
            if (h->opt.tc_split > 1) n_seg = std::min(8, h->opt.tc_split);  // forced
            if (n_seg > 1) {
                CU_TRY(h->seg_keys.reserve((size_t)n_seg * nq * K * sizeof(uint64_t)));
                TC.S.keys_out = h->seg_keys.as<uint64_t>();
                TC.S.n_slices = 0;
                TC.seg_page0 = page_lo;
                TC.seg_pages = (page_hi - page_lo + n_seg - 1) / n_seg;
            }
        }
        TC.n_seg = n_seg;
        // Survivor rings: SHALLOW for short lists.  A survivor waits in its ring while the bound it was tested against goes stale, and
        // on clustered data the entries queued behind it are mostly pairs a fresher bound would have dropped: with 64 instead of 256
        // entries per epilogue warp the 10Mx32 Gaussian clusters need 125 instead of 147 exact evaluations per query (main scan
        // 30.2 -> 27.2 ms), SIFT-like 144 instead of 171 (25.7 -> 24.8 ms per step); 512 entries: 166 evaluations, 33.2 ms.  A full
        // ring stalls its epilogue warp, and with it the pipeline, until the rescorer catches up -- that back-pressure is what keeps
        // the bounds fresh.  Uniform 16-D data never fills a ring (10.3 ms either way); K = 100 prefers 256 (33.4 vs 34.1 ms).
        int queue_cap = h->opt.tc_ring > 0 ? h->opt.tc_ring : (K <= 32 ? 64 : kTcQueueCap);
        if ((int)tc_smem_bytes(T.dpad, I.kp, K, tc_stages, queue_cap) > h->max_smem) queue_cap = kTcQueueCap;
        TC.queue_cap = queue_cap;
        TC.recheck = (h->opt.tc_recheck && (long long)I.n_slots <= (1ll << kTcSlotBits)) ? 1 : 0;
        size_t smem = tc_smem_bytes(T.dpad, I.kp, K, tc_stages, queue_cap);
        // One K slice per page (narrow points): two issuers, one per page parity, each issuing the page's two N=128 half-page
        // chains (+3 % over four issuers; whole-page N=256 chains cost 2.4 % there).  Several slices per page (wide points): the
        // MMA time counts, and one N=256 chain per page (128 instead of 2 x 92 cycles per K16 step) beats four issuers of
        // half-page chains by 2 % (D=128) to 5 % (D=32).
        const bool wide = I.kp > kTcKSlice;
        bool want_n256 = (h->opt.tc_n256 < 0) ? wide : (h->opt.tc_n256 != 0);
        int n_iss = (wide && !want_n256) ? 4 : 2;
        if (h->opt.tc_issuers == 2 || h->opt.tc_issuers == 4) n_iss = h->opt.tc_issuers;
        // a page with at least as many K slices as stages: one issuer takes every page (TcScanParams::single_issuer)
        const int n_slices = (I.kp + kTcKSlice - 1) / kTcKSlice;
        if (n_slices >= tc_stages) {
            n_iss = 2;
            TC.single_issuer = 1;
        }
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
        launches++;
        scans++;
        if (n_seg > 1) {
            k_merge_partial<<<(nq + 127) / 128, 128, 0, s>>>(T, h->seg_keys.as<uint64_t>(), n_seg, nq, K, brute ? 1 : 0, P.idx_out, P.d2_out, P.dist_out);
            CU_TRY(cudaGetLastError());
            launches++;
        }
        return SCHT_OK;
    };

    if (brute && !use_tc) {
        // exact FP32 exhaustive scan (k_scan walks every resident page itself)
    } else {
        // Which pages does a tile scan?  prefilter scan: ball test of the scan pages (k_select_pages); exact scan: tree traversal
        // with the hyper-plane bounds (k_traverse).  Probed first on a sample (about 24 tiles, at least every 64th and at most every
        // 8th; a small batch: all tiles): when the bounds keep most pages anyway they are not evaluated for the other tiles
        // (every page is listed).
        const bool can_select = use_tc && h->opt.page_select && select_fits(h);
        SelectParams SP{};
        TraverseParams TP{};
        if (use_tc) {
            SP.T = T;
            SP.I = I;
            SP.q = d_q;
            SP.q_order = P.q_order;
            SP.nq = nq;
            SP.K = K;
            SP.state = P.state;
            SP.ext_bound = d_ext_bound;
            SP.ranges = h->ranges.as<int2>();
            SP.n_ranges = h->n_ranges.as<int>();
            SP.counters = ctr;
            SP.tighten = 1;
            // the bound every query's test used goes to the scan kernel's filter: initialised with the external bound (or "none")
            CU_TRY(h->sel_bound.reserve((size_t)nq * 4));
            k_init_bounds<<<(nq + 255) / 256, 256, 0, s>>>(h->sel_bound.as<float>(), d_ext_bound, nq);
            CU_TRY(cudaGetLastError());
            launches++;
            SP.bound_out = h->sel_bound.as<float>();
            sel_bound = h->sel_bound.as<float>();
            P.range_groups = kMaxCuts;
        } else {
            TP.T = T;
            TP.q = d_q;
            TP.q_order = P.q_order;
            TP.q_leaf = P.q_leaf;
            TP.nq = nq;
            TP.K = K;
            TP.tq = tq_main;
            TP.state = P.state;
            TP.ext_bound = d_ext_bound;
            TP.ranges = h->ranges.as<int2>();
            TP.n_ranges = h->n_ranges.as<int>();
            TP.leaf_lo = partial ? std::max(leaf_lo, h->leaf_lo) : h->leaf_lo;
            TP.leaf_hi = partial ? std::min(leaf_hi, h->leaf_hi) : h->leaf_hi;
            TP.counters = ctr;
            TP.ball_only = 0;
        }
        auto bounds_pass = [&](int mode, int step) -> int {
            launches++;
            if (use_tc) {
                SP.sel_mode = mode;
                SP.sel_step = step;
                return launch_select(h, SP, s);
            }
            TP.sel_mode = mode;
            TP.sel_step = step;
            return launch_traverse(h, TP, s);
        };
        auto list_all = [&](int mode, int step) -> int {
            const int cnt = selection_size(n_tiles_main, mode, step);
            if (cnt > 0) k_list_all<<<(cnt + 127) / 128, 128, 0, s>>>(h->ranges.as<int2>(), h->n_ranges.as<int>(), n_tiles_main, use_tc ? kMaxCuts : T.n_cuts, mode, step,
                                                                      page_lo, page_hi);
            CU_TRY(cudaGetLastError());
            launches++;
            return SCHT_OK;
        };
        const bool sampling = n_tiles_main >= 64 && h->opt.traverse_sampling;
        const int kSample = sampling ? std::max(8, std::min(64, n_tiles_main / 24)) : 1;
        const long long sig = (long long)use_tc | ((long long)K << 1) | ((long long)P.pos_lo << 12) | ((long long)(brute ? 1 : 0) << 11) | ((long long)P.pos_hi << 36);
        int rc = SCHT_OK;
        if (use_tc && !can_select) {
            rc = list_all(0, 1);
        } else if (sampling && h->sample_skip > 0 && sig == h->sample_sig) {
            // the samples of the last batches of this shape all kept (nearly) every page: list them without asking again
            h->sample_skip--;
            rc = list_all(0, 1);
        } else {
            // probe: the sampled tiles (or all tiles of a small batch)
            unsigned long long listed = 0;
            rc = bounds_pass(sampling ? 1 : 0, kSample);
            if (rc) return rc;
            CU_TRY(cudaMemcpyAsync(&listed, ctr + 5, 8, cudaMemcpyDeviceToHost, s));
            CU_TRY(cudaStreamSynchronize(s));
            long long probed_q = 0;
            for (int t = 0; t < n_tiles_main; t += kSample) probed_q += std::min(tq_main, nq - t * tq_main);
            const double kept = (double)listed / std::max(1.0, (double)probed_q * n_list_pages);
            // a prefilter page costs ~1/10 of an exact page, so pruning must remove much more before it pays for itself
            const bool worth = kept < (use_tc ? 0.60 : 0.90);
            if (!worth) {
                if (sampling) {
                    rc = list_all(2, kSample);
                    if (sig != h->sample_sig) h->sample_streak = 0;
                    if (++h->sample_streak >= 2) h->sample_skip = h->opt.sample_reuse;
                }
            } else {
                h->sample_streak = 0;
                pruned = kept < 0.5;
                // first pass or not?  auto: when the bound test alone already leaves next to nothing (< 3 % of the pages) a first
                // pass cannot pay for its extra selection and launch; otherwise all queries take it (short lists) or only the loose
                // ones (K > 32, where the convergence of the lists is the expensive part).  Measured on the BASELINE workloads,
                // profiles/r02_first_pass.md.
                int fp = h->opt.first_pass;
                if (fp < 0) fp = kept < 0.03 ? 0 : (K <= 32 ? 2 : 1);
                if (use_tc && fp == 0) {
                    if (sampling) rc = bounds_pass(2, kSample);
                } else if (use_tc) {
                    // Two passes (scht_select.cuh): the queries whose seed bound is worse than the ball bound first scan the two
                    // nearest pages of their nearest super-group (lists -> state); then the bound test selects the pages of the
                    // main scan for all tiles with bounds that are tight for every query.
                    CU_TRY(cudaMemsetAsync(ctr + 5, 0, 8, s));
                    SP.first_pass = fp;
                    rc = bounds_pass(0, 1);
                    if (rc) return rc;
                    rc = launch_tc_scan(kWriteState, false);
                    if (rc) return rc;
                    CU_TRY(cudaMemsetAsync(ctr + 5, 0, 8, s));
                    SP.first_pass = 0;
                    rc = bounds_pass(0, 1);
                } else if (sampling) {
                    rc = bounds_pass(2, kSample);
                }
            }
            h->sample_sig = sig;
        }
        if (rc) return rc;
    }
    if (timed) CU_TRY(cudaEventRecord(h->ev[3], s));
    // K4b main scan
    P.mode = (brute && !use_tc) ? kScanAll : kScanList;
    P.write_mode = keys_wanted ? kWriteKeys : kWriteFinal;
    if (use_tc) {
        int rc = launch_tc_scan(P.write_mode, true);
        if (rc) return rc;
        h->tc_batches++;
        used_tc = true;
    } else {
        int rc = launch_scan(h, variant, P, s);
        if (rc) return rc;
        launches++;
        scans++;
    }
    if (timed) CU_TRY(cudaEventRecord(h->ev[4], s));
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
            float a = 0, b = 0, c2 = 0, tr = 0;
            CU_TRY(cudaEventElapsedTime(&a, h->ev[0], h->ev[4]));
            CU_TRY(cudaEventElapsedTime(&b, h->ev[1], h->ev[2]));
            CU_TRY(cudaEventElapsedTime(&c2, h->ev[3], h->ev[4]));
            CU_TRY(cudaEventElapsedTime(&tr, h->ev[2], h->ev[3]));
            st->ms_device += a;
            st->ms_scan += b + c2;
            st->ms_seed_scan += b;
            st->ms_main_scan += c2;
            st->ms_traverse += tr;
        }
    }
    return SCHT_OK;
}

int check_query_args(const scht_handle* h, const void* q, int64_t nq, int32_t dim, int32_t K, const void* idx, const void* d2) {
    if (!h) return scht::fail(SCHT_ERR_INVALID_ARG, "handle is NULL");
    if (nq < 0 || nq > 0x7fffffffLL) return scht::fail(SCHT_ERR_INVALID_ARG, "n_queries out of range");
    if (K < 1 || K > SCHT_MAX_K) return scht::fail(SCHT_ERR_INVALID_ARG, "K must be in [1, SCHT_MAX_K]");
    const int tdim = h->subs.empty() ? h->T.dim : h->subs[0]->T.dim;
    if (dim != tdim) {
        // Code/Convex_Tree.h:458-461 prints "dimension of query points is not the same as the data" and returns
        return scht::fail(SCHT_ERR_DIM_MISMATCH, "query dimension " + std::to_string(dim) + " != tree dimension " + std::to_string(tdim));
    }
    if (nq > 0 && (!q || !idx || !d2)) return scht::fail(SCHT_ERR_INVALID_ARG, "NULL query or output buffer");
    return SCHT_OK;
}

}  // namespace scht

namespace {

// one whole batch on the device (both phases; the brute-force entry turns its seed lists into the bound)
int run_batch(scht_handle* h, const float* d_q, int nq, int K, int kind, int leaf_lo, int leaf_hi, const QueryOut& out, cudaStream_t s, scht_stats* st,
              bool timed) {
    // Lazy seed scan (ScanParams::lazy_seed) when the main scan will not rely on "the seed pages are done": the brute-force entry
    // (lists start empty, the seed only provides the bound) and tree batches whose main scan is the prefilter scan, which re-scans
    // every listed page anyway -- guessed from the handle's last eligibility verdict; a wrong guess costs the exact scan a second
    // look at the seed pages, never a result bit.
    const bool lazy = h->opt.lazy_seed && h->I.kp > 0 && h->opt.tc_mode != 0 &&
                      (kind == kBatchBrute ? h->opt.brute_tc != 0 : (kind == kBatchTree && (h->opt.tc_mode == 1 || h->elig_last > 0)));
    int rc = batch_seed(h, d_q, nq, K, kind, leaf_lo, leaf_hi, s, st, timed, lazy);
    if (rc) return rc;
    const float* ext = nullptr;
    bool need_bound = kind == kBatchBrute;
    if (kind == kBatchPartial) {
        int plo, phi;
        leaf_range_positions(h, leaf_lo, leaf_hi, plo, phi);
        need_bound = plo > h->T.pos_lo || phi < h->T.pos_hi;  // the seed lists cover more than the leaf range: keep their K-th distance as the bound
    }
    if (need_bound) {
        CU_TRY(h->ext_bound.reserve((size_t)nq * 4));
        rc = batch_bounds(h, nq, K, h->ext_bound.as<float>(), s);
        if (rc) return rc;
        ext = h->ext_bound.as<float>();
    }
    return batch_main(h, d_q, nq, K, kind, leaf_lo, leaf_hi, ext, out, s, st, timed);
}

bool is_pinned_or_device(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type != cudaMemoryTypeUnregistered;
}

}  // namespace

namespace scht {

// scht_query / scht_brute_force on one GPU with HOST buffers.  The call is cut into batches (scht_set_batch_size); with more
// than one batch the copies ride on two extra streams: H2D of batch b+1 and D2H of batch b-1 run under the kernels of batch b
// (the reference's loop is strictly serial: upload, launch, wait, Code/GPU_Convex_Tree.h:208-323).  Pageable caller memory is
// staged through pinned buffers in that case, so that the copies are asynchronous at all.
int host_query(scht_handle* h, const float* queries, int64_t n_queries, int32_t dim, int32_t K, int32_t* idx_out, float* dist2_out,
               float* dist_out, scht_stats* stats, bool brute) {
    int rc = check_query_args(h, queries, n_queries, dim, K, idx_out, dist2_out);
    if (rc) return rc;
    if (stats) memset(stats, 0, sizeof(*stats));
    if (!h->subs.empty()) return multi_query(h, queries, n_queries, dim, K, idx_out, dist2_out, dist_out, stats, brute);
    double t0 = now_ms();
    CU_TRY(cudaSetDevice(h->device));
    int64_t bs = std::min<int64_t>(h->batch_size, 0x7fffffffLL / std::max(K, dim));
    if (h->opt.pipeline && h->opt.pipeline_split > 1 && n_queries <= bs && n_queries >= 8192)
        bs = (n_queries + h->opt.pipeline_split - 1) / h->opt.pipeline_split;
    const int64_t n_batches = (n_queries + bs - 1) / bs;
    const int kind = brute ? kBatchBrute : kBatchTree;
    const bool piped = h->opt.pipeline && n_batches > 1;
    const bool stage_in = piped && !is_pinned_or_device(queries);
    const bool stage_out = piped && !(is_pinned_or_device(idx_out) && is_pinned_or_device(dist2_out) && (!dist_out || is_pinned_or_device(dist_out)));
    const int n_out = dist_out ? 3 : 2;
    auto batch_n = [&](int64_t b) { return (int)std::min<int64_t>(bs, n_queries - b * bs); };
    auto upload = [&](int64_t b, cudaStream_t cs) -> int {
        const int slot = (int)(b & 1), nq = batch_n(b);
        CU_TRY(h->q[slot].reserve((size_t)nq * dim * 4));
        const float* src = queries + (size_t)b * bs * dim;
        if (stage_in) {
            CU_TRY(h->pin_q[slot].reserve((size_t)nq * dim * 4));
            memcpy(h->pin_q[slot].p, src, (size_t)nq * dim * 4);
            src = reinterpret_cast<const float*>(h->pin_q[slot].p);
        }
        CU_TRY(cudaMemcpyAsync(h->q[slot].p, src, (size_t)nq * dim * 4, cudaMemcpyHostToDevice, cs));
        return SCHT_OK;
    };
    // copy a finished batch's staged results into the caller's arrays (host memcpy; only with stage_out)
    auto unstage = [&](int64_t b) -> int {
        const int slot = (int)(b & 1), nq = batch_n(b);
        CU_TRY(cudaEventSynchronize(h->ev_out[slot]));
        const char* p = reinterpret_cast<const char*>(h->pin_out[slot].p);
        const size_t n = (size_t)nq * K * 4, off = (size_t)b * bs * K;
        memcpy(idx_out + off, p, n);
        memcpy(dist2_out + off, p + n, n);
        if (dist_out) memcpy(dist_out + off, p + 2 * n, n);
        return SCHT_OK;
    };
    if (!piped) {
        for (int64_t b = 0; b < n_batches; b++) {
            const int nq = batch_n(b);
            rc = upload(b, h->stream);
            if (rc) return rc;
            const int slot = (int)(b & 1);
            CU_TRY(h->idx[slot].reserve((size_t)nq * K * 4));
            CU_TRY(h->d2[slot].reserve((size_t)nq * K * 4));
            if (dist_out) CU_TRY(h->dist[slot].reserve((size_t)nq * K * 4));
            QueryOut out;
            out.idx = h->idx[slot].as<int32_t>();
            out.d2 = h->d2[slot].as<float>();
            out.dist = dist_out ? h->dist[slot].as<float>() : nullptr;
            rc = run_batch(h, h->q[slot].as<float>(), nq, K, kind, h->leaf_lo, h->leaf_hi, out, h->stream, stats, stats != nullptr);
            if (rc) return rc;
            const size_t off = (size_t)b * bs * K;
            CU_TRY(cudaMemcpyAsync(idx_out + off, out.idx, (size_t)nq * K * 4, cudaMemcpyDeviceToHost, h->stream));
            CU_TRY(cudaMemcpyAsync(dist2_out + off, out.d2, (size_t)nq * K * 4, cudaMemcpyDeviceToHost, h->stream));
            if (dist_out) CU_TRY(cudaMemcpyAsync(dist_out + off, out.dist, (size_t)nq * K * 4, cudaMemcpyDeviceToHost, h->stream));
            CU_TRY(cudaStreamSynchronize(h->stream));
        }
    } else {
        rc = upload(0, h->copy_stream);
        if (rc) return rc;
        CU_TRY(cudaEventRecord(h->ev_in[0], h->copy_stream));
        for (int64_t b = 0; b < n_batches; b++) {
            const int slot = (int)(b & 1), nq = batch_n(b);
            if (b + 1 < n_batches) {  // next batch's queries: the slot's previous kernels (batch b-1) must have finished reading it
                if (b >= 1) CU_TRY(cudaStreamWaitEvent(h->copy_stream, h->ev_done[slot ^ 1], 0));
                if (stage_in && b >= 1) CU_TRY(cudaEventSynchronize(h->ev_in[slot ^ 1]));  // the pinned staging of batch b-1 has been read
                rc = upload(b + 1, h->copy_stream);
                if (rc) return rc;
                CU_TRY(cudaEventRecord(h->ev_in[slot ^ 1], h->copy_stream));
            }
            CU_TRY(h->idx[slot].reserve((size_t)nq * K * 4));
            CU_TRY(h->d2[slot].reserve((size_t)nq * K * 4));
            if (dist_out) CU_TRY(h->dist[slot].reserve((size_t)nq * K * 4));
            QueryOut out;
            out.idx = h->idx[slot].as<int32_t>();
            out.d2 = h->d2[slot].as<float>();
            out.dist = dist_out ? h->dist[slot].as<float>() : nullptr;
            CU_TRY(cudaStreamWaitEvent(h->stream, h->ev_in[slot], 0));
            if (b >= 2) CU_TRY(cudaStreamWaitEvent(h->stream, h->ev_out[slot], 0));  // the slot's previous results (batch b-2) have left the device
            rc = run_batch(h, h->q[slot].as<float>(), nq, K, kind, h->leaf_lo, h->leaf_hi, out, h->stream, stats, stats != nullptr);
            if (rc) return rc;
            CU_TRY(cudaEventRecord(h->ev_done[slot], h->stream));
            CU_TRY(cudaStreamWaitEvent(h->out_stream, h->ev_done[slot], 0));
            const size_t n = (size_t)nq * K * 4, off = (size_t)b * bs * K;
            if (stage_out) {
                if (b >= 2) {
                    rc = unstage(b - 2);  // frees the slot's pinned staging
                    if (rc) return rc;
                }
                CU_TRY(h->pin_out[slot].reserve(n * n_out));
                char* p = reinterpret_cast<char*>(h->pin_out[slot].p);
                CU_TRY(cudaMemcpyAsync(p, out.idx, n, cudaMemcpyDeviceToHost, h->out_stream));
                CU_TRY(cudaMemcpyAsync(p + n, out.d2, n, cudaMemcpyDeviceToHost, h->out_stream));
                if (dist_out) CU_TRY(cudaMemcpyAsync(p + 2 * n, out.dist, n, cudaMemcpyDeviceToHost, h->out_stream));
            } else {
                CU_TRY(cudaMemcpyAsync(idx_out + off, out.idx, n, cudaMemcpyDeviceToHost, h->out_stream));
                CU_TRY(cudaMemcpyAsync(dist2_out + off, out.d2, n, cudaMemcpyDeviceToHost, h->out_stream));
                if (dist_out) CU_TRY(cudaMemcpyAsync(dist_out + off, out.dist, n, cudaMemcpyDeviceToHost, h->out_stream));
            }
            CU_TRY(cudaEventRecord(h->ev_out[slot], h->out_stream));
        }
        if (stage_out) {
            for (int64_t b = std::max<int64_t>(0, n_batches - 2); b < n_batches; b++) {
                rc = unstage(b);
                if (rc) return rc;
            }
        }
        CU_TRY(cudaStreamSynchronize(h->out_stream));
        CU_TRY(cudaStreamSynchronize(h->stream));
        CU_TRY(cudaStreamSynchronize(h->copy_stream));
    }
    if (stats) stats->ms_total = now_ms() - t0;
    return SCHT_OK;
}

}  // namespace scht

extern "C" int scht_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

extern "C" int scht_create(const scht_tree_desc* t, int device, scht_handle** out) {
    if (device == -1) return scht_create_multi(t, nullptr, 0, SCHT_SHARD_QUERIES, out);  // every visible device, queries sharded
    return create_handle(t, device, 0, t ? t->n_leaves : 0, nullptr, nullptr, out);
}

extern "C" int scht_create_shard(const scht_tree_desc* t, int device, int32_t leaf_begin, int32_t leaf_end, scht_handle** out) {
    return create_handle(t, device, leaf_begin, leaf_end, nullptr, nullptr, out);
}

extern "C" void scht_destroy(scht_handle* h) {
    if (!h) return;
    if (!h->subs.empty()) {
        multi_destroy(h);
        delete h;
        return;
    }
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    if (h->copy_stream) cudaStreamSynchronize(h->copy_stream);
    if (h->out_stream) cudaStreamSynchronize(h->out_stream);
    for (void* p : h->owned) cudaFree(p);
    for (DevBuf* b : {&h->q[0], &h->q[1], &h->idx[0], &h->idx[1], &h->d2[0], &h->d2[1], &h->dist[0], &h->dist[1], &h->q_leaf, &h->hist, &h->q_order,
                      &h->state, &h->ranges, &h->n_ranges, &h->counters, &h->seg_keys, &h->ext_bound, &h->bound_own, &h->sel_bound, &h->tmp_keys})
        b->release();
    for (PinBuf* b : {&h->pin_q[0], &h->pin_q[1], &h->pin_out[0], &h->pin_out[1]}) b->release();
    for (auto& e : h->ev)
        if (e) cudaEventDestroy(e);
    for (int i = 0; i < 2; i++) {
        if (h->ev_in[i]) cudaEventDestroy(h->ev_in[i]);
        if (h->ev_done[i]) cudaEventDestroy(h->ev_done[i]);
        if (h->ev_out[i]) cudaEventDestroy(h->ev_out[i]);
    }
    if (h->stream) cudaStreamDestroy(h->stream);
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    if (h->out_stream) cudaStreamDestroy(h->out_stream);
    delete h;
}

extern "C" int scht_set_batch_size(scht_handle* h, int64_t batch_size) {
    if (!h || batch_size < 1) return scht::fail(SCHT_ERR_INVALID_ARG, "scht_set_batch_size: batch_size must be >= 1");
    h->batch_size = batch_size;
    for (scht_handle* s : h->subs) s->batch_size = batch_size;
    return SCHT_OK;
}

namespace {
struct OptEntry {
    const char* name;
    int Options::*field;
    int lo, hi;
};
const OptEntry kOptions[] = {
    {"scan_variant", &Options::scan_variant, -1, 3},
    {"scan_stages", &Options::scan_stages, 0, 4},
    {"traverse_sampling", &Options::traverse_sampling, 0, 1},
    {"sample_reuse", &Options::sample_reuse, 0, 1 << 20},
    {"tc", &Options::tc_mode, 0, 2},
    {"tc_split", &Options::tc_split, 0, 8},
    {"tc_issuers", &Options::tc_issuers, 0, 4},
    {"tc_n256", &Options::tc_n256, -1, 1},
    {"seed_points", &Options::seed_points, 1, 1 << 30},
    {"page_select", &Options::page_select, 0, 1},
    {"brute_tc", &Options::brute_tc, 0, 1},
    {"pipeline", &Options::pipeline, 0, 1},
    {"pipeline_split", &Options::pipeline_split, 0, 64},
    {"rescorer_sleep_ns", &Options::rescorer_sleep_ns, 0, 1000000},
    {"first_pass", &Options::first_pass, -1, 2},
    {"tc_ring", &Options::tc_ring, 0, 1024},
    {"tc_recheck", &Options::tc_recheck, 0, 1},
    {"lazy_seed", &Options::lazy_seed, 0, 1},
    {"ring_wait_ns", &Options::ring_wait_ns, 0, 1000000},
    {"groups", &Options::groups, -1, 1 << 20},
};
}  // namespace

extern "C" int scht_set_option(scht_handle* h, const char* name, int64_t value) {
    if (!h || !name) return scht::fail(SCHT_ERR_INVALID_ARG, "scht_set_option: NULL argument");
    for (const OptEntry& e : kOptions) {
        if (strcmp(name, e.name) != 0) continue;
        if (value < e.lo || value > e.hi) return scht::fail(SCHT_ERR_INVALID_ARG, std::string("scht_set_option: value out of range for ") + name);
        if (e.field == &Options::groups) return scht::fail(SCHT_ERR_INVALID_ARG, "scht_set_option: 'groups' is fixed when the handle is created (SCHT_GROUPS)");
        if (e.field == &Options::tc_ring && value != 0 && (value < 32 || (value & (value - 1)) != 0))
            return scht::fail(SCHT_ERR_INVALID_ARG, "scht_set_option: tc_ring is 0 (default) or a power of two in 32..1024");
        if (e.field == &Options::tc_mode && value != 0 && h->subs.empty() && h->I.kp == 0 && h->opt.tc_mode == 0)
            return scht::fail(SCHT_ERR_INVALID_ARG, "scht_set_option: the handle was created without the tensor-core image (SCHT_TC=0)");
        std::vector<scht_handle*> all = h->subs.empty() ? std::vector<scht_handle*>{h} : h->subs;
        h->opt.*(e.field) = (int)value;
        for (scht_handle* x : all) {
            x->opt.*(e.field) = (int)value;
            x->sample_skip = 0;
            x->sample_streak = 0;
            x->elig_skip = 0;
            x->elig_streak = 0;
            if (e.field == &Options::seed_points) {
                int rc = rebuild_seed_ranges(x);
                if (rc) return rc;
            }
        }
        return SCHT_OK;
    }
    return scht::fail(SCHT_ERR_INVALID_ARG, std::string("scht_set_option: unknown option ") + name);
}

extern "C" int scht_get_option(const scht_handle* h, const char* name, int64_t* value) {
    if (!h || !name || !value) return scht::fail(SCHT_ERR_INVALID_ARG, "scht_get_option: NULL argument");
    const scht_handle* x = h->subs.empty() ? h : h->subs[0];
    for (const OptEntry& e : kOptions)
        if (strcmp(name, e.name) == 0) {
            *value = x->opt.*(e.field);
            return SCHT_OK;
        }
    return scht::fail(SCHT_ERR_INVALID_ARG, std::string("scht_get_option: unknown option ") + name);
}

extern "C" int scht_query(scht_handle* h, const float* queries, int64_t n_queries, int32_t dim, int32_t K, int32_t* idx_out,
                          float* dist2_out, float* dist_out, scht_stats* stats) {
    return host_query(h, queries, n_queries, dim, K, idx_out, dist2_out, dist_out, stats, false);
}

extern "C" int scht_brute_force(scht_handle* h, const float* queries, int64_t n_queries, int32_t dim, int32_t K, int32_t* idx_out,
                                float* dist2_out, float* dist_out, scht_stats* stats) {
    return host_query(h, queries, n_queries, dim, K, idx_out, dist2_out, dist_out, stats, true);
}

namespace {
int plain_only(const scht_handle* h, const char* what) {
    if (h && !h->subs.empty()) return scht::fail(SCHT_ERR_INVALID_ARG, std::string(what) + ": device-pointer entry points need a single-GPU handle");
    return SCHT_OK;
}
}  // namespace

extern "C" int scht_query_device(scht_handle* h, const float* d_queries, int64_t n_queries, int32_t dim, int32_t K, int32_t* d_idx_out,
                                 float* d_dist2_out, float* d_dist_out, void* cuda_stream, scht_stats* stats) {
    int rc = check_query_args(h, d_queries, n_queries, dim, K, d_idx_out, d_dist2_out);
    if (rc) return rc;
    if ((rc = plain_only(h, "scht_query_device"))) return rc;
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
        rc = run_batch(h, d_queries + (size_t)b0 * dim, nq, K, kBatchTree, h->leaf_lo, h->leaf_hi, out, s, stats, stats != nullptr);
        if (rc) return rc;
    }
    CU_TRY(cudaStreamSynchronize(s));
    if (stats) stats->ms_total = now_ms() - t0;
    return SCHT_OK;
}

extern "C" int scht_brute_force_device(scht_handle* h, const float* d_queries, int64_t n_queries, int32_t dim, int32_t K, int32_t* d_idx_out,
                                       float* d_dist2_out, float* d_dist_out, void* cuda_stream, scht_stats* stats) {
    int rc = check_query_args(h, d_queries, n_queries, dim, K, d_idx_out, d_dist2_out);
    if (rc) return rc;
    if ((rc = plain_only(h, "scht_brute_force_device"))) return rc;
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
        rc = run_batch(h, d_queries + (size_t)b0 * dim, nq, K, kBatchBrute, h->leaf_lo, h->leaf_hi, out, s, stats, stats != nullptr);
        if (rc) return rc;
    }
    CU_TRY(cudaStreamSynchronize(s));
    if (stats) stats->ms_total = now_ms() - t0;
    return SCHT_OK;
}

extern "C" int scht_find_approximate_leaves(scht_handle* h, const float* queries, int64_t n_queries, int32_t dim, int32_t* leaf_out) {
    if (!h) return scht::fail(SCHT_ERR_INVALID_ARG, "handle is NULL");
    if (!h->subs.empty()) h = h->subs[0];  // the tree top is complete on every device
    if (dim != h->T.dim) return scht::fail(SCHT_ERR_DIM_MISMATCH, "query dimension != tree dimension");
    if (n_queries < 0 || n_queries > 0x7fffffffLL / std::max(1, dim) || (n_queries > 0 && (!queries || !leaf_out)))
        return scht::fail(SCHT_ERR_INVALID_ARG, "scht_find_approximate_leaves: bad arguments");
    if (n_queries == 0) return SCHT_OK;
    CU_TRY(cudaSetDevice(h->device));
    int nq = (int)n_queries;
    CU_TRY(h->q[0].reserve((size_t)nq * dim * 4));
    CU_TRY(h->q_leaf.reserve((size_t)nq * 4));
    CU_TRY(cudaMemcpyAsync(h->q[0].p, queries, (size_t)nq * dim * 4, cudaMemcpyHostToDevice, h->stream));
    k_descend<<<(nq + 127) / 128, 128, 0, h->stream>>>(h->T, h->q[0].as<float>(), nq, h->q_leaf.as<int>(), nullptr, nullptr);
    CU_TRY(cudaGetLastError());
    CU_TRY(cudaMemcpyAsync(leaf_out, h->q_leaf.p, (size_t)nq * 4, cudaMemcpyDeviceToHost, h->stream));
    CU_TRY(cudaStreamSynchronize(h->stream));
    return SCHT_OK;
}

namespace {
int check_partial(scht_handle* h, const void* d_q, int64_t nq, int32_t dim, int32_t K, int32_t& leaf_begin, int32_t& leaf_end, const void* out) {
    int rc = check_query_args(h, d_q, nq, dim, K, out, out);
    if (rc) return rc;
    if ((rc = plain_only(h, "sharded-dataset entry points"))) return rc;
    if (leaf_begin < 0 && leaf_end < 0) {  // the handle's resident leaves
        leaf_begin = h->leaf_lo;
        leaf_end = h->leaf_hi;
    }
    if (leaf_begin < h->leaf_lo || leaf_end > h->leaf_hi || leaf_begin > leaf_end)
        return scht::fail(SCHT_ERR_INVALID_ARG, "leaf range outside the leaves resident on this handle");
    if (nq > 0x7fffffffLL / std::max(K, dim)) return scht::fail(SCHT_ERR_INVALID_ARG, "n_queries * max(K, dim) must stay below 2^31");
    return SCHT_OK;
}
}  // namespace

extern "C" int scht_query_partial_device(scht_handle* h, const float* d_queries, int64_t n_queries, int32_t dim, int32_t K,
                                         int32_t leaf_begin, int32_t leaf_end, uint64_t* d_keys_out, void* cuda_stream, scht_stats* stats) {
    int rc = check_partial(h, d_queries, n_queries, dim, K, leaf_begin, leaf_end, d_keys_out);
    if (rc) return rc;
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
        rc = run_batch(h, d_queries + (size_t)b0 * dim, nq, K, kBatchPartial, leaf_begin, leaf_end, out, s, stats, stats != nullptr);
        if (rc) return rc;
    }
    CU_TRY(cudaStreamSynchronize(s));
    if (stats) stats->ms_total = now_ms() - t0;
    return SCHT_OK;
}

extern "C" int scht_partial_seed_device(scht_handle* h, const float* d_queries, int64_t n_queries, int32_t dim, int32_t K, float* d_bound_out,
                                        void* cuda_stream) {
    int32_t lb = -1, le = -1;
    int rc = check_partial(h, d_queries, n_queries, dim, K, lb, le, d_bound_out);
    if (rc) return rc;
    if (n_queries == 0) return SCHT_OK;
    CU_TRY(cudaSetDevice(h->device));
    cudaStream_t s = cuda_stream ? (cudaStream_t)cuda_stream : h->stream;
    rc = batch_seed(h, d_queries, (int)n_queries, K, kBatchPartial, lb, le, s, nullptr, false);
    if (rc) return rc;
    rc = batch_bounds(h, (int)n_queries, K, d_bound_out, s);
    if (rc) return rc;
    CU_TRY(cudaStreamSynchronize(s));
    return SCHT_OK;
}

extern "C" int scht_partial_scan_device(scht_handle* h, const float* d_queries, int64_t n_queries, int32_t dim, int32_t K, const float* d_bound,
                                        uint64_t* d_keys_out, void* cuda_stream, scht_stats* stats) {
    int32_t lb = -1, le = -1;
    int rc = check_partial(h, d_queries, n_queries, dim, K, lb, le, d_keys_out);
    if (rc) return rc;
    if (stats) memset(stats, 0, sizeof(*stats));
    if (n_queries == 0) return SCHT_OK;
    if (h->bs.nq != (int)n_queries || h->bs.K != K || h->bs.kind != kBatchPartial)
        return scht::fail(SCHT_ERR_INVALID_ARG, "scht_partial_scan_device must follow scht_partial_seed_device of the same batch");
    CU_TRY(cudaSetDevice(h->device));
    cudaStream_t s = cuda_stream ? (cudaStream_t)cuda_stream : h->stream;
    QueryOut out;
    out.keys = d_keys_out;
    rc = batch_main(h, d_queries, (int)n_queries, K, kBatchPartial, lb, le, d_bound, out, s, stats, false);
    if (rc) return rc;
    h->bs.nq = 0;
    CU_TRY(cudaStreamSynchronize(s));
    return SCHT_OK;
}

extern "C" int scht_merge_partial_device(scht_handle* h, const uint64_t* d_keys, int32_t n_lists, const float* d_queries, int64_t n_queries,
                                         int32_t dim, int32_t K, int32_t* d_idx_out, float* d_dist2_out, float* d_dist_out, void* cuda_stream) {
    (void)d_queries;
    int rc = check_query_args(h, d_keys, n_queries, dim, K, d_idx_out, d_dist2_out);
    if (rc) return rc;
    if ((rc = plain_only(h, "scht_merge_partial_device"))) return rc;
    if (n_lists < 1 || n_lists > kMaxShards) return scht::fail(SCHT_ERR_INVALID_ARG, "n_lists must be in [1,16]");
    if (n_queries == 0) return SCHT_OK;
    CU_TRY(cudaSetDevice(h->device));
    cudaStream_t s = cuda_stream ? (cudaStream_t)cuda_stream : h->stream;
    rc = merge_lists(h, d_keys, n_lists, (int)n_queries, K, false, d_idx_out, d_dist2_out, d_dist_out, s);
    if (rc) return rc;
    CU_TRY(cudaStreamSynchronize(s));
    return SCHT_OK;
}

extern "C" int scht_handle_info(const scht_handle* h, int32_t* dim, int64_t* n_points, int32_t* n_nodes, int32_t* n_leaves, int32_t* device,
                                int64_t* device_bytes) {
    if (!h) return scht::fail(SCHT_ERR_INVALID_ARG, "handle is NULL");
    const scht_handle* x = h->subs.empty() ? h : h->subs[0];
    if (dim) *dim = x->T.dim;
    if (n_points) *n_points = x->T.n_points;
    if (n_nodes) *n_nodes = x->T.n_nodes;
    if (n_leaves) *n_leaves = x->T.n_leaves;
    if (device) *device = x->device;
    if (device_bytes) {
        int64_t b = 0;
        if (h->subs.empty()) b = (int64_t)h->tree_bytes;
        else
            for (const scht_handle* s : h->subs) b += (int64_t)s->tree_bytes;
        *device_bytes = b;
    }
    return SCHT_OK;
}

extern "C" int scht_handle_layout(const scht_handle* h, int32_t* n_devices, int32_t* mode, int32_t* n_scan_groups, int32_t* n_super_groups,
                                  int64_t* n_scan_pages, int32_t* leaf_begin, int32_t* leaf_end) {
    if (!h) return scht::fail(SCHT_ERR_INVALID_ARG, "handle is NULL");
    const scht_handle* x = h->subs.empty() ? h : h->subs[0];
    if (n_devices) *n_devices = h->subs.empty() ? 1 : (int)h->subs.size();
    if (mode) *mode = h->multi_mode;
    if (n_scan_groups) *n_scan_groups = x->I.kp ? x->I.n_groups : 0;
    if (n_super_groups) *n_super_groups = x->I.kp ? x->I.n_supers : 0;
    if (n_scan_pages) *n_scan_pages = x->I.kp ? x->I.n_spages : 0;
    if (leaf_begin) *leaf_begin = x->leaf_lo;
    if (leaf_end) *leaf_end = x->leaf_hi;
    return SCHT_OK;
}
