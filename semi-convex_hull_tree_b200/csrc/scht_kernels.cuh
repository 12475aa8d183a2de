// sm_100a device code of the Semi-convex Hull Tree exact k-NN query path.
//
// Replaces the reference's OpenCL kernels (Code/kernels/do_kNN_gpu.cl:298-446 do_finding_KNN_by_leaf_order,
// :462-536 do_brute_force_KNN) and the host descent loop (Code/GPU_Convex_Tree.h:168-193).  Not a port: the
// reference is thread-per-query with private arrays; here queries are bucketed by approximate leaf into tiles of
// 8 queries per warp, a producer lane streams "pages" (256 sorted points, dim-major) into shared memory with
// cp.async.bulk + mbarriers, and every staged page is reused by all queries of the tile.
//
// RESULT CONTRACT: for every query the K smallest keys
//     key = float_bits(d2) << 32 | r << 31 | pos
// d2  = the reference's distance arithmetic acc = (float)((double)acc + (double)d*(double)d), d = fl32(p_j - q_j),
//       j ascending (Code/basic_functions.h:167-198),
// r   = 0 for points of the query's approximate leaf, else 1 (the reference scans that leaf first,
//       Code/CPU_Convex_Tree.h:181-217, and NNResult::update keeps the first arrival among equals,
//       Code/Convex_Tree.h:121-151),
// pos = row in the leaf-sorted point matrix (leaves are visited in ascending order, Code/CPU_Convex_Tree.h:386-427).
// The hot loop evaluates d2 with packed FP32 FMAs (FADD2/FFMA2, one rounding per step); any pair that comes within
// 2^-20 (relative) of the query's current K-th distance is re-evaluated with the reference's float/double
// arithmetic before it is compared, so accepted values are the reference's bit for bit.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <type_traits>

#include "scht_types.cuh"

namespace scht {

// ---------------------------------------------------------------------------------------------------------
// small PTX helpers
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("{ .reg .b64 st; mbarrier.arrive.shared::cta.b64 st, [%0]; }" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("{ .reg .b64 st; mbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1; }" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
// try_wait suspends the thread in hardware until the phase completes or the hint (ns) expires: a long hint keeps waiting
// warps from re-issuing the poll (they share issue slots with the warps doing the work)
constexpr uint32_t kMbarSuspendNs = 20000;
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity), "r"(kMbarSuspendNs)
        : "memory");
}
#ifndef SCHT_PRODUCER_SLEEP_NS
#define SCHT_PRODUCER_SLEEP_NS 200
#endif
// Producer-side wait: the producer lane is normally several stages ahead, so it polls with a sleep in between instead
// of spinning (a spinning lane steals issue slots from the consumer warps that share its scheduler).
__device__ __forceinline__ void mbar_wait_backoff(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    for (;;) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (done) break;
        __nanosleep(SCHT_PRODUCER_SLEEP_NS);
    }
}
// 1-D bulk async copy global -> shared (TMA engine, SASS UBLKCP), completion counted on an mbarrier.
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// packed FP32 pairs (Blackwell FADD2 / FFMA2): two IEEE fp32 lanes per instruction, each rounded like the scalar op
__device__ __forceinline__ uint64_t pack2(float lo, float hi) {
    uint64_t d;
    asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi));
    return d;
}
__device__ __forceinline__ void unpack2(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t sub2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("sub.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}

__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("mul.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}

// The reference's accumulation step: result(float) += pow(d, 2) evaluated in double (Code/basic_functions.h:177,194).
__device__ __forceinline__ float ref_acc_sq(float acc, float d) { return (float)((double)acc + (double)d * (double)d); }
// A wide point's row spans several 128-byte lines: ask for all of them before the dependent chain walks the row, so the
// chain pays one memory round trip instead of one per batch of loads
__device__ __forceinline__ void prefetch_row(const float4* row, int dpad) {
    for (int c = 8; c < dpad / 4; c += 8) asm volatile("prefetch.global.L1 [%0];" ::"l"(row + c));
}

// ---------------------------------------------------------------------------------------------------------
// K1: approximate-leaf descent (replaces the host loop Code/GPU_Convex_Tree.h:168-193 == Code/CPU_Convex_Tree.h:181-197)
// thread per query; the two centre distances use the reference arithmetic so the branch decisions are identical.
// Also histograms the leaves (first step of the bucketing).
// ---------------------------------------------------------------------------------------------------------
__global__ void k_descend(DevTree T, const float* __restrict__ q, int nq, int* __restrict__ q_leaf, int* __restrict__ leaf_hist,
                          unsigned long long* __restrict__ counters) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nq) return;
    const float* qi = q + (size_t)i * T.dim;
    const bool vec4 = (T.dim & 3) == 0 && (reinterpret_cast<uintptr_t>(q) & 15) == 0;  // query rows 16-byte aligned (the centre rows always are)
    int node = 0, levels = 0;
    NodeRec nr = T.nodes[0];
    while (nr.leaf < 0) {
        const float* lc = T.node_lc + (size_t)node * T.dpad;
        const float* rc = T.node_rc + (size_t)node * T.dpad;
        // Only the COMPARISON has to be the reference's.  A plain FP32 FMA evaluation (any summation order) is within
        // (dim + 2) 2^-24 (relative) of the reference's value (same differences, one rounding per step instead of the
        // float(double) step), so when the two distances differ by more than 2^-14 of their sum the outcome is certain;
        // near-ties take the exact arithmetic.
        float l = 0.f, r = 0.f;
        if (vec4) {  // 16-byte loads, two independent chains per side: the level's latency is what this kernel pays
            float l1 = 0.f, r1 = 0.f;
            const float4* q4 = reinterpret_cast<const float4*>(qi);
            const float4* lc4 = reinterpret_cast<const float4*>(lc);
            const float4* rc4 = reinterpret_cast<const float4*>(rc);
#pragma unroll 4
            for (int j = 0; j < (T.dim >> 2); j++) {
                const float4 x = __ldg(q4 + j), a = __ldg(lc4 + j), b = __ldg(rc4 + j);
                const float a0 = x.x - a.x, a1 = x.y - a.y, a2 = x.z - a.z, a3 = x.w - a.w;
                const float b0 = x.x - b.x, b1 = x.y - b.y, b2 = x.z - b.z, b3 = x.w - b.w;
                l = fmaf(a0, a0, l), l1 = fmaf(a1, a1, l1), l = fmaf(a2, a2, l), l1 = fmaf(a3, a3, l1);
                r = fmaf(b0, b0, r), r1 = fmaf(b1, b1, r1), r = fmaf(b2, b2, r), r1 = fmaf(b3, b3, r1);
            }
            l += l1;
            r += r1;
        } else {
            for (int j = 0; j < T.dim; j++) {
                const float x = qi[j], dl = x - lc[j], dr = x - rc[j];
                l = fmaf(dl, dl, l);
                r = fmaf(dr, dr, r);
            }
        }
        if (!(fabsf(l - r) > 6.1035156e-5f * (l + r)) || T.dim > 1024) {
            l = 0.f, r = 0.f;
            for (int j = 0; j < T.dim; j++) {
                float x = qi[j];
                l = ref_acc_sq(l, x - lc[j]);
                r = ref_acc_sq(r, x - rc[j]);
            }
        }
        node = (l >= r) ? nr.right : node + 1;  // ties go right (Code/CPU_Convex_Tree.h:191)
        nr = T.nodes[node];
        levels++;
    }
    q_leaf[i] = nr.leaf;
    if (leaf_hist) atomicAdd(&leaf_hist[T.leaf_rank[nr.leaf]], 1);  // bucket = rank of the leaf (leaves ordered by scan group)
    if (counters) atomicAdd(&counters[3], (unsigned long long)(2 * levels));
}

// K3a: exclusive scan of the leaf histogram (single CTA; n_leaves is at most a few hundred thousand)
__global__ void k_leaf_prefix_sum(int* __restrict__ hist, int n) {
    __shared__ int warp_sums[32];
    __shared__ int carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int base = 0; base < n; base += blockDim.x) {
        int i = base + threadIdx.x;
        int v = (i < n) ? hist[i] : 0;
        int x = v;
        for (int o = 1; o < 32; o <<= 1) {
            int y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) warp_sums[w] = x;
        __syncthreads();
        if (w == 0) {
            int s = (lane < (int)(blockDim.x >> 5)) ? warp_sums[lane] : 0;
            for (int o = 1; o < 32; o <<= 1) {
                int y = __shfl_up_sync(0xffffffffu, s, o);
                if (lane >= o) s += y;
            }
            warp_sums[lane] = s;  // inclusive
        }
        __syncthreads();
        int carry = carry_s;
        int excl = carry + (w > 0 ? warp_sums[w - 1] : 0) + x - v;
        if (i < n) hist[i] = excl;
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) carry_s = carry + warp_sums[(blockDim.x >> 5) - 1];
        __syncthreads();
    }
}

// K3b: scatter query ids into their leaf bucket -> q_order (queries sharing an approximate leaf become neighbours,
// and leaves near the same scan group are neighbours too (DevTree::leaf_rank), so one tile's queries want the same pages).  Order inside a bucket is arbitrary; results do not depend on it.
__global__ void k_scatter(const int* __restrict__ q_leaf, const int* __restrict__ leaf_rank, int nq, int* __restrict__ cursor, int* __restrict__ q_order) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nq) return;
    int slot = atomicAdd(&cursor[leaf_rank[q_leaf[i]]], 1);
    q_order[slot] = i;
}
__global__ void k_iota(int* __restrict__ a, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] = i;
}

// ---------------------------------------------------------------------------------------------------------
// K4: leaf scan + top-K
// ---------------------------------------------------------------------------------------------------------
enum ScanMode : int {
    kScanAll = 0,   // every page, keys carry the ORIGINAL index, no approximate-leaf priority (brute force entry)
    kScanSeed = 1,  // pages covering the approximate leaves of the tile's queries; lists start empty
    kScanList = 2,  // page ranges produced by the traversal kernel; lists resume from `state`
};

enum WriteMode : int { kWriteFinal = 1, kWriteState = 0, kWriteKeys = 2 };

struct ScanParams {
    DevTree T;
    const float* q;        // [nq][dim] row-major, caller order
    const int* q_order;    // [nq] sorted slot -> query id
    const int* q_leaf;     // [nq] approximate leaf by query id (unused in kScanAll)
    int nq, K;
    int mode;
    int write_mode;        // kWriteFinal / kWriteState / kWriteKeys
    int pos_lo, pos_hi;    // only positions in [pos_lo,pos_hi) may enter the lists (sharded-dataset mode)
    int lazy_seed;         // the batch's seed scan (kScanSeed) keeps its lists by FAST-path distances and evaluates only the K entries
                           //   that are left with the reference arithmetic at the end (the lists it writes are exact keys of real points,
                           //   but not necessarily the K nearest of the seed pages: a later kScanList pass must not skip those pages)
    const int2* ranges;    // [n_tiles][kMaxCuts][kJobRanges] page ranges [x,y) (kScanList)
    const int* n_ranges;   // [n_tiles][kMaxCuts]
    int range_groups;      // groups of kJobRanges ranges to walk per tile (traversal: T.n_cuts, page selection: kMaxCuts)
    const float* ext_bound;  // [nq] by query id, or null: an upper bound of the K-th distance^2 known from elsewhere (seed scan of
                           //   another shard, seed scan of the tree for the brute-force entry): pairs beyond it are never evaluated
    uint64_t* state;       // [nq][K] keys by sorted slot
    uint64_t* keys_out;    // [nq][K] keys by query id (kWriteKeys)
    // kWriteKeys with n_slices > 0 (sharded-dataset mode inside one process): the K keys of query qid are stored straight into
    // the gather buffer of the device that owns the query slice g (slice_lo[g] <= qid < slice_lo[g+1]) -- peer memory over
    // NVLink -- at key_tab[g] + ((size_t)list * slice_len(g) + qid - slice_lo[g]) * K, so the exchange needs no separate copy
    int n_slices, list;
    uint64_t* key_tab[kMaxShards];
    int slice_lo[kMaxShards + 1];
    int* idx_out;          // [nq][K]
    float* d2_out;         // [nq][K]
    float* dist_out;       // [nq][K] or NULL
    unsigned long long* counters;  // [0] scanned (query,point) pairs, [1] exact re-evaluations, [2] list insertions
};

__device__ __forceinline__ uint64_t* keys_dst(const ScanParams& P, int qid) {
    if (P.n_slices <= 0) return P.keys_out + (size_t)qid * P.K;
    int g = 0;
    while (g + 1 < P.n_slices && qid >= P.slice_lo[g + 1]) g++;
    return P.key_tab[g] + ((size_t)P.list * (size_t)(P.slice_lo[g + 1] - P.slice_lo[g]) + (size_t)(qid - P.slice_lo[g])) * P.K;
}
__device__ __forceinline__ float ext_thr_of(const ScanParams& P, int qid) {
    return (P.ext_bound && qid >= 0) ? fminf(P.ext_bound[qid] * 1.000001f, 3.402823466e+38f) : 3.402823466e+38f;
}

__host__ __device__ inline size_t scan_smem_bytes(int n_cwarps, int dpad, int K, int stages = kStages) {
    size_t tq = (size_t)n_cwarps * kQPerWarp;
    size_t b = (size_t)stages * kRowsPerStage * kPage * 4;  // stages
    b += (size_t)dpad * tq * 4;                              // queries, dim-major
    b += tq * (size_t)K * 8;                                 // lists
    b += tq * 8 + tq * 6 * 4 + 16;                           // seed ranges, a0, a1, s0, s1, qid, external bound, misc
    b += 2 * stages * 8 + 64;                                // barriers
    return b;
}

// Warp-cooperative sorted insert of `key` into L[0..K) (ascending, unique keys); the last element drops out.
// Idempotent: a key already present is ignored, so re-scanning a page can never duplicate a neighbour.
__device__ __forceinline__ int list_insert(uint64_t* L, int K, uint64_t key, int lane) {
    uint64_t nv[4];
    bool dup = false;
#pragma unroll
    for (int s = 0; s < 4; s++) {
        int i = s * 32 + lane;
        nv[s] = 0;
        if (i < K) {
            uint64_t cur = L[i];
            uint64_t prev = (i > 0) ? L[i - 1] : 0ull;
            dup |= (cur == key);
            nv[s] = (cur < key) ? cur : ((i == 0 || prev < key) ? key : prev);
        }
    }
    if (__any_sync(0xffffffffu, dup)) return 0;
    __syncwarp();
#pragma unroll
    for (int s = 0; s < 4; s++) {
        int i = s * 32 + lane;
        if (i < K) L[i] = nv[s];
    }
    __syncwarp();
    return 1;
}

// Ascending, disjoint page ranges covering the seed neighbourhoods of a tile's queries.  The queries of a tile are
// neighbours in BUCKET order (leaves ordered by scan group, DevTree::leaf_rank), so their neighbourhoods arrive in any
// position order: insertion sort by first page, then merge what overlaps or touches (<= 64 entries, one thread).
__device__ inline int build_seed_ranges(const int* a0s, const int* a1s, int nq_tile, int pos_lo, int pos_hi, int2* out) {
    int n = 0;
    for (int i = 0; i < nq_tile; i++) {
        int a0 = max(a0s[i], pos_lo), a1 = min(a1s[i], pos_hi);
        if (a1 <= a0) continue;
        const int2 r = make_int2(a0 / kPage, (a1 + kPage - 1) / kPage);
        if (n > 0 && r.x == out[n - 1].x && r.y == out[n - 1].y) continue;  // same leaf as the previous query (the common case)
        int j = n++;
        while (j > 0 && out[j - 1].x > r.x) {
            out[j] = out[j - 1];
            j--;
        }
        out[j] = r;
    }
    int m = 0;
    for (int i = 0; i < n; i++) {
        if (m > 0 && out[i].x <= out[m - 1].y) out[m - 1].y = max(out[m - 1].y, out[i].y);
        else out[m++] = out[i];
    }
    return m;
}

// Walks the pages of a tile in ascending order: the ranges of `n_groups` consecutive groups (group g holds cnt[g]
// ranges at list + g*stride), minus (optionally) the pages of `seed`; a page is never returned twice.  Producer and
// consumers run the same walk, so they agree on the stage sequence without communicating.
struct PageWalk {
    const int2* list;
    const int* cnt;
    int n_groups, stride;
    const int2* seed;
    int n_seed;  // 0: nothing to skip
    int g, r, cnt_g, page, cur_end, sr;  // pages [page, cur_end) of the current range are still to be returned
    int w_lo, w_hi;                      // only pages inside the window [w_lo, w_hi) are returned (page segments of a split tile)
    __device__ PageWalk(const int2* l, const int* c, int ng, int st, const int2* s, int ns, int win_lo = 0, int win_hi = 0x7fffffff)
        : list(l), cnt(c), n_groups(ng), stride(st), seed(s), n_seed(ns), g(0), r(0), page(win_lo), cur_end(0), sr(0), w_lo(win_lo), w_hi(win_hi) {
        cnt_g = (ng > 0) ? c[0] : 0;
    }
    // Next run of consecutive pages [lo, hi) (no seed ranges to skip): the caller's inner loop is a plain counter, which
    // matters for the single-warp roles of k_scan_tc whose per-page instruction count sits on the critical path.
    __device__ __forceinline__ bool next_range(int& lo, int& hi) {
        for (;;) {
            while (g < n_groups && r >= cnt_g) {
                g++;
                r = 0;
                if (g < n_groups) cnt_g = cnt[g];
            }
            if (g >= n_groups) return false;
            const int2 rg = list[(size_t)g * stride + r];
            r++;
            const int x = max(rg.x, page), y = min(rg.y, w_hi);
            if (x < y) {
                lo = x;
                hi = y;
                page = y;
                return true;
            }
        }
    }
    __device__ __forceinline__ bool next(int& out) {
        for (;;) {
            if (page < cur_end) {  // fast path: registers only (the range list lives in global memory)
                if (n_seed) {
                    while (sr < n_seed && seed[sr].y <= page) sr++;
                    if (sr < n_seed && page >= seed[sr].x) {
                        page = seed[sr].y;
                        continue;
                    }
                }
                out = page++;
                return true;
            }
            while (g < n_groups && r >= cnt_g) {
                g++;
                r = 0;
                if (g < n_groups) cnt_g = cnt[g];
            }
            if (g >= n_groups) return false;
            const int2 rg = list[(size_t)g * stride + r];
            r++;
            if (rg.x > page) page = rg.x;
            cur_end = min(rg.y, w_hi);
        }
    }
};

// FULL: dpad is a multiple of 16, every stage holds 16 rows and the row loop is straight-line code.
// STAGES: depth of the slice ring (4; 2 for wide points, whose query block would otherwise cost the third resident CTA).
// LAZY: the seed scan of a batch with ScanParams::lazy_seed (a template parameter so that the other passes keep their code)
template <int NW, int MINB, bool FULL, int STAGES = kStages, bool LAZY = false>
__global__ void __launch_bounds__((NW + 1) * 32, MINB) k_scan(const ScanParams P) {
    constexpr int TQ = NW * kQPerWarp;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const DevTree& T = P.T;
    const int dpad = T.dpad, K = P.K;
    float* stage_mem = reinterpret_cast<float*>(smem_raw);
    float* qs = stage_mem + STAGES * kRowsPerStage * kPage;  // [dpad][TQ]
    uint64_t* lists = reinterpret_cast<uint64_t*>(qs + (size_t)dpad * TQ);
    int2* seed_rg = reinterpret_cast<int2*>(lists + (size_t)TQ * K);  // [TQ]
    int* a0s = reinterpret_cast<int*>(seed_rg + TQ);
    int* a1s = a0s + TQ;
    int* s0s = a1s + TQ;   // seed neighbourhood of each query (a superset of its approximate leaf)
    int* s1s = s0s + TQ;
    int* qids = s1s + TQ;
    float* exts = reinterpret_cast<float*>(qids + TQ);  // external bound of each query (FLT_MAX without one), shared memory: only the slow path reads it
    int* misc = reinterpret_cast<int*>(exts + TQ);  // [0] number of seed ranges
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(misc + 4);
    uint64_t* empty_bar = full_bar + STAGES;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile = blockIdx.x;
    const int q_begin = tile * TQ;
    const int nq_tile = min(TQ, P.nq - q_begin);

    // ---- tile set-up -------------------------------------------------------------------------------------
    if (tid == 0) {
        for (int s = 0; s < STAGES; s++) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], NW);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < TQ; i += blockDim.x) {
        int qid = (i < nq_tile) ? P.q_order[q_begin + i] : -1;
        qids[i] = qid;
        exts[i] = ext_thr_of(P, qid);
        int a0 = 0, a1 = 0, s0 = 0, s1 = 0;
        if (qid >= 0 && P.mode != kScanAll) {
            int leaf = P.q_leaf[qid];
            a0 = T.leaf_start[leaf];
            a1 = a0 + T.leaf_count[leaf];
            s0 = T.seed_start[leaf];
            s1 = T.seed_end[leaf];
        }
        a0s[i] = a0;
        a1s[i] = a1;
        s0s[i] = s0;
        s1s[i] = s1;
    }
    __syncthreads();
    for (int i = tid; i < TQ * dpad; i += blockDim.x) {
        int qi = i / dpad, j = i - qi * dpad;
        int qid = qids[qi];
        float v = (qid >= 0 && j < T.dim) ? P.q[(size_t)qid * T.dim + j] : 0.f;
        qs[j * TQ + qi] = v;
    }
    if (P.mode == kScanList) {
        for (int i = tid; i < TQ * K; i += blockDim.x) {
            int qi = i / K;
            lists[i] = (qi < nq_tile) ? P.state[(size_t)(q_begin + qi) * K + (i - qi * K)] : kEmptyKey;
        }
    } else {
        for (int i = tid; i < TQ * K; i += blockDim.x) lists[i] = kEmptyKey;
    }
    if (tid == 0) misc[0] = (P.mode == kScanAll) ? 0 : build_seed_ranges(s0s, s1s, nq_tile, P.pos_lo, P.pos_hi, seed_rg);
    __syncthreads();

    // ---- the page list of this tile (identical walk in the producer and in every consumer warp) ------------
    const int n_seed = misc[0];
    if (tid == 0) {
        misc[1] = 1;  // one range: every page (kScanAll; slot TQ-1 is free there: no seed ranges)
        if (P.mode == kScanAll) seed_rg[TQ - 1] = make_int2(T.page_lo, T.page_hi);
    }
    __syncthreads();
    const int2* list;
    const int* cnt;
    int n_groups = 1, stride = 0, n_skip = 0;
    if (P.mode == kScanAll) {
        list = &seed_rg[TQ - 1];
        cnt = &misc[1];
    } else if (P.mode == kScanSeed) {
        list = seed_rg;
        cnt = &misc[0];
    } else {
        list = P.ranges + (size_t)tile * kRangeCap;
        cnt = P.n_ranges + (size_t)tile * kMaxCuts;
        n_groups = P.range_groups;
        stride = kJobRanges;
        n_skip = P.lazy_seed ? 0 : n_seed;  // pages of the seed scan are already in the lists (unless the seed scan was lazy)
    }
    PageWalk walk(list, cnt, n_groups, stride, seed_rg, n_skip);
    const int n_slices = (dpad + kRowsPerStage - 1) / kRowsPerStage;
    int page;

    if (warp == NW) {
        // ================= producer: one lane streams page slices through the stage ring =================
        if (lane == 0) {
            uint32_t it = 0;
            while (walk.next(page)) {
                for (int s = 0; s < n_slices; s++, it++) {
                    int st = it % STAGES;
                    uint32_t ph = (it / STAGES) & 1;
                    mbar_wait_backoff(&empty_bar[st], ph ^ 1);
                    int rows = min(kRowsPerStage, dpad - s * kRowsPerStage);
                    uint32_t bytes = (uint32_t)rows * kPage * 4;
                    mbar_arrive_expect_tx(&full_bar[st], bytes);
                    bulk_g2s(stage_mem + (size_t)st * kRowsPerStage * kPage,
                             T.pages + ((size_t)page * dpad + (size_t)s * kRowsPerStage) * kPage, bytes, &full_bar[st]);
                }
            }
        }
        return;
    }

    // ================= consumers: warp owns 8 queries, lane owns 8 points of every page =====================
    const int qw0 = warp * kQPerWarp;  // first tile-local query of this warp
    uint64_t* my_lists = lists + (size_t)qw0 * K;
    // fast-path bound per query: K-th d2 widened by 2^-20 (covers the rare double-rounding difference between the
    // packed FMA chain and the reference arithmetic); padded query slots get -1 and never collect.
    float thr[kQPerWarp];
#pragma unroll
    for (int qi = 0; qi < kQPerWarp; qi++) {
        float kth = __uint_as_float((uint32_t)(my_lists[(size_t)qi * K + K - 1] >> 32));
        thr[qi] = (qw0 + qi < nq_tile) ? fminf(fminf(kth * 1.000001f, 3.402823466e+38f), exts[qw0 + qi]) : -1.f;
    }
    const bool brute = (P.mode == kScanAll);
    constexpr bool lazy = LAZY;  // (launched for kScanSeed with P.lazy_seed only)
    // Narrow points (the whole page is ONE stage): the stage is handed back to the producer only after the page's exact
    // path, which then reads its candidates' coordinates from shared memory instead of L2 (a ~700-cycle round trip per
    // batch of 8 loads, most of the seed scan's time).
    const bool hold_stage = (n_slices == 1);
    unsigned long long n_pairs = 0;
    unsigned int n_exact = 0, n_ins = 0;

    uint32_t it = 0;
    while (walk.next(page)) {
        uint64_t acc[kQPerWarp][4];

        // one dimension (row of the stage) for the warp's 8 queries x this lane's 8 points.  FIRST: acc = d*d, which
        // rounds exactly like fma(d, d, +0) and saves zeroing 64 accumulators per page.
        auto row_step = [&](const float* sp_row, const float* qp_row, auto first) {
            float4 pa = *reinterpret_cast<const float4*>(sp_row);
            float4 pb = *reinterpret_cast<const float4*>(sp_row + 128);
            float4 qa = *reinterpret_cast<const float4*>(qp_row);
            float4 qb = *reinterpret_cast<const float4*>(qp_row + 4);
            uint64_t pv[4] = {pack2(pa.x, pa.y), pack2(pa.z, pa.w), pack2(pb.x, pb.y), pack2(pb.z, pb.w)};
            float qv[8] = {qa.x, qa.y, qa.z, qa.w, qb.x, qb.y, qb.z, qb.w};
#pragma unroll
            for (int qi = 0; qi < kQPerWarp; qi++) {
                uint64_t qq = pack2(qv[qi], qv[qi]);  // becomes the scalar-broadcast operand form of FADD2
#pragma unroll
                for (int pp = 0; pp < 4; pp++) {
                    uint64_t d = sub2(pv[pp], qq);
                    acc[qi][pp] = decltype(first)::value ? mul2(d, d) : fma2(d, d, acc[qi][pp]);
                }
            }
        };

        // one stage = one slice of <= 16 dimensions of the page; the first slice starts the accumulators
        auto stage_step = [&](int s, auto first) {
            int st = it % STAGES;
            uint32_t ph = (it / STAGES) & 1;
            mbar_wait(&full_bar[st], ph);
            const float* sp = stage_mem + (size_t)st * kRowsPerStage * kPage + 4 * lane;
            const float* qp = qs + (size_t)s * kRowsPerStage * TQ + qw0;
            row_step(sp, qp, first);
            if (FULL) {
#pragma unroll
                for (int row = 1; row < kRowsPerStage; row++) row_step(sp + row * kPage, qp + row * TQ, std::false_type{});
            } else {
                int rows = min(kRowsPerStage, dpad - s * kRowsPerStage);
#pragma unroll 1
                for (int row = 1; row < rows; row++) row_step(sp + row * kPage, qp + row * TQ, std::false_type{});
            }
            __syncwarp();
            if (!hold_stage && lane == 0) mbar_arrive(&empty_bar[st]);
            it++;
        };
        const int held_st = it % STAGES;  // (only meaningful when hold_stage)
        stage_step(0, std::true_type{});
#pragma unroll 1
        for (int s = 1; s < n_slices; s++) stage_step(s, std::false_type{});
        n_pairs += (unsigned long long)min(kPage, T.n_points - page * kPage);

        // ---- bound bootstrap: while a query's list is not full its bound is +inf and EVERY point of the page would go
        // through the exact path (that start-up used to dominate the seed scan).  Instead, a bisection on the bit patterns of
        // the page's 256 fast-path distances finds a value t with at least K valid points at or below it; points beyond
        // t(1+2^-20) are strictly farther than K others (the fast path is within 2 ulp of the reference arithmetic), so they
        // cannot be among the K nearest and are dropped without an exact evaluation.  thr[] is warp-uniform.
        {
            bool boot = false;
#pragma unroll
            for (int qi = 0; qi < kQPerWarp; qi++) boot |= thr[qi] >= 3.0e38f;
            if (boot && K <= kPage) {
                uint32_t ok = 0;  // points of this lane that may enter a list (inside the shard; padded ones are +inf anyway)
#pragma unroll
                for (int e = 0; e < 8; e++) {
                    const int pos = page * kPage + ((e >> 2) ? 128 : 0) + 4 * lane + (e & 3);
                    if (pos >= P.pos_lo && pos < P.pos_hi) ok |= 1u << e;
                }
#pragma unroll
                for (int qi = 0; qi < kQPerWarp; qi++) {
                    if (thr[qi] >= 3.0e38f) {
                        uint32_t v[8];
#pragma unroll
                        for (int pp = 0; pp < 4; pp++) {
                            float lo, hi;
                            unpack2(acc[qi][pp], lo, hi);
                            v[2 * pp] = ((ok >> (2 * pp)) & 1u) ? __float_as_uint(lo) : 0x7f800000u;
                            v[2 * pp + 1] = ((ok >> (2 * pp + 1)) & 1u) ? __float_as_uint(hi) : 0x7f800000u;
                        }
                        uint32_t vmin = 0xffffffffu, vmax = 0;
                        int cnt = 0;
#pragma unroll
                        for (int e = 0; e < 8; e++) {
                            const bool fin = v[e] < 0x7f7fffffu;  // finite non-negative fast-path distance (NaN patterns are above)
                            vmin = min(vmin, fin ? v[e] : 0xffffffffu);
                            vmax = max(vmax, fin ? v[e] : 0u);
                            cnt += fin ? 1 : 0;
                        }
                        cnt = __reduce_add_sync(0xffffffffu, cnt);
                        if (cnt >= K) {
                            uint32_t lo = __reduce_min_sync(0xffffffffu, vmin), hi = __reduce_max_sync(0xffffffffu, vmax);
#pragma unroll 1
                            for (int iter = 0; iter < 18 && lo < hi; iter++) {
                                const uint32_t mid = lo + ((hi - lo) >> 1);
                                int c = 0;
#pragma unroll
                                for (int e = 0; e < 8; e++) c += (v[e] <= mid) ? 1 : 0;
                                c = __reduce_add_sync(0xffffffffu, c);
                                if (c >= K) hi = mid;       // invariant: at least K valid points are <= hi
                                else lo = mid + 1;
                            }
                            thr[qi] = fminf(__uint_as_float(hi) * 1.000001f, 3.0e38f);
                        }
                    }
                }
            }
        }

        // ---- candidate filter: anything within the widened K-th distance goes to the exact path ----------
        bool any = false;
#pragma unroll
        for (int qi = 0; qi < kQPerWarp; qi++)
#pragma unroll
            for (int pp = 0; pp < 4; pp++) {
                float lo, hi;
                unpack2(acc[qi][pp], lo, hi);
                any |= (lo <= thr[qi]) | (hi <= thr[qi]);
            }
        if (__any_sync(0xffffffffu, any)) {
            // per-lane candidate mask: bit 8*qi+e  <->  query qi, point e of this lane
            // (e = 2*pp+h: positions 4*lane+{0..3} for e<4, 128+4*lane+{0..3} for e>=4)
            uint32_t cm_lo = 0, cm_hi = 0;
#pragma unroll
            for (int qi = 0; qi < kQPerWarp; qi++)
#pragma unroll
                for (int pp = 0; pp < 4; pp++) {
                    float lo, hi;
                    unpack2(acc[qi][pp], lo, hi);
                    uint32_t b = (lo <= thr[qi] ? 1u : 0u) | (hi <= thr[qi] ? 2u : 0u);
                    if (qi < 4) cm_lo |= b << (8 * qi + 2 * pp);
                    else cm_hi |= b << (8 * (qi - 4) + 2 * pp);
                }
            // Exact path.  One ROUND: every lane re-evaluates ONE of its own candidates -- of whichever of the warp's 8 queries --
            // with the reference arithmetic; then the round's results are inserted query by query.  Pooling the queries matters
            // because the latency of an evaluation (row loads + the sequential FP64 chain) is paid once per round whatever the
            // number of busy lanes: a page leaves ~15 candidates per query, i.e. one or two rounds per query taken one query at a
            // time (8-16 rounds per page and warp), but only max-per-lane ~ 5-7 rounds when the 64 (query, point) bits of a lane
            // are drained together.
            uint64_t cm = ((uint64_t)cm_hi << 32) | cm_lo;
#pragma unroll 1
            while (__any_sync(0xffffffffu, cm != 0)) {
                bool cand = false;
                float d2 = 0.f;
                int pos = 0, cqi = -1;
                if (cm) {
                    const int b = __ffsll((long long)cm) - 1;
                    cm &= cm - 1;
                    cqi = b >> 3;
                    const int e = b & 7;
                    const int in_page = ((e >> 2) ? 128 : 0) + 4 * lane + (e & 3);
                    pos = page * kPage + in_page;
                    const float* qq = qs + qw0 + cqi;
                    // loads first (8 dims in flight), then the sequential reference chain: the chain would otherwise
                    // serialise one round trip per dimension.  A held stage is read from shared memory (dim-major);
                    // otherwise the point's row comes from the row-major copy (two 16-byte loads per 8 dims)
                    // lazy seed scan: a plain FP32 evaluation (two FMA chains) stands in for the reference chain; what is left in
                    // the lists at the end is evaluated exactly (below)
                    float f0 = 0.f, f1 = 0.f;
                    if (hold_stage) {
                        const float* pg = stage_mem + (size_t)held_st * kRowsPerStage * kPage + in_page;
                        for (int j0 = 0; j0 < T.dim; j0 += 8) {
                            float pv[8];
#pragma unroll
                            for (int u = 0; u < 8; u++) pv[u] = (j0 + u < T.dim) ? pg[(size_t)(j0 + u) * kPage] : 0.f;
#pragma unroll
                            for (int u = 0; u < 8; u++)
                                if (j0 + u < T.dim) {
                                    const float d = pv[u] - qq[(j0 + u) * TQ];
                                    if (!lazy) d2 = ref_acc_sq(d2, d);
                                    else if (u & 1) f1 = fmaf(d, d, f1);
                                    else f0 = fmaf(d, d, f0);
                                }
                        }
                    } else {
                        const float4* pr = reinterpret_cast<const float4*>(T.rows + (size_t)pos * dpad);
                        prefetch_row(pr, dpad);
                        for (int j0 = 0; j0 < T.dim; j0 += 8) {
                            const float4 x = __ldg(pr + (j0 >> 2));
                            const float4 y = (j0 + 4 < dpad) ? __ldg(pr + (j0 >> 2) + 1) : make_float4(0.f, 0.f, 0.f, 0.f);
                            const float pv[8] = {x.x, x.y, x.z, x.w, y.x, y.y, y.z, y.w};
#pragma unroll
                            for (int u = 0; u < 8; u++)
                                if (j0 + u < T.dim) {
                                    const float d = pv[u] - qq[(j0 + u) * TQ];
                                    if (!lazy) d2 = ref_acc_sq(d2, d);
                                    else if (u & 1) f1 = fmaf(d, d, f1);
                                    else f0 = fmaf(d, d, f0);
                                }
                        }
                    }
                    if (lazy) d2 = f0 + f1;
                    n_exact++;
                    cand = (d2 < 3.402823466e+38f) && (pos >= P.pos_lo) && (pos < P.pos_hi);
                }
                uint32_t low = 0xffffffffu;
                if (cand) {
                    const int a0 = a0s[qw0 + cqi], a1 = a1s[qw0 + cqi];
                    low = brute ? (uint32_t)T.orig_idx[pos] : (((pos >= a0 && pos < a1) ? 0u : 0x80000000u) | (uint32_t)pos);
                }
                // the round's results, query by query.  Smallest first: each step takes the warp-minimum key of the query's remaining
                // candidates (two REDUX); as soon as the minimum cannot enter the list none of the others can, so rejected candidates
                // cost nothing (while the lists are still filling, most candidates are rejected by the ones inserted before them).
#pragma unroll 1
                for (int qi = 0; qi < kQPerWarp; qi++) {
                    const bool mine = cand && cqi == qi;
                    if (!__any_sync(0xffffffffu, mine)) continue;
                    uint64_t* L = my_lists + (size_t)qi * K;
                    uint32_t hi = mine ? __float_as_uint(d2) : 0xffffffffu;  // valid d2 bits are < 0x7f7fffff
                    for (;;) {
                        const uint32_t mhi = __reduce_min_sync(0xffffffffu, hi);
                        if (mhi == 0xffffffffu) break;
                        const uint32_t mlo = __reduce_min_sync(0xffffffffu, hi == mhi ? low : 0xffffffffu);
                        const uint64_t key = ((uint64_t)mhi << 32) | mlo;
                        if (key >= L[K - 1]) break;  // nothing smaller is left
                        n_ins += list_insert(L, K, key, lane);
                        if (hi == mhi && low == mlo) hi = 0xffffffffu;  // consumed (keys are unique)
                    }
                }
            }
            float new_thr[kQPerWarp];
#pragma unroll
            for (int qi = 0; qi < kQPerWarp; qi++) {
                const float kth = __uint_as_float((uint32_t)(my_lists[(size_t)qi * K + K - 1] >> 32));
                new_thr[qi] = (qw0 + qi < nq_tile) ? fminf(fminf(kth * 1.000001f, 3.402823466e+38f), exts[qw0 + qi]) : -1.f;
            }
#pragma unroll
            for (int qi = 0; qi < kQPerWarp; qi++) thr[qi] = new_thr[qi];
        }
        if (hold_stage) {
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[held_st]);
        }
    }

    // ---- lazy seed scan: the entries that are left get their reference distances (the warp's 8 lists together, 32 entries
    // per round), then every list is put back in order (the fast and the exact order differ on near-ties only)
    if (lazy) {
        __syncwarp();
        const int total = kQPerWarp * K;
        for (int e0 = 0; e0 < total; e0 += 32) {
            const int e = e0 + lane;
            if (e < total) {
                const int qi = e / K;
                uint64_t* slot = my_lists + e;  // (= my_lists + qi * K + k)
                const uint64_t key = *slot;
                if (qw0 + qi < nq_tile && key != kEmptyKey) {
                    const int pos = (int)((uint32_t)key & 0x7fffffffu);
                    const float* qq = qs + qw0 + qi;
                    float d2 = 0.f;
                    if (T.rows) {
                        const float4* pr = reinterpret_cast<const float4*>(T.rows + (size_t)pos * dpad);
                        prefetch_row(pr, dpad);
                        for (int j0 = 0; j0 < T.dim; j0 += 8) {
                            const float4 x = __ldg(pr + (j0 >> 2));
                            const float4 y = (j0 + 4 < dpad) ? __ldg(pr + (j0 >> 2) + 1) : make_float4(0.f, 0.f, 0.f, 0.f);
                            const float pv[8] = {x.x, x.y, x.z, x.w, y.x, y.y, y.z, y.w};
#pragma unroll
                            for (int u = 0; u < 8; u++)
                                if (j0 + u < T.dim) d2 = ref_acc_sq(d2, pv[u] - qq[(j0 + u) * TQ]);
                        }
                    } else {
                        const float* pg = T.pages + (size_t)(pos / kPage) * dpad * kPage + (pos % kPage);
                        for (int j0 = 0; j0 < T.dim; j0 += 8) {
                            float pv[8];
#pragma unroll
                            for (int u = 0; u < 8; u++) pv[u] = (j0 + u < T.dim) ? __ldg(pg + (size_t)(j0 + u) * kPage) : 0.f;
#pragma unroll
                            for (int u = 0; u < 8; u++)
                                if (j0 + u < T.dim) d2 = ref_acc_sq(d2, pv[u] - qq[(j0 + u) * TQ]);
                        }
                    }
                    *slot = (d2 < 3.402823466e+38f) ? (((uint64_t)__float_as_uint(d2) << 32) | (uint32_t)key) : kEmptyKey;
                    n_exact++;
                }
            }
        }
        __syncwarp();
        for (int qi = 0; qi < kQPerWarp; qi++) {
            if (qw0 + qi >= nq_tile) break;
            uint64_t* L = my_lists + (size_t)qi * K;
            bool bad = false;
            for (int k = lane; k + 1 < K; k += 32) bad |= L[k] > L[k + 1];
            if (!__any_sync(0xffffffffu, bad)) continue;
            // stable rank sort (keys of real points are unique; empty slots compare equal and keep their order)
            uint64_t mine[4];
            int rank[4];
#pragma unroll
            for (int s = 0; s < 4; s++) {
                const int k = s * 32 + lane;
                mine[s] = (k < K) ? L[k] : kEmptyKey;
                rank[s] = k;
                if (k < K) {
                    int r = 0;
                    for (int j = 0; j < K; j++) {
                        const uint64_t o = L[j];
                        r += (o < mine[s] || (o == mine[s] && j < k)) ? 1 : 0;
                    }
                    rank[s] = r;
                }
            }
            __syncwarp();
#pragma unroll
            for (int s = 0; s < 4; s++)
                if (s * 32 + lane < K) L[rank[s]] = mine[s];
            __syncwarp();
        }
    }

    // ---- results ------------------------------------------------------------------------------------------
    __syncwarp();
    for (int qi = 0; qi < kQPerWarp; qi++) {
        if (qw0 + qi >= nq_tile) break;
        const uint64_t* L = my_lists + (size_t)qi * K;
        const int qid = qids[qw0 + qi];
        if (P.write_mode == kWriteFinal) {
            for (int k = lane; k < K; k += 32) {
                uint64_t key = L[k];
                int idx = -1;
                float d2 = 3.402823466e+38f;
                if (key != kEmptyKey) {
                    uint32_t low = (uint32_t)key;
                    idx = brute ? (int)low : T.orig_idx[low & 0x7fffffffu];
                    d2 = __uint_as_float((uint32_t)(key >> 32));
                }
                P.idx_out[(size_t)qid * K + k] = idx;
                P.d2_out[(size_t)qid * K + k] = d2;
                if (P.dist_out) P.dist_out[(size_t)qid * K + k] = __fsqrt_rn(d2);
            }
        } else if (P.write_mode == kWriteState) {
            for (int k = lane; k < K; k += 32) P.state[(size_t)(q_begin + qw0 + qi) * K + k] = L[k];
        } else {
            uint64_t* dst = keys_dst(P, qid);
            for (int k = lane; k < K; k += 32) dst[k] = L[k];
        }
    }
    if (P.counters) {
        int nvalid = max(0, min(kQPerWarp, nq_tile - qw0));
        for (int o = 16; o; o >>= 1) n_exact += __shfl_xor_sync(0xffffffffu, n_exact, o);
        if (lane == 0) {
            atomicAdd(&P.counters[0], n_pairs * (unsigned long long)nvalid);
            atomicAdd(&P.counters[1], (unsigned long long)n_exact);
            atomicAdd(&P.counters[2], (unsigned long long)n_ins);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// K2: traversal / pruning.  One warp per tile of TQ queries (lane = 2 queries for TQ=64), walking the tree in DFS
// pre-order (= array order).  The dot product of a query with the normal of each edge on the current root->node
// path is memoised per level in shared memory (the reference's `_pro` memo, Code/ConvexNode.h:151-206, but per
// path level instead of per node id), so a node costs one D-dim dot + depth adds per query.  A subtree is
// skipped when the lower bound of EVERY query of the tile exceeds that query's K-th distance after the seed scan
// (bound: Code/ConvexNode.h:113-144, max over active constraints of -(a.q+b); it holds for every descendant
// because a node's half-spaces contain all its points, Code/Convex_Tree.h:755-771).  Unlike the reference the
// test carries a rounding margin, so the search stays provably exact.  Output: ascending page ranges per tile.
// ---------------------------------------------------------------------------------------------------------
struct TraverseParams {
    DevTree T;
    const float* q;
    const int* q_order;
    const int* q_leaf;
    int nq, K, tq;           // tq = queries per tile of the leaf-scan kernel (<= 32*QL)
    const uint64_t* state;   // [nq][K] keys after the seed scan
    const float* ext_bound;  // [nq] by query id or null (ScanParams::ext_bound)
    int2* ranges;            // [n_tiles][kMaxCuts][kJobRanges]
    int* n_ranges;           // [n_tiles][kMaxCuts]
    int leaf_lo, leaf_hi;    // only leaves in [leaf_lo, leaf_hi) are emitted (sharded-dataset mode)
    int sel_mode, sel_step;  // tile subset: 0 all tiles, 1 tiles with t % step == 0 (sample), 2 the other tiles
    int ball_only;           // 1: only the bounding-ball test (prefilter scan: a per-query test of a leaf costs about as
                             //    much as letting the tensor core scan its page, so only the ~free tile-level test is used)
    unsigned long long* counters;  // [4] node tests (per query), [5] pages listed
};

__host__ __device__ inline size_t traverse_smem_bytes(int warps, int tq, int dpad) {
    return (size_t)warps * ((size_t)dpad * tq * 4 + (size_t)kMaxLevel * tq * 4 + (size_t)(2 * dpad + 2 * kMaxLevel) * 4);
}

// i-th tile of a tile subset (see TraverseParams::sel_mode)
__host__ __device__ inline int tile_of_selection(int i, int mode, int step) {
    if (mode == 1) return i * step;
    if (mode == 2) return i + i / (step - 1) + 1;
    return i;
}
__host__ __device__ inline int selection_size(int n_tiles, int mode, int step) {
    int sampled = (n_tiles + step - 1) / step;
    return mode == 1 ? sampled : mode == 2 ? n_tiles - sampled : n_tiles;
}

// Lists every page of [page_lo, page_hi) for the tiles of a subset (used when a sampled traversal shows that the
// bounds prune next to nothing: evaluating them for the remaining tiles would cost more than it saves).
__global__ void k_list_all(int2* __restrict__ ranges, int* __restrict__ n_ranges, int n_tiles, int n_cuts, int mode, int step, int page_lo,
                           int page_hi) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= selection_size(n_tiles, mode, step)) return;
    int tile = tile_of_selection(i, mode, step);
    if (tile >= n_tiles) return;
    for (int c = 0; c < n_cuts; c++) n_ranges[(size_t)tile * kMaxCuts + c] = (c == 0) ? 1 : 0;
    ranges[(size_t)tile * kMaxCuts * kJobRanges] = make_int2(page_lo, page_hi);
}

template <int QL>  // queries per lane: a tile holds up to 32*QL queries
__global__ void __launch_bounds__(128) k_traverse(const TraverseParams P) {
    constexpr int TQ = 32 * QL;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const DevTree& T = P.T;
    const int dpad = T.dpad;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int n_tiles = (P.nq + P.tq - 1) / P.tq;
    const long long job = (long long)blockIdx.x * nwarps + warp;  // one warp per (selected tile, subtree job)
    const int sel = (int)(job / T.n_cuts), cut = (int)(job % T.n_cuts);
    const int tile = tile_of_selection(sel, P.sel_mode, P.sel_step);
    if (tile >= n_tiles) return;
    float* qs = reinterpret_cast<float*>(smem_raw) + (size_t)warp * ((size_t)dpad * TQ + (size_t)kMaxLevel * TQ + 2 * dpad + 2 * kMaxLevel);  // [dpad][TQ]
    float* dots = qs + (size_t)dpad * TQ;  // [level][TQ]
    float* cen = dots + (size_t)kMaxLevel * TQ;  // [dpad] centroid of the tile's queries
    float* cdots = cen + dpad;                   // [level] the centroid's dot memo
    float* nrm_s = cdots + kMaxLevel;            // [dpad] the visited node's normal, staged once per node
    float* beta_s = nrm_s + dpad;                // [level] the visited node's offsets
    const int q_begin = tile * P.tq;  // tile stride = the leaf-scan kernel's tile (<= TQ slots used)
    const int nq_tile = min(P.tq, P.nq - q_begin);

    float thr2[QL], absm[QL];
    bool valid[QL];
#pragma unroll
    for (int u = 0; u < QL; u++) {
        int qi = lane + 32 * u;
        thr2[u] = 0.f;
        absm[u] = 0.f;
        valid[u] = false;  // padded slot: never keeps a subtree alive
        if (qi < nq_tile) {
            valid[u] = true;
            int qid = P.q_order[q_begin + qi];
            float n2 = 0.f;
            for (int j = 0; j < dpad; j++) {
                float v = (j < T.dim) ? P.q[(size_t)qid * T.dim + j] : 0.f;
                qs[j * TQ + qi] = v;
                n2 = fmaf(v, v, n2);
            }
            // rounding margin of the bound: |error| <= c * D * 2^-24 * (||q|| + max||x||), see DESIGN.md
            absm[u] = 16.f * (float)(T.dim + 2) * 5.9604645e-8f * (sqrtf(n2) * 1.0001f + T.xnorm_max);
            uint32_t kb = (uint32_t)(P.state[(size_t)(q_begin + qi) * P.K + P.K - 1] >> 32);
            float kth2 = __uint_as_float(kb);
            if (P.ext_bound) kth2 = fminf(kth2, P.ext_bound[qid]);
            thr2[u] = kth2 * 1.00001f;  // K-th d2 after the seed (FLT_MAX -> inf: never prunes)
        } else {
            for (int j = 0; j < dpad; j++) qs[j * TQ + qi] = 0.f;
        }
    }
    __syncwarp();
    // Bounding ball of the tile's queries (centre = centroid, radius r): the bound of a node is 1-Lipschitz in the query
    // (its normals have unit length), so bound(q) >= bound(centre) - r for every query of the tile.  One centre test
    // prunes a subtree for all queries at 1/(32*QL) of the cost of the per-query tests; those run only when it fails.
    float r_tile = 0.f, k_max = 0.f, absm_c = 0.f;
    {
        for (int j = 0; j < dpad; j++) {
            float sj = 0.f;
#pragma unroll
            for (int u = 0; u < QL; u++) sj += valid[u] ? qs[j * TQ + lane + 32 * u] : 0.f;
            for (int o = 16; o; o >>= 1) sj += __shfl_xor_sync(0xffffffffu, sj, o);
            if (lane == 0) cen[j] = sj / (float)nq_tile;
        }
        __syncwarp();
        float cn2 = 0.f;
        for (int j = 0; j < dpad; j++) cn2 = fmaf(cen[j], cen[j], cn2);
#pragma unroll
        for (int u = 0; u < QL; u++) {
            if (!valid[u]) continue;
            float d2 = 0.f;
            for (int j = 0; j < dpad; j++) {
                float d = qs[j * TQ + lane + 32 * u] - cen[j];
                d2 = fmaf(d, d, d2);
            }
            r_tile = fmaxf(r_tile, d2);
            k_max = fmaxf(k_max, thr2[u]);
        }
        for (int o = 16; o; o >>= 1) {
            r_tile = fmaxf(r_tile, __shfl_xor_sync(0xffffffffu, r_tile, o));
            k_max = fmaxf(k_max, __shfl_xor_sync(0xffffffffu, k_max, o));
        }
        r_tile = sqrtf(r_tile) * 1.0001f + 1e-30f;
        absm_c = 16.f * (float)(T.dim + 2) * 5.9604645e-8f * (sqrtf(cn2) * 1.0001f + T.xnorm_max);
    }

    int2* out = P.ranges + ((size_t)tile * kMaxCuts + cut) * kJobRanges;
    int n_out = 0, cur_b = -1, cur_e = -1;  // open range [cur_b, cur_e)
    unsigned long long tests = 0, listed = 0;
    // Append [b,e) (ascending).  When the job runs out of slots the last range is widened over the gap: the extra
    // pages only cost time (every scanned pair is evaluated exactly and list insertion is idempotent).
    auto flush = [&](int b, int e) {
        if (n_out == kJobRanges) {
            listed += (unsigned long long)(e - out[n_out - 1].y);
            if (lane == 0) out[n_out - 1].y = e;
        } else {
            listed += (unsigned long long)(e - b);
            if (lane == 0) out[n_out] = make_int2(b, e);
            n_out++;
        }
        __syncwarp();
    };
    // Visit one node: memoise the dot of every query with the node's own normal at its level, evaluate the bound
    // over the node's constraints; true when EVERY query of the tile can skip the subtree.
    auto visit = [&](int n, const NodeRec& nr) -> bool {
        if (nr.depth == 0) return false;
        bool prune_all = true;
        const int k = nr.depth;
        const float* nrm = T.node_normal + (size_t)n * dpad;
        const float* beta = T.node_beta + nr.beta_off;
        {
            // centre test: lanes over dimensions for the dot, lanes over constraints for the bound
            // (the node's normal and offsets are read from global memory once, coalesced, and staged in shared memory:
            // the per-query loops below would otherwise expose one L1/L2 round trip per element)
            float part = 0.f;
            for (int j = lane; j < dpad; j += 32) {
                const float a = nrm[j];
                nrm_s[j] = a;
                part = fmaf(a, cen[j], part);
            }
            const int kc = min(k, kMaxLevel);  // kMaxLevel == 32: one constraint per lane
            const float my_beta = (lane < kc) ? beta[lane] : 0.f;
            beta_s[lane] = my_beta;
            for (int o = 16; o; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
            if (k <= kMaxLevel && lane == 0) cdots[k - 1] = part;
            __syncwarp();
            float t = (lane < kc) ? fmaxf(-(cdots[lane] + my_beta), 0.f) : 0.f;
            const float lbc = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(t)));
            const float safe_c = lbc * 0.9999f - absm_c - r_tile * 1.0001f;
            tests += 1;
            if (safe_c > 0.f && safe_c * safe_c > k_max) return true;
            if (P.ball_only) return false;
        }
#pragma unroll
        for (int u = 0; u < QL; u++) {
            int qi = lane + 32 * u;
            float d0 = 0.f, d1 = 0.f;
            for (int j = 0; j < dpad; j += 4) {
                float4 a = *reinterpret_cast<const float4*>(nrm_s + j);
                d0 = fmaf(a.x, qs[(j + 0) * TQ + qi], d0);
                d1 = fmaf(a.y, qs[(j + 1) * TQ + qi], d1);
                d0 = fmaf(a.z, qs[(j + 2) * TQ + qi], d0);
                d1 = fmaf(a.w, qs[(j + 3) * TQ + qi], d1);
            }
            const float d = d0 + d1;
            if (k <= kMaxLevel) dots[(k - 1) * TQ + qi] = d;
            float lb0 = 0.f, lb1 = 0.f;
            const int kk = min(k, kMaxLevel);
            int j = 0;
            for (; j + 1 < kk; j += 2) {
                lb0 = fmaxf(lb0, -(dots[j * TQ + qi] + beta_s[j]));
                lb1 = fmaxf(lb1, -(dots[(j + 1) * TQ + qi] + beta_s[j + 1]));
            }
            if (j < kk) lb0 = fmaxf(lb0, -(dots[j * TQ + qi] + beta_s[j]));
            const float safe = fmaxf(lb0, lb1) * 0.9999f - absm[u];
            const bool prune = !valid[u] || ((safe > 0.f) && (safe * safe > thr2[u]));
            prune_all &= prune;
        }
        tests += (unsigned long long)nq_tile;
        return __all_sync(0xffffffffu, prune_all);
    };
    auto emit_leaf = [&](const NodeRec& nr) {
        if (nr.leaf >= P.leaf_lo && nr.leaf < P.leaf_hi && nr.pos1 > nr.pos0) {
            int pb = nr.pos0 / kPage, pe = (nr.pos1 + kPage - 1) / kPage;
            if (cur_b < 0) {
                cur_b = pb;
                cur_e = pe;
            } else if (pb <= cur_e) {
                cur_e = max(cur_e, pe);
            } else {
                flush(cur_b, cur_e);
                cur_b = pb;
                cur_e = pe;
            }
        }
    };

    // 1. the path root -> cut node (fills the dot memo of the levels above the subtree; any pruned ancestor ends the job)
    const int* path = T.cut_path + (size_t)cut * (T.cut_depth + 1);
    int cut_node = 0;
    bool dead = false;
    for (int i = 0; i <= T.cut_depth && !dead; i++) {
        int n = path[i];
        if (n < 0) break;
        cut_node = n;
        dead = visit(n, T.nodes[n]);
    }
    // 2. the subtree of the cut node in pre-order (= array order), skipping pruned subtrees
    if (!dead) {
        const NodeRec root = T.nodes[cut_node];
        if (root.leaf >= 0) {
            emit_leaf(root);
        } else {
            int n = cut_node + 1;
            const int end = cut_node + root.size;
            while (n < end) {
                const NodeRec nr = T.nodes[n];
                if (visit(n, nr)) {
                    n += nr.size;
                    continue;
                }
                if (nr.leaf >= 0) emit_leaf(nr);
                n += 1;
            }
        }
    }
    if (cur_b >= 0) flush(cur_b, cur_e);
    if (lane == 0) {
        P.n_ranges[(size_t)tile * kMaxCuts + cut] = n_out;
        if (P.counters) {
            atomicAdd(&P.counters[4], tests);
            atomicAdd(&P.counters[5], listed * (unsigned long long)nq_tile);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// merge of per-shard partial lists (sharded-dataset mode): n_lists sorted key lists per query -> final arrays
// thread per query; lists laid out [n_lists][nq][K].
// ---------------------------------------------------------------------------------------------------------
__global__ void k_merge_partial(DevTree T, const uint64_t* __restrict__ keys, int n_lists, int nq, int K, int brute_keys, int* __restrict__ idx_out,
                                float* __restrict__ d2_out, float* __restrict__ dist_out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nq) return;
    int head[kMaxShards];
    for (int l = 0; l < n_lists; l++) head[l] = 0;
    for (int k = 0; k < K; k++) {
        uint64_t best = ~0ull;
        int bl = -1;
        for (int l = 0; l < n_lists; l++) {
            if (head[l] >= K) continue;
            uint64_t v = keys[((size_t)l * nq + i) * K + head[l]];
            if (v < best) { best = v; bl = l; }
        }
        int idx = -1;
        float d2 = 3.402823466e+38f;
        if (bl >= 0) {
            // a key identifies a point: lists that started from the same state (page segments of one tile) hold duplicates
            for (int l = 0; l < n_lists; l++)
                if (head[l] < K && keys[((size_t)l * nq + i) * K + head[l]] == best) head[l]++;
            if (best != kEmptyKey) {
                idx = brute_keys ? (int)(uint32_t)best : T.orig_idx[(uint32_t)best & 0x7fffffffu];  // brute-force keys carry the original index
                d2 = __uint_as_float((uint32_t)(best >> 32));
            }
        }
        idx_out[(size_t)i * K + k] = idx;
        d2_out[(size_t)i * K + k] = d2;
        if (dist_out) dist_out[(size_t)i * K + k] = __fsqrt_rn(d2);
    }
}

// bound[qid] = K-th squared distance of the query's list after the seed scan (FLT_MAX while the list is not full)
__global__ void k_extract_bounds(const uint64_t* __restrict__ state, const int* __restrict__ q_order, int nq, int K, float* __restrict__ bound) {
    int slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= nq) return;
    bound[q_order[slot]] = __uint_as_float((uint32_t)(state[(size_t)slot * K + K - 1] >> 32));
}
// out[i] = min over the n arrays (sharded-dataset mode: the seed bounds of all shards; peer pointers are read over NVLink)
struct BoundPtrs {
    const float* p[kMaxShards];
};
__global__ void k_min_bounds(BoundPtrs B, int n, int nq, float* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nq) return;
    float m = 3.402823466e+38f;
    for (int l = 0; l < n; l++) m = fminf(m, B.p[l][i]);
    out[i] = m;
}
// dst[i] = src[i] (or FLT_MAX: no external bound)
__global__ void k_init_bounds(float* __restrict__ dst, const float* __restrict__ src, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src ? src[i] : 3.402823466e+38f;
}
// keeps only the keys whose position lies in [pos_lo, pos_hi) (order preserved), or empties the lists (pos_hi <= pos_lo)
__global__ void k_clip_state(uint64_t* __restrict__ state, int nq, int K, int pos_lo, int pos_hi) {
    int slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= nq) return;
    uint64_t* L = state + (size_t)slot * K;
    int w = 0;
    for (int k = 0; k < K; k++) {
        const uint64_t key = L[k];
        const int pos = (int)((uint32_t)key & 0x7fffffffu);
        if (key != kEmptyKey && pos >= pos_lo && pos < pos_hi) L[w++] = key;
    }
    for (; w < K; w++) L[w] = kEmptyKey;
}

}  // namespace scht
