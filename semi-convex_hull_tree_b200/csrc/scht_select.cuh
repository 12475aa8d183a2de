// Page selection for the tensor-core prefilter scan: which SCAN pages can hold a neighbour of some query of a tile?
//
// The scan-order image (ScanImage, scht_types.cuh) keeps the resident points in spatially tight groups; every group and
// every scan page (256 slots) has a bounding ball (centre c, radius r).  For a query q with current K-th distance kth
// (after the seed scan of its approximate leaf, Code/CPU_Convex_Tree.h:181-217) the triangle inequality gives the lower
// bound  |q - p| >= |q - c| - r  for every point p of the ball, so the page is skipped when |q - c| - r > kth -- the
// same role as the reference's hyper-plane test (Code/ConvexNode.h:113-144: skip a leaf whose lower bound exceeds the
// K-th distance), with a bound that stays tight on clustered data where a leaf's half-spaces are loose.  All
// comparisons carry the rounding slack ball_eps(dim), so a page holding a true neighbour is never skipped; which pages
// are scanned never changes a result bit (DESIGN.md "Result contract").
//
// One CTA per 128-query tile (thread = query).  Level 1: every super-group ball against every query; level 2: the pages
// of the super-groups that survive for at least one query, tested by the queries that need the super-group.  Output:
// ascending page ranges per tile in the layout PageWalk reads ([kMaxCuts][kJobRanges] flat = kRangeCap ranges); overflow
// widens the last range (scans more, stays exact).
//
// Tightening (SelectParams::tighten): before the bound test every query takes the minimum of its seed bound and
//     min over super-groups with >= K points of  |q - c| + r
// -- all points of such a ball lie within |q - c| + r of the query, so its K-th neighbour cannot be farther.  A query
// whose approximate leaf lies in the "wrong" cluster has a seed bound several times too large, and one such query would
// make its whole tile list a large part of the image; with this bound every query's test is at least as tight as "the
// nearest cluster-sized ball".  The tightened bound (squared) is also written out for the scan kernel's filter.
//
// First pass (SelectParams::first_pass): a query whose seed bound is worse than that ball bound ("loose": its approximate
// leaf was the wrong place to look) names the two pages nearest to it inside its nearest super-group; the tile scans
// those few pages first (k_scan_tc, lists -> state), after which the loose queries hold real neighbours and the bound test
// of the main selection is tight for every query.  Tiles without loose queries list nothing.
#pragma once
#include "scht_kernels.cuh"

namespace scht {

struct SelectParams {
    DevTree T;
    ScanImage I;
    const float* q;
    const int* q_order;
    int nq, K;
    const uint64_t* state;    // [nq][K] keys by bucket slot (after the seed scan)
    const float* ext_bound;   // [nq] by query id: external upper bound of the K-th distance^2 (or null)
    int2* ranges;             // [n_tiles][kRangeCap]
    int* n_ranges;            // [n_tiles][kMaxCuts]
    int sel_mode, sel_step;   // tile subset (tile_of_selection)
    int tighten;              // 1: tighten every query's bound with the super-group balls first (see above)
    int first_pass;           // 1: list only the pages the loose queries should scan first (needs tighten); 2: the same for every query
    float* bound_out;         // [nq] by query id, or null: the bound^2 each query's test used (an upper bound of its K-th distance^2)
    unsigned long long* counters;  // [4] ball tests (per query), [5] pages listed x queries, [7] (query, page) pairs that pass the bound test
};

constexpr int kSelQ = 128;    // queries per tile (= kTcQ)
constexpr int kSelGroups = 32;  // super-group balls staged per round (one bit each in a 32-bit mask)
constexpr int kSelPages = 8;    // page balls staged per round
static_assert(kSuperPages <= 64, "k_select_pages marks the pages of one super-group in two 32-bit words");

__host__ __device__ inline size_t select_smem_bytes(int dp_regs, int dpad, int n_supers) {
    const int ds = dp_regs > 0 ? dp_regs : dpad;
    size_t b = (size_t)kSelGroups * ds * 4 + 2 * kSelGroups * 4;
    if (dp_regs == 0) b += (size_t)dpad * kSelQ * 4;
    b += (size_t)((n_supers + 31) / 32) * 4;  // marks of the first pass
    return b + 16;
}

// DP > 0: dpad <= DP and the query lives in registers as DP/2 packed pairs; DP == 0: any dpad, query in shared memory
template <int DP>
__global__ void __launch_bounds__(kSelQ) k_select_pages(const SelectParams P) {
    extern __shared__ __align__(16) unsigned char sel_smem[];
    __shared__ unsigned int s_mask;
    __shared__ unsigned int s_pm[2];  // first pass: page marks of one super-group (kSuperPages <= 64)
    __shared__ int s_nout;
    const DevTree& T = P.T;
    const ScanImage& I = P.I;
    const int dpad = T.dpad;
    const int DS = DP > 0 ? DP : dpad;                    // staged row length (floats), a multiple of 4
    float* cs = reinterpret_cast<float*>(sel_smem);       // [kSelGroups][DS] staged centres
    float* rs = cs + (size_t)kSelGroups * DS;             // [kSelGroups] staged radii
    int* ns = reinterpret_cast<int*>(rs + kSelGroups);    // [kSelGroups] staged point counts (super-groups)
    float* qs = reinterpret_cast<float*>(ns + kSelGroups);  // DP == 0: [dpad][kSelQ]
    unsigned int* near_map = reinterpret_cast<unsigned int*>(qs + (DP > 0 ? 0 : (size_t)dpad * kSelQ));  // [(n_supers+31)/32] bits: candidates of the test loop / marks of the first pass
    const int t = threadIdx.x, lane = t & 31;
    const int n_tiles = (P.nq + kSelQ - 1) / kSelQ;
    const int tile = tile_of_selection((int)blockIdx.x, P.sel_mode, P.sel_step);
    if (tile >= n_tiles) return;
    const int q_begin = tile * kSelQ;
    const int nq_tile = min(kSelQ, P.nq - q_begin);
    const bool valid = t < nq_tile;
    const float e = ball_eps(T.dim);

    // ---- the query and its bound ----
    const int qid = valid ? P.q_order[q_begin + t] : -1;
    uint64_t qreg[DP > 0 ? DP / 2 : 1];
    if (DP > 0) {
#pragma unroll
        for (int j2 = 0; j2 < (DP > 0 ? DP / 2 : 1); j2++) {
            const float a = (valid && 2 * j2 < T.dim) ? P.q[(size_t)qid * T.dim + 2 * j2] : 0.f;
            const float b = (valid && 2 * j2 + 1 < T.dim) ? P.q[(size_t)qid * T.dim + 2 * j2 + 1] : 0.f;
            qreg[j2] = pack2(a, b);
        }
    } else {
        for (int j = 0; j < dpad; j++) qs[j * kSelQ + t] = (valid && j < T.dim) ? P.q[(size_t)qid * T.dim + j] : 0.f;
    }
    float thr = -1.f;  // padded slot: needs nothing
    if (valid) {
        float kth2 = __uint_as_float((uint32_t)(P.state[(size_t)(q_begin + t) * P.K + P.K - 1] >> 32));
        if (P.ext_bound) kth2 = fminf(kth2, P.ext_bound[qid]);
        thr = sqrtf(kth2) * (1.f + e) * (1.f + e);  // FLT_MAX (list not full) -> 1.8e19: nothing is skipped
    }
    if (t == 0) s_mask = 0u;
    __syncthreads();

    // squared distance of this thread's query to a staged centre row
    auto dist2 = [&](const float* c) -> float {
        if (DP > 0) {
            uint64_t a0 = pack2(0.f, 0.f), a1 = a0;
#pragma unroll
            for (int j4 = 0; j4 < (DP > 0 ? DP / 4 : 1); j4++) {
                const ulonglong2 c2 = *reinterpret_cast<const ulonglong2*>(c + 4 * j4);
                const uint64_t d0 = sub2(qreg[2 * j4], c2.x), d1 = sub2(qreg[2 * j4 + 1], c2.y);
                a0 = fma2(d0, d0, a0);
                a1 = fma2(d1, d1, a1);
            }
            float x0, x1, y0, y1;
            unpack2(a0, x0, x1);
            unpack2(a1, y0, y1);
            return (x0 + x1) + (y0 + y1);
        } else {
            float a0 = 0.f, a1 = 0.f;
            for (int j = 0; j < dpad; j += 2) {
                const float d0 = qs[j * kSelQ + t] - c[j], d1 = qs[(j + 1) * kSelQ + t] - c[j + 1];
                a0 = fmaf(d0, d0, a0);
                a1 = fmaf(d1, d1, a1);
            }
            return a0 + a1;
        }
    };
    // true when the ball (centre row c, radius r) may hold a point within this query's bound
    auto needs = [&](const float* c, float r) -> bool {
        const float dc = sqrtf(dist2(c));
        return !((dc * (1.f - e) - r * (1.f + e)) > thr);  // NaN / inf never skip
    };
    // stage n balls (rows of `cen`, radii `rad`) starting at index i0
    auto stage = [&](const float* cen, const float* rad, const int* cnt, int i0, int n) {
        for (int x = t; x < n * DS; x += kSelQ) {
            const int r = x / DS, j = x - r * DS;
            cs[x] = (j < dpad) ? cen[(size_t)(i0 + r) * dpad + j] : 0.f;
        }
        if (t < n) {
            rs[t] = rad[i0 + t];
            ns[t] = cnt[i0 + t];
        }
    };

    int2* out = P.ranges + (size_t)tile * kRangeCap;
    int n_out = 0, cur_b = -1, cur_e = -1;  // thread 0 only
    unsigned long long listed = 0, tests = 0, need_pairs = 0;
    auto flush = [&]() {
        if (n_out == kRangeCap) {
            listed += (unsigned long long)(cur_e - out[n_out - 1].y);
            out[n_out - 1].y = cur_e;
        } else {
            listed += (unsigned long long)(cur_e - cur_b);
            out[n_out++] = make_int2(cur_b, cur_e);
        }
    };
    auto add_pages = [&](int pb, int pe) {  // page runs arrive in ascending order (a boundary page may repeat)
        if (cur_b < 0) {
            cur_b = pb;
            cur_e = pe;
        } else if (pb <= cur_e) {
            cur_e = max(cur_e, pe);
        } else {
            flush();
            cur_b = pb;
            cur_e = pe;
        }
    };

    bool loose = false;  // the ball bound beats this query's seed bound
    int near_s = -1;     // super-group with the nearest centre
    // cand_map: bit per super-group, set when some query of the tile could not rule it out with the bound it had at the time (its
    // seed bound, tightened along the way): the test loop below only revisits those.  Without tightening: all set.
    const int n_words = (I.n_supers + 31) / 32;
    for (int w = t; w < n_words; w += kSelQ) near_map[w] = P.tighten ? 0u : 0xffffffffu;
    if (P.tighten) {
        float ub = 3.4e38f, bd = 3.4e38f, run_thr = thr;
        for (int g0 = 0; g0 < I.n_supers; g0 += kSelGroups) {
            const int ng = min(kSelGroups, I.n_supers - g0);
            __syncthreads();
            stage(I.scen, I.srad, I.scnt, g0, ng);
            __syncthreads();
            unsigned int cm = 0;
            if (valid)
                for (int gi = 0; gi < ng; gi++) {
                    const float d2c = dist2(cs + (size_t)gi * DS);
                    const float dc = sqrtf(d2c);
                    if (ns[gi] >= P.K) {
                        ub = fminf(ub, dc * (1.f + e) + rs[gi]);
                        run_thr = fminf(run_thr, ub * (1.f + e) * (1.f + e) * (1.f + e));
                    }
                    if (d2c < bd) bd = d2c, near_s = g0 + gi;
                    if (!((dc * (1.f - e) - rs[gi] * (1.f + e)) > run_thr)) cm |= 1u << gi;
                }
            cm = __reduce_or_sync(0xffffffffu, cm);
            if (lane == 0 && cm) atomicOr(&near_map[g0 >> 5], cm);
        }
        tests += valid ? (unsigned long long)I.n_supers : 0ull;
        const float ub_thr = ub * (1.f + e) * (1.f + e) * (1.f + e);
        loose = valid && (ub_thr < thr || P.first_pass == 2);
        if (valid) thr = fminf(thr, ub_thr);
    }
    __syncthreads();
    if (P.bound_out && valid && !P.first_pass) P.bound_out[qid] = fminf(thr * thr, 3.402823466e+38f);
    if (P.first_pass) {
        // ---- the pages the loose queries scan first: of the nearest super-group of each, its two nearest pages ----
        __syncthreads();
        for (int w = t; w < n_words; w += kSelQ) near_map[w] = 0u;
        __syncthreads();
        if (loose && near_s >= 0) atomicOr(&near_map[near_s >> 5], 1u << (near_s & 31));
        __syncthreads();
        for (int w = 0; w < n_words; w++) {
            unsigned int m = near_map[w];  // block-uniform
            while (m) {
                const int sgi = 32 * w + __ffs(m) - 1;
                m &= m - 1;
                const int pb = I.spage[sgi], pe = I.spage[sgi + 1];
                const bool mine = loose && near_s == sgi;
                float d1 = 3.4e38f, d2 = 3.4e38f;
                int p1 = -1, p2 = -1;
                for (int p0 = pb; p0 < pe; p0 += kSelPages) {
                    const int np = min(kSelPages, pe - p0);
                    __syncthreads();
                    stage(I.pcen, I.prad, I.pcnt, p0, np);
                    if (t < 2) s_pm[t] = 0u;
                    __syncthreads();
                    if (mine)
                        for (int pi = 0; pi < np; pi++) {
                            const float d = dist2(cs + (size_t)pi * DS);
                            if (d < d1) {
                                d2 = d1, p2 = p1;
                                d1 = d, p1 = p0 + pi;
                            } else if (d < d2) {
                                d2 = d, p2 = p0 + pi;
                            }
                        }
                }
                if (p1 >= 0) atomicOr(&s_pm[(p1 - pb) >> 5], 1u << ((p1 - pb) & 31));
                if (p2 >= 0) atomicOr(&s_pm[(p2 - pb) >> 5], 1u << ((p2 - pb) & 31));
                __syncthreads();
                if (t == 0)
                    for (int k = 0; k < 2; k++) {
                        unsigned int pm = s_pm[k];
                        while (pm) {
                            const int pg = pb + 32 * k + __ffs(pm) - 1;
                            pm &= pm - 1;
                            add_pages(pg, pg + 1);
                        }
                    }
            }
        }
    } else
    for (int g0 = 0; g0 < I.n_supers; g0 += kSelGroups) {
        const int ng = min(kSelGroups, I.n_supers - g0);
        const unsigned int cand_w = near_map[g0 >> 5];  // block-uniform (written before the barrier above)
        if (cand_w == 0u) continue;
        __syncthreads();  // previous round's readers of cs are done
        stage(I.scen, I.srad, I.scnt, g0, ng);
        __syncthreads();
        unsigned int gm = 0;
        if (valid)
            for (int gi = 0; gi < ng; gi++)
                if (((cand_w >> gi) & 1u) && needs(cs + (size_t)gi * DS, rs[gi])) gm |= 1u << gi;
        tests += valid ? (unsigned long long)__popc(cand_w) : 0ull;
        const unsigned int wm = __reduce_or_sync(0xffffffffu, gm);
        if (lane == 0 && wm) atomicOr(&s_mask, wm);
        __syncthreads();
        unsigned int tile_mask = s_mask;
        __syncthreads();
        if (t == 0) s_mask = 0u;
        while (tile_mask) {  // block-uniform: the super-groups some query of the tile needs
            const int gi = __ffs(tile_mask) - 1;
            tile_mask &= tile_mask - 1;
            const int g = g0 + gi;
            const int pb = I.spage[g], pe = I.spage[g + 1];
            const bool mine = (gm >> gi) & 1u;
            for (int p0 = pb; p0 < pe; p0 += kSelPages) {
                const int np = min(kSelPages, pe - p0);
                __syncthreads();  // cs free; s_mask reset visible
                stage(I.pcen, I.prad, I.pcnt, p0, np);
                __syncthreads();
                unsigned int pm = 0;
                if (mine)
                    for (int pi = 0; pi < np; pi++)
                        if (needs(cs + (size_t)pi * DS, rs[pi])) pm |= 1u << pi;
                if (mine) tests += (unsigned long long)np;
                need_pairs += (unsigned long long)__popc(pm);
                const unsigned int wpm = __reduce_or_sync(0xffffffffu, pm);
                if (lane == 0 && wpm) atomicOr(&s_mask, wpm);
                __syncthreads();
                if (t == 0) {
                    unsigned int m = s_mask;
                    s_mask = 0u;
                    while (m) {
                        const int pi = __ffs(m) - 1;
                        m &= m - 1;
                        add_pages(p0 + pi, p0 + pi + 1);
                    }
                }
            }
        }
    }
    if (t == 0) {
        if (cur_b >= 0) flush();
        s_nout = n_out;
    }
    __syncthreads();
    const int total = s_nout;
    if (t < kMaxCuts) P.n_ranges[(size_t)tile * kMaxCuts + t] = max(0, min(kJobRanges, total - t * kJobRanges));
    if (P.counters) {
        for (int o = 16; o; o >>= 1) {
            tests += __shfl_xor_sync(0xffffffffu, tests, o);
            need_pairs += __shfl_xor_sync(0xffffffffu, need_pairs, o);
        }
        if (lane == 0) {
            atomicAdd(&P.counters[4], tests);
            if (need_pairs) atomicAdd(&P.counters[7], need_pairs);
        }
        if (t == 0) atomicAdd(&P.counters[5], listed * (unsigned long long)nq_tile);
    }
}

}  // namespace scht
