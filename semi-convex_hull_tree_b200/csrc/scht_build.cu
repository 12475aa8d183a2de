// Device tree builder: the constructor path of Convex_Tree<T>
//     build_tree (reference Code/Convex_Tree.h:708-857) -> build_simplified_tree (:685-705) -> sort_raw_data (:593-633)
// executed on one B200, with the reference's float/double arithmetic and its rand() stream, so that the flat arrays are
// BIT-IDENTICAL to the host builder's (tree_builder.cpp) and hence to the reference's own tree.
//
// Why one persistent cooperative kernel.  The reference draws one rand() per internal node in DFS pre-order
// (Code/Convex_Tree.h:789), so the r-th draw belongs to the r-th internal node of the pre-order: the start of a right
// subtree depends on how many internal nodes its left sibling produced.  Bit-identity therefore forces the nodes to be
// created in DFS order; what is parallel is the work INSIDE a node (three distance passes, the re-tightening of every
// inherited hyper-plane, the stable partition).  k_build_tree walks the DFS with every CTA of a cooperative grid
// executing the same control flow (node ids, rand index, CSR offsets are recomputed redundantly and never communicated);
// data-dependent values (arg-max rows, partition sizes) go through global memory around grid.sync().  Per internal node:
// 4 grid syncs, per leaf: 1.  The rows of a node are contiguous (the partition is stable and ping-pongs between two row
// buffers), so the final buffer IS m_sorted_data and a node's range is its leaf range.
//
// Arithmetic that must match the host bit for bit (all of it is order-independent per row, only reductions are parallel):
//   * distances: acc = (float)((double)acc + (double)d*(double)d), d = fl32(p_j - c_j); sqrtf          (basic_functions.h:129-143)
//   * getBeta_1: tmp = fl32(tmp + fl32(p_j*a_j)) (no FMA contraction), v = fl32(tmp + beta0), FIRST minimum     (Convex_Tree.h:563-589)
//   * index_max / index_min: first extremum -> 64-bit keys (value bits, row index) reduced with atomicMax/atomicMin
//   * scalar_product: float products summed in a double                                                       (basic_functions.h:20-29)
#include <cooperative_groups.h>
#include <cuda_runtime.h>

#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/scht.h"
#include "scht_error.h"
#include "scht_host_tree.h"

namespace cg = cooperative_groups;

namespace {

constexpr int kBThreads = 1024;
constexpr int kBMaxDepth = 64;  // most constraints (= depth) a node may have on the device path
constexpr int kBCols = 8;       // hyper-planes evaluated per pass over a row
constexpr int kSlotStride = kBMaxDepth + 2;
constexpr int kSlotArgmax = kBMaxDepth;

enum BuildStatus { kBuildOk = 0, kBuildNodeOverflow = 1, kBuildBetaOverflow = 2, kBuildTooDeep = 3, kBuildRandOverflow = 4 };

struct BuildParams {
    float* rows[2];    // [N][D] row-major, ping-pong
    int32_t* orig[2];  // [N] original index of each row
    float* dl;         // [N] distance to the left pivot
    uint8_t* flag;     // [N] 1 = goes left
    const int32_t* rand;
    int rand_cap;
    int n;
    int dim;
    int threshold;
    int32_t *left, *right, *parent, *node_a, *node_b, *beta_off;
    float *lc, *rc, *normal, *beta;
    int node_cap;
    long long beta_cap;
    unsigned long long* slots;  // [3][kSlotStride]
    int* cta_cnt;               // [3][gridDim]
    int* result;                // status, n_nodes, n_internal, beta_size(lo), depth_stat, degenerate, beta_size(hi)
};

struct Frame {
    int a, b, parent, k, parity, is_right, pbo_lo, pbo_hi;
};

__device__ __forceinline__ float b_acc_sq(float acc, float d) { return (float)((double)acc + (double)d * (double)d); }

// monotone map float -> uint32 (-0 must be canonicalised by the caller)
__device__ __forceinline__ uint32_t ord_f32(float v) {
    uint32_t b = __float_as_uint(v);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

template <bool IS_MIN>
__device__ __forceinline__ unsigned long long warp_red(unsigned long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        unsigned long long w = __shfl_xor_sync(0xffffffffu, v, o);
        v = IS_MIN ? (w < v ? w : v) : (w > v ? w : v);
    }
    return v;
}

// Block-wide reduction of NV keys per thread; thread 0 applies them to global slots with atomics.
template <bool IS_MIN, int NV>
__device__ __forceinline__ void block_red_to_slots(unsigned long long (&v)[NV], int nvalid, unsigned long long* red /*[32][NV]*/,
                                                   unsigned long long* slots) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int u = 0; u < NV; u++) {
        unsigned long long r = warp_red<IS_MIN>(v[u]);
        if (lane == 0) red[warp * NV + u] = r;
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int u = 0; u < NV; u++) {
            unsigned long long r = warp_red<IS_MIN>(red[lane * NV + u]);
            if (lane == 0 && u < nvalid) {
                if (IS_MIN) atomicMin(&slots[u], r);
                else atomicMax(&slots[u], r);
            }
        }
    }
    __syncthreads();
}

// distance from `center` (shared memory) to row p, the reference's way
__device__ __forceinline__ float row_dist(const float* __restrict__ p, const float* center, int dim) {
    float acc = 0.f;
    for (int j = 0; j < dim; j++) acc = b_acc_sq(acc, __fsub_rn(p[j], center[j]));
    return __fsqrt_rn(acc);
}

__device__ __forceinline__ void copy_row(float* __restrict__ dst, const float* __restrict__ src, int dim) {
    if ((dim & 3) == 0) {
        const float4* s4 = reinterpret_cast<const float4*>(src);
        float4* d4 = reinterpret_cast<float4*>(dst);
        for (int j = 0; j < (dim >> 2); j++) d4[j] = s4[j];
    } else {
        for (int j = 0; j < dim; j++) dst[j] = src[j];
    }
}

__global__ void __launch_bounds__(kBThreads, 1) k_build_tree(BuildParams P) {
    cg::grid_group grid = cg::this_grid();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int D = P.dim;
    float* nrmT = reinterpret_cast<float*>(smem_raw);  // [D][kBMaxDepth]: normal of path edge c, transposed
    float* beta0 = nrmT + (size_t)D * kBMaxDepth;      // [kBMaxDepth]
    float* center = beta0 + kBMaxDepth;                // [D]
    float* lc_s = center + D;                          // [D]
    float* rc_s = lc_s + D;                            // [D]
    float* diff_s = rc_s + D;                          // [D]
    float* mean_s = diff_s + D;                        // [D]
    float* own_s = mean_s + D;                         // [D]
    __shared__ unsigned long long red[32 * kBCols];
    __shared__ int wsum[33];
    __shared__ Frame stack[kBMaxDepth + 4];
    __shared__ float nrm_bcast;
    __shared__ int nl_bcast;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int G = gridDim.x, cta = blockIdx.x;
    const bool lead = (cta == 0);

    int sp = 0;
    if (tid == 0) stack[0] = Frame{0, P.n, -1, 0, 0, 0, 0, 0};
    sp = 1;
    __syncthreads();

    int next_id = 0, rand_i = 0, depth_stat = 1, degenerate = 0, status = kBuildOk;
    long long beta_size = 0;
    unsigned phase = 0;

    while (sp > 0) {
        const Frame f = stack[sp - 1];
        sp--;
        __syncthreads();  // everyone holds f before the slot can be overwritten
        const int id = next_id;
        const int k = f.k, a = f.a, rows = f.b - f.a;
        if (id >= P.node_cap) { status = kBuildNodeOverflow; break; }
        if (k > kBMaxDepth) { status = kBuildTooDeep; break; }
        if (beta_size + k > P.beta_cap) { status = kBuildBetaOverflow; break; }
        next_id++;
        const bool leaf = rows <= P.threshold;  // Convex_Tree.h:776
        if (!leaf && rand_i >= P.rand_cap) { status = kBuildRandOverflow; break; }
        const float* __restrict__ src = P.rows[f.parity];
        const int32_t* __restrict__ src_orig = P.orig[f.parity];

        // ---- child set-up: own hyper-plane from the parent's pivots (Convex_Tree.h:826-834,847-849) ----------------
        if (f.parent >= 0) {
            if (tid < D) {
                float l = P.lc[(size_t)f.parent * D + tid], r = P.rc[(size_t)f.parent * D + tid];
                diff_s[tid] = __fsub_rn(l, r);
                mean_s[tid] = __fdiv_rn(__fadd_rn(l, r), 2.f);
            }
            __syncthreads();
            if (tid == 0) {
                float nacc = 0.f;
                for (int j = 0; j < D; j++) nacc = b_acc_sq(nacc, diff_s[j]);
                nrm_bcast = __fsqrt_rn(nacc);
            }
            __syncthreads();
            if (tid < D) {
                float la = __fdiv_rn(diff_s[tid], nrm_bcast);
                float v = f.is_right ? __fmul_rn(-1.f, la) : la;
                nrmT[tid * kBMaxDepth + (k - 1)] = v;
                own_s[tid] = v;
                if (lead) P.normal[(size_t)id * D + tid] = v;
            }
            const long long pbo = ((long long)f.pbo_hi << 32) | (unsigned)f.pbo_lo;
            if (tid < k - 1) beta0[tid] = P.beta[pbo + tid];
            __syncthreads();
            if (tid == 0) {
                double s = 0.0;
                for (int j = 0; j < D; j++) s += (double)__fmul_rn(own_s[j], mean_s[j]);
                beta0[k - 1] = -(float)s;
            }
            __syncthreads();
        } else if (lead && tid < D) {
            P.normal[(size_t)id * D + tid] = 0.f;
        }
        if (lead && tid == 0) {
            P.parent[id] = f.parent;
            P.left[id] = -1;
            P.right[id] = -1;
            P.node_a[id] = f.a;
            P.node_b[id] = f.b;
            P.beta_off[id] = (int32_t)beta_size;
            if (f.parent >= 0) {
                if (f.is_right) P.right[f.parent] = id;
                else P.left[f.parent] = id;
            }
        }
        if (lead && tid < D) {
            P.lc[(size_t)id * D + tid] = 0.f;
            P.rc[(size_t)id * D + tid] = 0.f;
        }

        // this CTA's contiguous share of the node's rows
        int chunk = (rows + G - 1) / G;
        if (chunk < kBThreads) chunk = kBThreads;
        const long long lo_ll = (long long)a + (long long)cta * chunk;
        const int lo = lo_ll < f.b ? (int)lo_ll : f.b;
        const int hi = (lo_ll + chunk < f.b) ? (int)(lo_ll + chunk) : f.b;

        // ================= phase A: re-tighten the k hyper-planes; distances from the random pick ===================
        unsigned long long* slotA = P.slots + (phase % 3) * kSlotStride;
        if (lead && tid < kSlotStride) {
            unsigned long long* nxt = P.slots + ((phase + 1) % 3) * kSlotStride;
            nxt[tid] = (tid == kSlotArgmax) ? 0ull : ~0ull;
        }
        if (lo < hi) {
            for (int c0 = 0; c0 < k; c0 += kBCols) {
                unsigned long long best[kBCols];
#pragma unroll
                for (int u = 0; u < kBCols; u++) best[u] = ~0ull;
                for (int i = lo + tid; i < hi; i += kBThreads) {
                    const float* __restrict__ p = src + (size_t)i * D;
                    float acc[kBCols];
#pragma unroll
                    for (int u = 0; u < kBCols; u++) acc[u] = 0.f;
                    for (int j = 0; j < D; j++) {
                        const float pj = p[j];
                        const float* nr = nrmT + j * kBMaxDepth + c0;
#pragma unroll
                        for (int u = 0; u < kBCols; u++) acc[u] = __fadd_rn(acc[u], __fmul_rn(pj, nr[u]));
                    }
#pragma unroll
                    for (int u = 0; u < kBCols; u++) {
                        if (c0 + u < k) {
                            float v = __fadd_rn(__fadd_rn(acc[u], beta0[c0 + u]), 0.f);  // + 0: -0 -> +0 (compares equal on the host)
                            unsigned long long key = ((unsigned long long)ord_f32(v) << 32) | (unsigned)(i - a);
                            best[u] = key < best[u] ? key : best[u];
                        }
                    }
                }
                int nvalid = k - c0 < kBCols ? k - c0 : kBCols;
                block_red_to_slots<true, kBCols>(best, nvalid, red, slotA + c0);
            }
            if (!leaf) {
                const int pick = P.rand[rand_i] % rows;  // Convex_Tree.h:789
                if (tid < D) center[tid] = src[(size_t)(a + pick) * D + tid];
                __syncthreads();
                unsigned long long best[1] = {0ull};
                for (int i = lo + tid; i < hi; i += kBThreads) {
                    float d = row_dist(src + (size_t)i * D, center, D);
                    unsigned long long key = ((unsigned long long)__float_as_uint(d) << 32) | (0xffffffffu - (unsigned)(i - a));
                    best[0] = key > best[0] ? key : best[0];
                }
                block_red_to_slots<false, 1>(best, 1, red, slotA + kSlotArgmax);
            } else if (f.parity == 1) {
                // a leaf that ended in the odd buffer: bring it home (the even buffer is m_sorted_data)
                for (int i = lo + tid; i < hi; i += kBThreads) {
                    copy_row(P.rows[0] + (size_t)i * D, src + (size_t)i * D, D);
                    P.orig[0][i] = src_orig[i];
                }
            }
        }
        if (!leaf) rand_i++;
        grid.sync();
        phase++;

        // this node's offsets: -scalar_product(normal, extreme row) (Convex_Tree.h:586-588)
        if (lead && tid < k) {
            const unsigned best = (unsigned)(slotA[tid] & 0xffffffffu);
            const float* __restrict__ p = src + (size_t)(a + (int)best) * D;
            double s = 0.0;
            for (int j = 0; j < D; j++) s += (double)__fmul_rn(nrmT[j * kBMaxDepth + tid], p[j]);
            P.beta[beta_size + tid] = -(float)s;
        }
        const long long my_bo = beta_size;
        beta_size += k;
        if (leaf) continue;

        // ================= phase B: left pivot = farthest from the pick; distances to it ============================
        {
            const int li = (int)(0xffffffffu - (unsigned)(slotA[kSlotArgmax] & 0xffffffffu));
            unsigned long long* slotB = P.slots + (phase % 3) * kSlotStride;
            if (lead && tid < kSlotStride) {
                unsigned long long* nxt = P.slots + ((phase + 1) % 3) * kSlotStride;
                nxt[tid] = (tid == kSlotArgmax) ? 0ull : ~0ull;
            }
            if (tid < D) lc_s[tid] = src[(size_t)(a + li) * D + tid];
            __syncthreads();
            if (lo < hi) {
                unsigned long long best[1] = {0ull};
                for (int i = lo + tid; i < hi; i += kBThreads) {
                    float d = row_dist(src + (size_t)i * D, lc_s, D);
                    P.dl[i] = d;
                    unsigned long long key = ((unsigned long long)__float_as_uint(d) << 32) | (0xffffffffu - (unsigned)(i - a));
                    best[0] = key > best[0] ? key : best[0];
                }
                block_red_to_slots<false, 1>(best, 1, red, slotB + kSlotArgmax);
            }
            grid.sync();
            phase++;

            // ============= phase C: right pivot = farthest from the left one; side of every row ======================
            const int ri = (int)(0xffffffffu - (unsigned)(slotB[kSlotArgmax] & 0xffffffffu));
            int* cntC = P.cta_cnt + (phase % 3) * G;
            if (lead && tid < kSlotStride) {
                unsigned long long* nxt = P.slots + ((phase + 1) % 3) * kSlotStride;
                nxt[tid] = (tid == kSlotArgmax) ? 0ull : ~0ull;
            }
            if (tid < D) rc_s[tid] = src[(size_t)(a + ri) * D + tid];
            __syncthreads();
            int mine = 0;
            for (int i = lo + tid; i < hi; i += kBThreads) {
                float d = row_dist(src + (size_t)i * D, rc_s, D);
                float diff = __fsub_rn(P.dl[i], d);  // Convex_Tree.h:809-821
                int fl = diff <= 0.f;
                P.flag[i] = (uint8_t)fl;
                mine += fl;
            }
            for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
            if (lane == 0) wsum[warp] = mine;
            __syncthreads();
            if (warp == 0) {
                int v = wsum[lane];
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                if (lane == 0) cntC[cta] = v;
            }
            grid.sync();
            phase++;

            // ============= phase D: stable partition into the other buffer ==========================================
            if (lead && tid < kSlotStride) {
                unsigned long long* nxt = P.slots + ((phase + 1) % 3) * kSlotStride;
                nxt[tid] = (tid == kSlotArgmax) ? 0ull : ~0ull;
            }
            // left rows of the CTAs before this one, and of the whole node
            {
                int before = 0, total = 0;
                for (int c = tid; c < G; c += kBThreads) {
                    int v = cntC[c];
                    total += v;
                    if (c < cta) before += v;
                }
                for (int o = 16; o > 0; o >>= 1) {
                    before += __shfl_xor_sync(0xffffffffu, before, o);
                    total += __shfl_xor_sync(0xffffffffu, total, o);
                }
                __syncthreads();
                if (lane == 0) { wsum[warp] = before; red[warp] = (unsigned long long)total; }
                __syncthreads();
                if (warp == 0) {
                    int bsum = wsum[lane], tsum = (int)red[lane];
                    for (int o = 16; o > 0; o >>= 1) {
                        bsum += __shfl_xor_sync(0xffffffffu, bsum, o);
                        tsum += __shfl_xor_sync(0xffffffffu, tsum, o);
                    }
                    if (lane == 0) { wsum[32] = bsum; nl_bcast = tsum; }
                }
                __syncthreads();
            }
            const int nl = nl_bcast;
            const int left_before = wsum[32];
            __syncthreads();
            if (nl == 0 || nl == rows) {
                // unsplittable node (more than `threshold` identical rows): closed as an oversized leaf, like the host builder
                degenerate = 1;
                if (f.parity == 1) {
                    for (int i = lo + tid; i < hi; i += kBThreads) {
                        copy_row(P.rows[0] + (size_t)i * D, src + (size_t)i * D, D);
                        P.orig[0][i] = src_orig[i];
                    }
                }
                grid.sync();
                phase++;
                continue;
            }
            if (lead && tid < D) {
                P.lc[(size_t)id * D + tid] = lc_s[tid];
                P.rc[(size_t)id * D + tid] = rc_s[tid];
            }
            {
                float* __restrict__ dst = P.rows[f.parity ^ 1];
                int32_t* __restrict__ dst_orig = P.orig[f.parity ^ 1];
                int lpos = a + left_before;                   // next left destination
                int rpos = a + nl + (lo - a - left_before);  // next right destination
                for (int t = lo; t < hi; t += kBThreads) {
                    const int i = t + tid;
                    const bool valid = i < hi;
                    const int fl = valid ? (int)P.flag[i] : 0;
                    const unsigned bl = __ballot_sync(0xffffffffu, fl);
                    const int wl = __popc(bl);
                    const int rank_l = __popc(bl & ((1u << lane) - 1u));
                    if (lane == 0) wsum[warp] = wl;
                    __syncthreads();
                    int wpre = 0, tile_l = 0;
                    {
                        int v = wsum[lane];
                        int incl = v;
                        for (int o = 1; o < 32; o <<= 1) {
                            int n = __shfl_up_sync(0xffffffffu, incl, o);
                            if (lane >= o) incl += n;
                        }
                        wpre = __shfl_sync(0xffffffffu, incl - v, warp);
                        tile_l = __shfl_sync(0xffffffffu, incl, 31);
                    }
                    if (valid) {
                        const int lrank = wpre + rank_l;
                        const int idx_in_tile = tid;
                        const int d = fl ? (lpos + lrank) : (rpos + (idx_in_tile - lrank));
                        copy_row(dst + (size_t)d * D, src + (size_t)i * D, D);
                        dst_orig[d] = src_orig[i];
                    }
                    const int tile_n = (hi - t) < kBThreads ? (hi - t) : kBThreads;
                    lpos += tile_l;
                    rpos += tile_n - tile_l;
                    __syncthreads();
                }
            }
            grid.sync();
            phase++;

            // children: right pushed first so that the left subtree is built first (DFS pre-order)
            if (tid == 0) {
                stack[sp] = Frame{a + nl, f.b, id, k + 1, f.parity ^ 1, 1, (int)(my_bo & 0xffffffffll), (int)(my_bo >> 32)};
                stack[sp + 1] = Frame{a, a + nl, id, k + 1, f.parity ^ 1, 0, (int)(my_bo & 0xffffffffll), (int)(my_bo >> 32)};
            }
            sp += 2;
            if (k + 2 > depth_stat) depth_stat++;  // `depth` statistic, Convex_Tree.h:837-839 (root is built with cur_depth 1)
            __syncthreads();
        }
    }
    if (lead && tid == 0) {
        P.result[0] = status;
        P.result[1] = next_id;
        P.result[2] = rand_i;
        P.result[3] = (int)(beta_size & 0xffffffffll);
        P.result[4] = depth_stat;
        P.result[5] = degenerate;
        P.result[6] = (int)(beta_size >> 32);
    }
}

__global__ void k_iota(int32_t* p, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = i;
}

// glibc rand() stream (TYPE_3), same generator as tree_builder.cpp
void glibc_rand_stream(uint32_t seed, std::vector<int32_t>& out, size_t count) {
    if (seed == 0) seed = 1;
    std::vector<uint32_t> r(344 + count);
    int32_t s[34];
    s[0] = (int32_t)seed;
    for (int i = 1; i < 31; i++) {
        int64_t hi = s[i - 1] / 127773, lo = s[i - 1] % 127773;
        int64_t word = 16807 * lo - 2836 * hi;
        if (word < 0) word += 2147483647;
        s[i] = (int32_t)word;
    }
    for (int i = 31; i < 34; i++) s[i] = s[i - 31];
    for (int i = 0; i < 34; i++) r[i] = (uint32_t)s[i];
    for (size_t i = 34; i < 344 + count; i++) r[i] = r[i - 31] + r[i - 3];
    out.resize(count);
    for (size_t i = 0; i < count; i++) out[i] = (int32_t)(r[344 + i] >> 1);
}

struct DevMem {
    std::vector<void*> ptrs;
    ~DevMem() {
        for (void* p : ptrs) cudaFree(p);
    }
    template <class T>
    cudaError_t alloc(T** out, size_t count) {
        void* p = nullptr;
        cudaError_t e = cudaMalloc(&p, (count ? count : 1) * sizeof(T));
        if (e == cudaSuccess) ptrs.push_back(p);
        *out = reinterpret_cast<T*>(p);
        return e;
    }
};

#define CU_B(expr)                                                                                            \
    do {                                                                                                      \
        cudaError_t e__ = (expr);                                                                             \
        if (e__ != cudaSuccess)                                                                               \
            return scht::fail(e__ == cudaErrorMemoryAllocation ? SCHT_ERR_OOM : SCHT_ERR_CUDA,               \
                              std::string("scht_build_tree_device: ") + #expr + ": " + cudaGetErrorString(e__)); \
    } while (0)

}  // namespace

extern "C" int scht_build_tree_device(const float* data, int64_t n_points, int32_t dim, float leaf_pts_percent,
                                      uint32_t rand_seed, int32_t device, scht_host_tree** out) {
    if (!out) return scht::fail(SCHT_ERR_INVALID_ARG, "scht_build_tree_device: out is NULL");
    *out = nullptr;
    if (!data || n_points < 1 || n_points > 0x7fffff00LL || dim < 1)
        return scht::fail(SCHT_ERR_INVALID_ARG, "scht_build_tree_device: bad data/n_points/dim");
    if (!(leaf_pts_percent > 0.f) || leaf_pts_percent >= 0.5f)
        return scht::fail(SCHT_ERR_INVALID_ARG, "illegal percent, it should be in (0,0.5)");
    int ndev = scht_device_count();
    if (ndev == 0) return scht::fail(SCHT_ERR_NO_DEVICE, "no CUDA device visible (the device builder has no CPU fallback; scht_build_tree is the host builder)");
    if (device < 0 || device >= ndev) return scht::fail(SCHT_ERR_INVALID_ARG, "scht_build_tree_device: bad device ordinal");
    cudaDeviceProp prop;
    CU_B(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) return scht::fail(SCHT_ERR_NO_DEVICE, "scht_build_tree_device: this library is built for sm_100a only");
    if (!prop.cooperativeLaunch) return scht::fail(SCHT_ERR_NO_DEVICE, "scht_build_tree_device: device cannot launch cooperative kernels");
    CU_B(cudaSetDevice(device));

    const int N = (int)n_points, D = dim;
    const int thr = (int)((float)(unsigned)n_points * leaf_pts_percent);  // Code/Convex_Tree.h:525
    const size_t smem = ((size_t)D * kBMaxDepth + kBMaxDepth + 6 * (size_t)D) * sizeof(float);
    if (smem > (size_t)prop.sharedMemPerBlockOptin)
        return scht::fail(SCHT_ERR_INVALID_ARG, "scht_build_tree_device: dimension too large for the device builder");
    auto t0 = std::chrono::steady_clock::now();

    CU_B(cudaFuncSetAttribute(k_build_tree, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    CU_B(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_build_tree, kBThreads, smem));
    if (per_sm < 1) return scht::fail(SCHT_ERR_INTERNAL, "scht_build_tree_device: kernel does not fit an SM");
    const int G = prop.multiProcessorCount;  // one CTA per SM

    DevMem mem;
    BuildParams P{};
    CU_B(mem.alloc(&P.rows[0], (size_t)N * D));
    CU_B(mem.alloc(&P.rows[1], (size_t)N * D));
    CU_B(mem.alloc(&P.orig[0], (size_t)N));
    CU_B(mem.alloc(&P.orig[1], (size_t)N));
    CU_B(mem.alloc(&P.dl, (size_t)N));
    CU_B(mem.alloc(&P.flag, (size_t)N));
    CU_B(mem.alloc(&P.slots, (size_t)3 * kSlotStride));
    CU_B(mem.alloc(&P.cta_cnt, (size_t)3 * G));
    CU_B(mem.alloc(&P.result, (size_t)8));
    P.n = N;
    P.dim = D;
    P.threshold = thr;

    // node capacity: leaves come out at 1.3-3x N/threshold on real data; grow and retry on overflow (worst case 2N-1 nodes)
    long long node_cap = 16ll * (N / (thr > 0 ? thr : 1) + 1) + 4096;
    const long long node_max = 2ll * N + 1;
    long long beta_per_node = 24;
    int result[8] = {};
    std::vector<int32_t> rnd;
    for (int attempt = 0;; attempt++) {
        if (node_cap > node_max) node_cap = node_max;
        DevMem nodes;
        int32_t* d_rand = nullptr;
        CU_B(nodes.alloc(&P.left, (size_t)node_cap));
        CU_B(nodes.alloc(&P.right, (size_t)node_cap));
        CU_B(nodes.alloc(&P.parent, (size_t)node_cap));
        CU_B(nodes.alloc(&P.node_a, (size_t)node_cap));
        CU_B(nodes.alloc(&P.node_b, (size_t)node_cap));
        CU_B(nodes.alloc(&P.beta_off, (size_t)node_cap));
        CU_B(nodes.alloc(&P.lc, (size_t)node_cap * D));
        CU_B(nodes.alloc(&P.rc, (size_t)node_cap * D));
        CU_B(nodes.alloc(&P.normal, (size_t)node_cap * D));
        P.beta_cap = node_cap * beta_per_node;
        CU_B(nodes.alloc(&P.beta, (size_t)P.beta_cap));
        CU_B(nodes.alloc(&d_rand, (size_t)node_cap));
        glibc_rand_stream(rand_seed, rnd, (size_t)node_cap);
        CU_B(cudaMemcpy(d_rand, rnd.data(), (size_t)node_cap * sizeof(int32_t), cudaMemcpyHostToDevice));
        P.rand = d_rand;
        P.rand_cap = (int)std::min<long long>(node_cap, 0x7fffffff);
        P.node_cap = (int)std::min<long long>(node_cap, 0x7fffffff);

        CU_B(cudaMemcpy(P.rows[0], data, (size_t)N * D * sizeof(float), cudaMemcpyHostToDevice));
        k_iota<<<(N + 255) / 256, 256>>>(P.orig[0], N);
        CU_B(cudaGetLastError());
        {
            std::vector<unsigned long long> init((size_t)3 * kSlotStride, ~0ull);
            for (int s = 0; s < 3; s++) init[(size_t)s * kSlotStride + kSlotArgmax] = 0ull;
            CU_B(cudaMemcpy(P.slots, init.data(), init.size() * sizeof(unsigned long long), cudaMemcpyHostToDevice));
            CU_B(cudaMemset(P.cta_cnt, 0, (size_t)3 * G * sizeof(int)));
            CU_B(cudaMemset(P.result, 0xff, 8 * sizeof(int)));
        }
        void* args[] = {&P};
        CU_B(cudaLaunchCooperativeKernel((void*)k_build_tree, dim3(G), dim3(kBThreads), args, smem, 0));
        CU_B(cudaDeviceSynchronize());
        CU_B(cudaMemcpy(result, P.result, sizeof(result), cudaMemcpyDeviceToHost));
        if (result[0] == kBuildOk) {
            // ---- download and assemble the host tree ----------------------------------------------------------------
            const int nn = result[1];
            const long long nbeta = ((long long)result[6] << 32) | (unsigned)result[3];
            auto* t = new (std::nothrow) scht_host_tree();
            if (!t) return scht::fail(SCHT_ERR_OOM, "scht_build_tree_device: out of host memory");
            try {
                std::vector<int32_t> na(nn), nb(nn);
                t->left.resize(nn); t->right.resize(nn); t->parent.resize(nn); t->leaf_index.assign(nn, -1);
                t->beta_off.resize((size_t)nn + 1);
                t->left_center.resize((size_t)nn * D); t->right_center.resize((size_t)nn * D); t->normal.resize((size_t)nn * D);
                t->beta.resize((size_t)std::max<long long>(nbeta, 1), 0.f);
                t->sorted_points.resize((size_t)N * D); t->sorted_to_orig.resize((size_t)N);
                cudaError_t e = cudaSuccess;
                auto dl = [&](void* dst, const void* src, size_t bytes) {
                    if (e == cudaSuccess && bytes) e = cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost);
                };
                dl(t->left.data(), P.left, sizeof(int32_t) * nn);
                dl(t->right.data(), P.right, sizeof(int32_t) * nn);
                dl(t->parent.data(), P.parent, sizeof(int32_t) * nn);
                dl(na.data(), P.node_a, sizeof(int32_t) * nn);
                dl(nb.data(), P.node_b, sizeof(int32_t) * nn);
                dl(t->beta_off.data(), P.beta_off, sizeof(int32_t) * nn);
                dl(t->left_center.data(), P.lc, sizeof(float) * (size_t)nn * D);
                dl(t->right_center.data(), P.rc, sizeof(float) * (size_t)nn * D);
                dl(t->normal.data(), P.normal, sizeof(float) * (size_t)nn * D);
                dl(t->beta.data(), P.beta, sizeof(float) * (size_t)nbeta);
                dl(t->sorted_points.data(), P.rows[0], sizeof(float) * (size_t)N * D);
                dl(t->sorted_to_orig.data(), P.orig[0], sizeof(int32_t) * (size_t)N);
                if (e != cudaSuccess) {
                    delete t;
                    return scht::fail(SCHT_ERR_CUDA, std::string("scht_build_tree_device: download: ") + cudaGetErrorString(e));
                }
                t->beta_off[nn] = (int32_t)nbeta;
                // build_simplified_tree (Code/Convex_Tree.h:685-705): leaf ids in node order; sort_raw_data: leaf ranges
                int L = 0;
                for (int i = 0; i < nn; i++)
                    if (t->left[i] < 0 && t->right[i] < 0) t->leaf_index[i] = L++;
                t->leaf_start.resize(L); t->leaf_count.resize(L); t->leaf_node_id.resize(L);
                int64_t off = 0;
                for (int i = 0; i < nn; i++) {
                    int li = t->leaf_index[i];
                    if (li < 0) continue;
                    if (na[i] != off) { delete t; return scht::fail(SCHT_ERR_INTERNAL, "scht_build_tree_device: leaves are not contiguous"); }
                    t->leaf_start[li] = na[i]; t->leaf_count[li] = nb[i] - na[i]; t->leaf_node_id[li] = i;
                    off = nb[i];
                }
                if (off != N) { delete t; return scht::fail(SCHT_ERR_INTERNAL, "scht_build_tree_device: tree build lost points (NaN coordinates?)"); }
            } catch (const std::bad_alloc&) {
                delete t;
                return scht::fail(SCHT_ERR_OOM, "scht_build_tree_device: out of host memory");
            }
            t->build_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            t->bind(D, n_points, result[4], thr);
            *out = t;
            return SCHT_OK;
        }
        if (result[0] == kBuildTooDeep)
            return scht::fail(SCHT_ERR_INVALID_ARG, "scht_build_tree_device: tree deeper than 64 levels; use the host builder (scht_build_tree)");
        if (result[0] == kBuildNodeOverflow || result[0] == kBuildRandOverflow) {
            if (node_cap >= node_max) return scht::fail(SCHT_ERR_INTERNAL, "scht_build_tree_device: node capacity exhausted");
            node_cap *= 4;
        } else if (result[0] == kBuildBetaOverflow) {
            beta_per_node *= 2;
            if (beta_per_node > kBMaxDepth) beta_per_node = kBMaxDepth + 1;
        } else {
            return scht::fail(SCHT_ERR_INTERNAL, "scht_build_tree_device: kernel did not report a status");
        }
        if (attempt > 12) return scht::fail(SCHT_ERR_INTERNAL, "scht_build_tree_device: too many capacity retries");
    }
}
