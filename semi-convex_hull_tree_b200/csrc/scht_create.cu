// Create path of the C ABI (include/scht.h): validates a flat tree description, uploads the (shard of the) tree to one
// GPU and derives the device-only structures of DESIGN.md §3:
//   * dim-major pages (+ row-major copy for wide points) in TREE order  -- exact scan, seed scan, exact re-evaluation
//   * the SCAN-ORDER image of the tensor-core prefilter: resident points grouped into spatially tight scan groups
//     (two Lloyd iterations over strided seeds), FP16 UMMA image, slot -> position table, bounding balls per scan page
//     and per group (k_select_pages prunes with them)
// Replaces GPU_Convex_Tree<T>'s constructor tail (reference: Code/GPU_Convex_Tree.h:150-163,431-479,522-748).
// All O(N*D) passes run on the device; the host only walks the node records.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cub/device/device_radix_sort.cuh>
#include <numeric>
#include <string>
#include <thread>
#include <vector>

#include "scht_handle.h"

using namespace scht;

namespace {

// ---------------------------------------------------------------------------------------------------------
// page builder: row-major sorted points -> dim-major pages (+ padded row-major copy)
// ---------------------------------------------------------------------------------------------------------
__global__ void k_build_pages(const float* __restrict__ rows, int64_t row0, int64_t nrows, int dim, int dpad, int64_t n_points,
                              float* __restrict__ pages, float* __restrict__ rows_out) {
    // one thread per (position, dim) of the pages touched by rows [row0, row0+nrows) rounded to pages
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t total = nrows * dpad;
    if (i >= total) return;
    int64_t r = i / dpad;
    int j = (int)(i - r * dpad);
    int64_t pos = row0 + r;
    float v = (j < dim) ? ((pos < n_points) ? rows[r * dim + j] : kPadCoord) : 0.f;
    int64_t page = pos / kPage;
    int lane = (int)(pos - page * kPage);
    pages[(page * dpad + j) * kPage + lane] = v;
    if (rows_out) rows_out[pos * dpad + j] = v;
}

__device__ __forceinline__ float page_coord(const DevTree& T, int pos, int j) {
    return T.pages[((size_t)(pos >> 8) * T.dpad + j) * kPage + (pos & 255)];
}

// max |p_j - mean_j| per dimension, max |p - mean|^2 and max |p|^2 over the resident positions (order-independent:
// every reduction is a maximum, so the image is reproducible run to run)
__global__ void __launch_bounds__(256) k_shard_maxes(DevTree T, const float* __restrict__ mean, unsigned int* __restrict__ amax_bits,
                                                     unsigned long long* __restrict__ n2max_bits, unsigned long long* __restrict__ xn2max_bits) {
    extern __shared__ unsigned int s_amax[];  // [dpad]
    const int tid = threadIdx.x, lane = tid & 31;
    for (int j = tid; j < T.dpad; j += blockDim.x) s_amax[j] = 0u;
    __syncthreads();
    const long long i = (long long)blockIdx.x * blockDim.x + tid;
    const bool valid = T.pos_lo + i < T.pos_hi;
    const int pos = valid ? (int)(T.pos_lo + i) : T.pos_lo;
    double n2 = 0.0, x2 = 0.0;
    for (int j = 0; j < T.dim; j++) {
        const double v = (double)page_coord(T, pos, j);
        const double c = v - (double)mean[j];
        n2 += c * c;
        x2 += v * v;
        unsigned int a = valid ? __float_as_uint(__double2float_ru(fabs(c))) : 0u;
        a = __reduce_max_sync(0xffffffffu, a);
        if (lane == 0) atomicMax(&s_amax[j], a);
    }
    unsigned long long nb = valid ? (unsigned long long)__double_as_longlong(n2) : 0ull;
    unsigned long long xb = valid ? (unsigned long long)__double_as_longlong(x2) : 0ull;
    for (int o = 16; o; o >>= 1) {
        nb = max(nb, __shfl_xor_sync(0xffffffffu, nb, o));
        xb = max(xb, __shfl_xor_sync(0xffffffffu, xb, o));
    }
    if (lane == 0) {
        atomicMax(n2max_bits, nb);
        atomicMax(xn2max_bits, xb);
    }
    __syncthreads();
    for (int j = tid; j < T.dim; j += blockDim.x) atomicMax(&amax_bits[j], s_amax[j]);
}

// ---------------------------------------------------------------------------------------------------------
// scan groups: Lloyd iterations on the resident points
// ---------------------------------------------------------------------------------------------------------
__global__ void k_gather_seeds(DevTree T, int C, float* __restrict__ cen) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)C * T.dpad) return;
    const int c = (int)(i / T.dpad), j = (int)(i - (long long)c * T.dpad);
    const long long n = (long long)T.pos_hi - T.pos_lo;
    const int pos = T.pos_lo + (int)(((2ll * c + 1) * n) / (2ll * C));  // strided over the tree order: spread like the data
    cen[i] = page_coord(T, pos, j);
}

// nearest centroid of every source point.  ROWS: the points are rows of `src` ([count][dpad], e.g. leaf means),
// else the resident positions first .. first+count of the dim-major pages.  Thread = point; centroids pass through
// shared memory in chunks of 32 x 16 dimensions.
template <bool ROWS>
__global__ void __launch_bounds__(256) k_assign(DevTree T, const float* __restrict__ src, int first, int count, const float* __restrict__ cen, int C,
                                                int* __restrict__ out) {
    __shared__ __align__(16) float cs[32][16];
    const int tid = threadIdx.x;
    const long long i = (long long)blockIdx.x * blockDim.x + tid;
    const bool valid = i < count;
    const int ii = valid ? (int)i : 0;
    const int dpad = T.dpad;
    float best = 3.4e38f;
    int bi = 0;
    for (int c0 = 0; c0 < C; c0 += 32) {
        float acc[32];
#pragma unroll
        for (int ci = 0; ci < 32; ci++) acc[ci] = 0.f;
        for (int j0 = 0; j0 < dpad; j0 += 16) {
            __syncthreads();
            for (int e = tid; e < 512; e += 256) {
                const int ci = e >> 4, u = e & 15, c = c0 + ci, j = j0 + u;
                cs[ci][u] = (c < C && j < dpad) ? cen[(size_t)c * dpad + j] : 0.f;
            }
            __syncthreads();
            float p[16];
#pragma unroll
            for (int u = 0; u < 16; u++) {
                const int j = j0 + u;
                float v = 0.f;
                if (j < dpad) v = ROWS ? src[(size_t)ii * dpad + j] : page_coord(T, first + ii, j);
                p[u] = v;
            }
#pragma unroll
            for (int ci = 0; ci < 32; ci++) {
#pragma unroll
                for (int u4 = 0; u4 < 4; u4++) {
                    const float4 c4 = *reinterpret_cast<const float4*>(&cs[ci][4 * u4]);
                    float d;
                    d = p[4 * u4 + 0] - c4.x; acc[ci] = fmaf(d, d, acc[ci]);
                    d = p[4 * u4 + 1] - c4.y; acc[ci] = fmaf(d, d, acc[ci]);
                    d = p[4 * u4 + 2] - c4.z; acc[ci] = fmaf(d, d, acc[ci]);
                    d = p[4 * u4 + 3] - c4.w; acc[ci] = fmaf(d, d, acc[ci]);
                }
            }
        }
#pragma unroll
        for (int ci = 0; ci < 32; ci++)
            if (c0 + ci < C && acc[ci] < best) {
                best = acc[ci];
                bi = c0 + ci;
            }
    }
    if (valid) out[ii] = bi;
}

__global__ void k_make_keys(const int* __restrict__ grp, int first, int count, unsigned long long* __restrict__ keys) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) keys[i] = ((unsigned long long)(unsigned int)(grp ? grp[i] : 0) << 32) | (unsigned int)(first + (int)i);
}

// gslot[g] = first sorted index whose group id is >= g (g = 0..C); groups are contiguous ranges of the sorted keys
__global__ void k_group_bounds(const unsigned long long* __restrict__ keys, int n, int C, int* __restrict__ gslot) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g > C) return;
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = lo + ((hi - lo) >> 1);
        if ((int)(keys[mid] >> 32) < g) lo = mid + 1;
        else hi = mid;
    }
    gslot[g] = lo;
}

__device__ __forceinline__ float block_sum(float v, float* red) {  // blockDim 256; every thread gets the total
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; w++) t += red[w];
    return t;
}
__device__ __forceinline__ float block_max(float v, float* red) {
    for (int o = 16; o; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; w++) t = fmaxf(t, red[w]);
    return t;
}

// centroid of every group = mean of its members (fixed reduction order: reproducible); an empty group keeps its centroid
__global__ void __launch_bounds__(256) k_group_means(DevTree T, const unsigned long long* __restrict__ keys, const int* __restrict__ gslot,
                                                     float* __restrict__ cen) {
    __shared__ float red[8];
    const int g = blockIdx.x, b = gslot[g], e = gslot[g + 1];
    if (e <= b) return;
    for (int j = 0; j < T.dpad; j++) {
        float s = 0.f;
        for (int i = b + threadIdx.x; i < e; i += blockDim.x) s += page_coord(T, (int)(unsigned int)keys[i], j);
        s = block_sum(s, red);
        if (threadIdx.x == 0) cen[(size_t)g * T.dpad + j] = s / (float)(e - b);
    }
}

// key (old group, position) -> key (order index of the group, position)
__global__ void k_remap_keys(const unsigned long long* __restrict__ in, int n, const int* __restrict__ newid, unsigned long long* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = ((unsigned long long)(unsigned int)newid[(int)(in[i] >> 32)] << 32) | (unsigned int)in[i];
}
// sorted keys -> slots: member i of group g goes to goff[g] + (i - gstart[g]); slot_pos was filled with -1 (padding)
__global__ void k_scatter_slots(const unsigned long long* __restrict__ keys, int n, const int* __restrict__ gstart, const int* __restrict__ goff,
                                int* __restrict__ slot_pos) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int g = (int)(keys[i] >> 32);
    slot_pos[goff[g] + ((int)i - gstart[g])] = (int)(unsigned int)keys[i];
}
__global__ void k_iota_slots(int first, int n, int n_slots, int* __restrict__ slot_pos) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_slots) slot_pos[i] = (i < n) ? first + (int)i : -1;
}

// bounding ball of every scan page: centre = mean of its points, radius = max distance (rounded up, ball_eps); pcnt = points
__global__ void __launch_bounds__(256) k_page_balls(DevTree T, const int* __restrict__ slot_pos, float* __restrict__ pcen, float* __restrict__ prad,
                                                    int* __restrict__ pcnt) {
    extern __shared__ float cen_s[];  // [dpad]
    __shared__ float red[8];
    const int sp = blockIdx.x;
    const int pos = slot_pos[(size_t)sp * kPage + threadIdx.x];
    const bool valid = pos >= 0;
    const float cnt = block_sum(valid ? 1.f : 0.f, red);
    for (int j = 0; j < T.dpad; j++) {
        const float s = block_sum(valid ? page_coord(T, pos, j) : 0.f, red);
        if (threadIdx.x == 0) {
            const float c = s / fmaxf(cnt, 1.f);
            cen_s[j] = c;
            pcen[(size_t)sp * T.dpad + j] = c;
        }
    }
    __syncthreads();
    float d2 = 0.f;
    if (valid)
        for (int j = 0; j < T.dpad; j++) {
            const float d = page_coord(T, pos, j) - cen_s[j];
            d2 = fmaf(d, d, d2);
        }
    const float m = block_max(d2, red);
    if (threadIdx.x == 0) {
        prad[sp] = sqrtf(m) * (1.f + ball_eps(T.dim)) + 1e-37f;
        pcnt[sp] = (int)cnt;
    }
}

// bounding ball of a run of slots [first[b], first[b+1]) * unit (unit = 256: super-groups given by page boundaries)
// centre: mean of the points of the run
__global__ void __launch_bounds__(256) k_run_balls(DevTree T, const int* __restrict__ slot_pos, const int* __restrict__ first, int unit, float* __restrict__ cen,
                                                   float* __restrict__ rad, int* __restrict__ count) {
    extern __shared__ float cen_s[];  // [dpad]
    __shared__ float red[8];
    const int g = blockIdx.x;
    const long long b = (long long)first[g] * unit, e = (long long)first[g + 1] * unit;
    float cnt = 0.f;
    for (long long i = b + threadIdx.x; i < e; i += blockDim.x) cnt += slot_pos[i] >= 0 ? 1.f : 0.f;
    cnt = block_sum(cnt, red);
    for (int j = 0; j < T.dpad; j++) {
        float s = 0.f;
        for (long long i = b + threadIdx.x; i < e; i += blockDim.x) {
            const int pos = slot_pos[i];
            if (pos >= 0) s += page_coord(T, pos, j);
        }
        s = block_sum(s, red);
        if (threadIdx.x == 0) {
            const float c = s / fmaxf(cnt, 1.f);
            cen_s[j] = c;
            cen[(size_t)g * T.dpad + j] = c;
        }
    }
    __syncthreads();
    float m = 0.f;
    for (long long i = b + threadIdx.x; i < e; i += blockDim.x) {
        const int pos = slot_pos[i];
        if (pos < 0) continue;
        float d2 = 0.f;
        for (int j = 0; j < T.dpad; j++) {
            const float d = page_coord(T, pos, j) - cen_s[j];
            d2 = fmaf(d, d, d2);
        }
        m = fmaxf(m, d2);
    }
    m = block_max(m, red);
    if (threadIdx.x == 0) {
        rad[g] = sqrtf(m) * (1.f + ball_eps(T.dim)) + 1e-37f;
        count[g] = (int)cnt;
    }
}

// radius of every k-means group around its centroid (members = sorted keys [gslot[g], gslot[g+1]))
__global__ void __launch_bounds__(256) k_group_radii(DevTree T, const unsigned long long* __restrict__ keys, const int* __restrict__ gslot,
                                                     const float* __restrict__ gcen, float* __restrict__ grad) {
    extern __shared__ float cen_s[];  // [dpad]
    __shared__ float red[8];
    const int g = blockIdx.x, b = gslot[g], e = gslot[g + 1];
    for (int j = threadIdx.x; j < T.dpad; j += blockDim.x) cen_s[j] = gcen[(size_t)g * T.dpad + j];
    __syncthreads();
    float m = 0.f;
    for (int i = b + threadIdx.x; i < e; i += blockDim.x) {
        const int pos = (int)(unsigned int)keys[i];
        float d2 = 0.f;
        for (int j = 0; j < T.dpad; j++) {
            const float d = page_coord(T, pos, j) - cen_s[j];
            d2 = fmaf(d, d, d2);
        }
        m = fmaxf(m, d2);
    }
    m = block_max(m, red);
    if (threadIdx.x == 0) grad[g] = sqrtf(m);
}

// exact coordinates, row-major, in scan order (one thread per (slot, 4 dims))
__global__ void k_build_srows(DevTree T, const int* __restrict__ slot_pos, long long n_slots, float* __restrict__ srows) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int q4 = T.dpad >> 2;
    if (i >= n_slots * q4) return;
    const long long slot = i / q4;
    const int j = (int)(i - slot * q4) * 4;
    const int pos = slot_pos[slot];
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (pos >= 0) {
        if (T.rows) v = *reinterpret_cast<const float4*>(T.rows + (size_t)pos * T.dpad + j);
        else v = make_float4(page_coord(T, pos, j), page_coord(T, pos, j + 1), page_coord(T, pos, j + 2), page_coord(T, pos, j + 3));
    }
    *reinterpret_cast<float4*>(srows + (size_t)slot * T.dpad + j) = v;
}

// -------------------------------------------------------------------------------------------------------------
// tensor-core page image in SCAN order (one thread per slot): row r of scan page sp holds fp16(s (p_j - mean_j)) and, in
// three extra K slots, s^2 |p - mean|^2 / u as an exact hi + mid + lo FP16 split (see scht_tc.cuh)
// -------------------------------------------------------------------------------------------------------------
__global__ void k_build_tcpages(DevTree T, const int* __restrict__ slot_pos, int n_spages, const float* __restrict__ mean, float scale, float norm_unit,
                                int kp, unsigned char* __restrict__ tcpages) {
    const long long slot = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= (long long)n_spages * kPage) return;
    const int page = (int)(slot / kPage), r = (int)(slot % kPage);
    const int pos = slot_pos[slot];
    const bool real = pos >= 0;
    __half* img = reinterpret_cast<__half*>(tcpages + (size_t)page * tc_page_bytes(kp));
    double n2 = 0.0, rest = 0.0;
    for (int k = 0; k < kp; k++) {
        __half hv = __float2half_rn(0.f);
        if (k < T.dim) {
            if (real) {
                double c = (double)page_coord(T, pos, k) - (double)mean[k];
                n2 += c * c;
                hv = __float2half_rn((float)(c * (double)scale));
            }
        } else if (k < T.dim + 3) {
            // s^2 |p'|^2 / u as hi + mid + lo (each FP16, together 33 bits); padded slots: a value no bound reaches
            if (k == T.dim) rest = real ? n2 * (double)scale * (double)scale / (double)norm_unit : 60000.0;
            hv = __float2half_rn((float)rest);
            rest -= (double)__half2float(hv);
        }
        img[(((size_t)(k >> 3) * 32 + (r >> 3)) * 8 + (r & 7)) * 8 + (k & 7)] = hv;
    }
}

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
template <class T>
int dev_alloc(scht_handle* h, size_t count, T** dst) {
    void* d = nullptr;
    size_t bytes = std::max<size_t>(count, 1) * sizeof(T);
    CU_TRY(cudaMalloc(&d, bytes));
    h->owned.push_back(d);
    h->tree_bytes += bytes;
    *dst = reinterpret_cast<T*>(d);
    return SCHT_OK;
}

int env_int(const char* name, int dflt) {
    const char* ev = getenv(name);
    return ev ? atoi(ev) : dflt;
}

// mean of a strided sample of the resident points (block j = dimension j; fixed summation order: reproducible)
__global__ void __launch_bounds__(256) k_sample_mean(DevTree T, int ns, float* __restrict__ mean) {
    __shared__ float red[8];
    const int j = blockIdx.x;
    const long long n = (long long)T.pos_hi - T.pos_lo;
    float s = 0.f;
    for (int i = threadIdx.x; i < ns; i += blockDim.x) s += page_coord(T, T.pos_lo + (int)(((2ll * i + 1) * n) / (2ll * ns)), j);
    s = block_sum(s, red);
    if (threadIdx.x == 0) mean[j] = (j < T.dim) ? s / (float)ns : 0.f;
}

#define TRY_H(x)                          \
    do {                                  \
        int rc__ = (x);                   \
        if (rc__) return rc__;            \
    } while (0)

// Order of the k-means groups: recursive bisection of their centroids (two far-apart pivots, every centroid to the nearer
// one -- the split rule of the tree itself, Code/Convex_Tree.h:789-821, applied to a few thousand centroids on the host), read
// off depth first: neighbours in the order are neighbours in space.  Empty groups are left out.
void order_groups(const std::vector<float>& cen, int dpad, const std::vector<int>& size, std::vector<int>& order) {
    std::vector<int> ids;
    for (int g = 0; g < (int)size.size(); g++)
        if (size[g] > 0) ids.push_back(g);
    order.clear();
    order.reserve(ids.size());
    auto dist2 = [&](int a, const double* m) {
        double s = 0;
        const float* p = &cen[(size_t)a * dpad];
        for (int j = 0; j < dpad; j++) {
            const double d = (double)p[j] - m[j];
            s += d * d;
        }
        return s;
    };
    auto dist2g = [&](int a, int b) {
        double s = 0;
        const float *p = &cen[(size_t)a * dpad], *q = &cen[(size_t)b * dpad];
        for (int j = 0; j < dpad; j++) {
            const double d = (double)p[j] - (double)q[j];
            s += d * d;
        }
        return s;
    };
    std::vector<std::pair<int, int>> stack;  // ranges [b, e) of ids still to split (processed right to left so the order is DFS)
    stack.emplace_back(0, (int)ids.size());
    std::vector<double> mean(dpad);
    std::vector<int> tmp;
    while (!stack.empty()) {
        const int b = stack.back().first, e = stack.back().second;
        stack.pop_back();
        if (e - b <= 1) {
            if (e > b) order.push_back(ids[b]);
            continue;
        }
        std::fill(mean.begin(), mean.end(), 0.0);
        for (int i = b; i < e; i++)
            for (int j = 0; j < dpad; j++) mean[j] += cen[(size_t)ids[i] * dpad + j];
        for (int j = 0; j < dpad; j++) mean[j] /= (e - b);
        int pa = b, pb = b;
        double best = -1;
        for (int i = b; i < e; i++) {
            const double d = dist2(ids[i], mean.data());
            if (d > best) best = d, pa = i;
        }
        best = -1;
        for (int i = b; i < e; i++) {
            const double d = dist2g(ids[i], ids[pa]);
            if (d > best) best = d, pb = i;
        }
        const int ga = ids[pa], gb = ids[pb];
        tmp.clear();
        int w = b;
        for (int i = b; i < e; i++) {
            if (dist2g(ids[i], ga) <= dist2g(ids[i], gb)) ids[w++] = ids[i];
            else tmp.push_back(ids[i]);
        }
        if (w == b || w == e) {  // all centroids coincide: keep them as they are
            for (size_t k = 0; k < tmp.size(); k++) ids[w + k] = tmp[k];
            for (int i = b; i < e; i++) order.push_back(ids[i]);
            continue;
        }
        for (size_t k = 0; k < tmp.size(); k++) ids[w + k] = tmp[k];
        stack.emplace_back(w, e);  // right half later
        stack.emplace_back(b, w);  // left half first
    }
}

// ---- the scan-order image (ScanImage) of the resident points -----------------------------------------------
int build_scan_image(scht_handle* h, const float* d_mean, const std::vector<float>& pmax_h, double n2max, const float* leaf_means) {
    DevTree& T = h->T;
    ScanImage& I = h->I;
    cudaStream_t s = h->stream;
    const int dim = T.dim, dpad = T.dpad;
    const int n = T.pos_hi - T.pos_lo;
    const int kp = (dim + 3 + 15) & ~15;
    size_t free_b = 0, total_b = 0;
    cudaMemGetInfo(&free_b, &total_b);
    if ((size_t)((double)n * 1.1 / kPage + 2) * (tc_page_bytes(kp) + (size_t)kPage * dpad * 4) + (size_t)n * 28 >= free_b / 2) return SCHT_OK;  // no room for the image: the exact FP32 scan serves every batch

    double amax_all = 0, asum = 0;
    for (int j = 0; j < dim; j++) {
        amax_all = std::max(amax_all, (double)pmax_h[j]);
        asum += pmax_h[j];
    }
    I.pnorm_max = (float)(std::sqrt(n2max) * 1.0001);
    I.pabs_sum = (float)(asum * 1.0001);
    // power-of-two scale that puts the largest centred coordinate into (128, 256]: well inside FP16's range
    I.scale = (amax_all > 0 && std::isfinite(amax_all)) ? (float)std::exp2(std::floor(std::log2(256.0 / amax_all))) : 1.f;
    const double sn = (double)I.scale * I.scale * n2max;
    // norm slots hold s^2|p'|^2 / u with u a power of two that keeps them below 16384 (FP16 range, room for 60000 on padding)
    I.norm_unit = (sn > 16384.0) ? (float)std::exp2(std::ceil(std::log2(sn / 16384.0))) : 1.f;
    // s^2 must stay a normal float in the filter bound (k_scan_tc::bound_of): 2^-60 <= s <= 2^60
    if (!(std::isfinite(I.pnorm_max) && std::isfinite(I.scale) && I.scale >= 8.6736174e-19f && I.scale <= 1.1529215e18f && sn < 1e37)) return SCHT_OK;

    // ---- k-means groups (build-time only): ~512 points each ----
    int C = h->opt.groups;
    if (C < 0) C = (n < 8192) ? 1 : std::min(32768, n / 512);
    C = std::max(1, std::min(C, std::max(1, n / kPage)));
    unsigned long long *keys_a = nullptr, *keys_b = nullptr;
    int *grp = nullptr, *d_newid = nullptr, *d_gstart = nullptr, *d_goff = nullptr;
    float *d_gcen = nullptr, *d_grad = nullptr, *d_lm = nullptr;
    int* d_lg = nullptr;
    void* cub_tmp = nullptr;
    size_t cub_bytes = 0;
    int end_bit = 32;
    for (int c = C - 1; c > 0; c >>= 1) end_bit++;
    auto free_tmp = [&]() {
        for (void* p : {(void*)keys_a, (void*)keys_b, (void*)grp, (void*)d_newid, (void*)d_gstart, (void*)d_goff, (void*)d_gcen, (void*)d_grad, (void*)d_lm, (void*)d_lg, cub_tmp})
            if (p) cudaFree(p);
    };
#define CU_T(expr)                                                                                          \
    do {                                                                                                     \
        cudaError_t e__ = (expr);                                                                            \
        if (e__ != cudaSuccess) {                                                                            \
            free_tmp();                                                                                      \
            return scht::fail(e__ == cudaErrorMemoryAllocation ? SCHT_ERR_OOM : SCHT_ERR_CUDA,              \
                              std::string(#expr) + ": " + cudaGetErrorString(e__));                         \
        }                                                                                                    \
    } while (0)
#define TRY_T(x)            \
    do {                    \
        int rc__ = (x);     \
        if (rc__) {         \
            free_tmp();     \
            return rc__;    \
        }                   \
    } while (0)
    const unsigned nb = (unsigned)((n + 255) / 256);
    std::vector<int> goff_h, spage_h;   // slot offset of every ordered group; first page of every super-group
    std::vector<float> cen_ord;         // centroids in scan order (bucket order of the leaves)
    int n_spages = (n + kPage - 1) / kPage;
    std::vector<char> hard;             // hard[p]: a page break (padding) ends right before page p
    int* slot_pos = nullptr;
    if (C > 1) {
        CU_T(cudaMalloc(&keys_a, (size_t)n * 8));
        CU_T(cudaMalloc(&keys_b, (size_t)n * 8));
        CU_T(cudaMalloc(&grp, (size_t)n * 4));
        CU_T(cudaMalloc(&d_gcen, (size_t)C * dpad * 4));
        CU_T(cudaMalloc(&d_grad, (size_t)C * 4));
        CU_T(cudaMalloc(&d_gstart, (size_t)(C + 1) * 4));
        CU_T(cudaMalloc(&d_goff, (size_t)(C + 1) * 4));
        CU_T(cudaMalloc(&d_newid, (size_t)C * 4));
        CU_T(cub::DeviceRadixSort::SortKeys(nullptr, cub_bytes, keys_a, keys_b, n, 0, end_bit, s));
        CU_T(cudaMalloc(&cub_tmp, std::max<size_t>(cub_bytes, 16)));
        k_gather_seeds<<<(unsigned)(((size_t)C * dpad + 255) / 256), 256, 0, s>>>(T, C, d_gcen);
        for (int it = 0; it < 2; it++) {  // assign -> sort by (group, position) -> centroids = member means
            k_assign<false><<<nb, 256, 0, s>>>(T, nullptr, T.pos_lo, n, d_gcen, C, grp);
            k_make_keys<<<nb, 256, 0, s>>>(grp, T.pos_lo, n, keys_a);
            CU_T(cub::DeviceRadixSort::SortKeys(cub_tmp, cub_bytes, keys_a, keys_b, n, 0, end_bit, s));
            k_group_bounds<<<(C + 1 + 255) / 256, 256, 0, s>>>(keys_b, n, C, d_gstart);
            k_group_means<<<C, 256, 0, s>>>(T, keys_b, d_gstart, d_gcen);
            CU_T(cudaGetLastError());
        }
        k_group_radii<<<C, 256, dpad * sizeof(float), s>>>(T, keys_b, d_gstart, d_gcen, d_grad);
        std::vector<float> cen_h((size_t)C * dpad), rad_h(C);
        std::vector<int> gstart_h(C + 1), size_h(C);
        CU_T(cudaMemcpyAsync(cen_h.data(), d_gcen, cen_h.size() * 4, cudaMemcpyDeviceToHost, s));
        CU_T(cudaMemcpyAsync(rad_h.data(), d_grad, rad_h.size() * 4, cudaMemcpyDeviceToHost, s));
        CU_T(cudaMemcpyAsync(gstart_h.data(), d_gstart, gstart_h.size() * 4, cudaMemcpyDeviceToHost, s));
        CU_T(cudaStreamSynchronize(s));
        for (int g = 0; g < C; g++) size_h[g] = gstart_h[g + 1] - gstart_h[g];
        // ---- host: order of the groups, page breaks, super-groups ----
        std::vector<int> order;
        order_groups(cen_h, dpad, size_h, order);
        const int Cn = (int)order.size();
        std::vector<int> newid(C, 0);
        for (int i = 0; i < Cn; i++) newid[order[i]] = i;
        // a page break after group i when empty space separates it from group i+1 (gap between the two balls larger than half
        // the bigger radius): the padding keeps a page from straddling two clusters.  Widest gaps first, at most 8 % padding.
        std::vector<std::pair<double, int>> cand;
        for (int i = 0; i + 1 < Cn; i++) {
            const int a = order[i], b = order[i + 1];
            double d2 = 0;
            for (int j = 0; j < dpad; j++) {
                const double d = (double)cen_h[(size_t)a * dpad + j] - (double)cen_h[(size_t)b * dpad + j];
                d2 += d * d;
            }
            const double gap = std::sqrt(d2) - rad_h[a] - rad_h[b];
            if (gap > 0.5 * std::max(rad_h[a], rad_h[b])) cand.emplace_back(gap, i);
        }
        std::sort(cand.begin(), cand.end(), [](const std::pair<double, int>& x, const std::pair<double, int>& y) { return x.first > y.first; });
        const size_t max_breaks = (size_t)((double)n * 0.08 / (kPage / 2));
        std::vector<char> brk(Cn, 0);
        for (size_t k = 0; k < cand.size() && k < max_breaks; k++) brk[cand[k].second] = 1;
        goff_h.assign(Cn + 1, 0);
        long long off = 0;
        std::vector<long long> hard_slots;
        for (int i = 0; i < Cn; i++) {
            goff_h[i] = (int)off;
            off += size_h[order[i]];
            if (brk[i]) {
                off = (off + kPage - 1) / kPage * kPage;
                hard_slots.push_back(off);
            }
        }
        goff_h[Cn] = (int)off;
        n_spages = (int)((off + kPage - 1) / kPage);
        hard.assign(n_spages + 1, 0);
        for (long long x : hard_slots)
            if (x / kPage <= n_spages) hard[x / kPage] = 1;
        cen_ord.resize((size_t)Cn * dpad);
        for (int i = 0; i < Cn; i++) memcpy(&cen_ord[(size_t)i * dpad], &cen_h[(size_t)order[i] * dpad], sizeof(float) * dpad);
        // ---- device: keys by (order index, position) -> slots ----
        CU_T(cudaMemcpyAsync(d_newid, newid.data(), (size_t)C * 4, cudaMemcpyHostToDevice, s));
        CU_T(cudaMemcpyAsync(d_goff, goff_h.data(), (size_t)(Cn + 1) * 4, cudaMemcpyHostToDevice, s));
        k_remap_keys<<<nb, 256, 0, s>>>(keys_b, n, d_newid, keys_a);
        CU_T(cub::DeviceRadixSort::SortKeys(cub_tmp, cub_bytes, keys_a, keys_b, n, 0, end_bit, s));
        k_group_bounds<<<(Cn + 1 + 255) / 256, 256, 0, s>>>(keys_b, n, Cn, d_gstart);
        TRY_T(dev_alloc(h, (size_t)n_spages * kPage, &slot_pos));
        CU_T(cudaMemsetAsync(slot_pos, 0xff, (size_t)n_spages * kPage * 4, s));
        k_scatter_slots<<<nb, 256, 0, s>>>(keys_b, n, d_gstart, d_goff, slot_pos);
        CU_T(cudaGetLastError());
    } else {
        hard.assign(n_spages + 1, 0);
        TRY_T(dev_alloc(h, (size_t)n_spages * kPage, &slot_pos));
        k_iota_slots<<<(unsigned)(((size_t)n_spages * kPage + 255) / 256), 256, 0, s>>>(T.pos_lo, n, n_spages * kPage, slot_pos);  // one group: tree order
        CU_T(cudaGetLastError());
    }
    // super-groups: runs of <= kSuperPages pages that do not cross a page break
    spage_h.push_back(0);
    for (int p = 1; p <= n_spages; p++)
        if (p == n_spages || hard[p] || p - spage_h.back() >= kSuperPages) spage_h.push_back(p);
    const int S = (int)spage_h.size() - 1;

    float *pcen = nullptr, *prad = nullptr, *scen = nullptr, *srad = nullptr, *srows = nullptr;
    int *pcnt = nullptr, *spage = nullptr, *scnt = nullptr;
    unsigned char* tcp = nullptr;
    TRY_T(dev_alloc(h, (size_t)n_spages * dpad, &pcen));
    TRY_T(dev_alloc(h, (size_t)n_spages, &prad));
    TRY_T(dev_alloc(h, (size_t)n_spages, &pcnt));
    TRY_T(dev_alloc(h, (size_t)S * dpad, &scen));
    TRY_T(dev_alloc(h, (size_t)S, &srad));
    TRY_T(dev_alloc(h, (size_t)S + 1, &spage));
    TRY_T(dev_alloc(h, (size_t)S, &scnt));
    TRY_T(dev_alloc(h, (size_t)n_spages * tc_page_bytes(kp), &tcp));
    TRY_T(dev_alloc(h, (size_t)n_spages * kPage * dpad, &srows));
    CU_T(cudaMemcpyAsync(spage, spage_h.data(), (size_t)(S + 1) * 4, cudaMemcpyHostToDevice, s));
    k_build_tcpages<<<(unsigned)(((size_t)n_spages * kPage + 255) / 256), 256, 0, s>>>(T, slot_pos, n_spages, d_mean, I.scale, I.norm_unit, kp, tcp);
    k_build_srows<<<(unsigned)(((size_t)n_spages * kPage * (dpad >> 2) + 255) / 256), 256, 0, s>>>(T, slot_pos, (long long)n_spages * kPage, srows);
    k_page_balls<<<n_spages, 256, dpad * sizeof(float), s>>>(T, slot_pos, pcen, prad, pcnt);
    k_run_balls<<<S, 256, dpad * sizeof(float), s>>>(T, slot_pos, spage, kPage, scen, srad, scnt);
    CU_T(cudaGetLastError());

    // ---- bucket order of the leaves: by (ordered scan group nearest to the leaf's mean, leaf id), so that the queries of a
    // tile (bucketed by approximate leaf) want the same scan pages ----
    std::vector<int> rank(T.n_leaves);
    std::iota(rank.begin(), rank.end(), 0);
    if (C > 1 && leaf_means && !cen_ord.empty()) {
        const int Cn = (int)(cen_ord.size() / dpad);
        std::vector<float> lm((size_t)T.n_leaves * dpad, 0.f);
        for (int l = 0; l < T.n_leaves; l++) memcpy(&lm[(size_t)l * dpad], leaf_means + (size_t)l * dim, sizeof(float) * dim);
        CU_T(cudaMalloc(&d_lm, lm.size() * 4));
        CU_T(cudaMalloc(&d_lg, (size_t)T.n_leaves * 4));
        CU_T(cudaMemcpyAsync(d_lm, lm.data(), lm.size() * 4, cudaMemcpyHostToDevice, s));
        CU_T(cudaMemcpyAsync(d_gcen, cen_ord.data(), cen_ord.size() * 4, cudaMemcpyHostToDevice, s));
        k_assign<true><<<(T.n_leaves + 255) / 256, 256, 0, s>>>(T, d_lm, 0, T.n_leaves, d_gcen, Cn, d_lg);
        std::vector<int> lg(T.n_leaves);
        CU_T(cudaMemcpyAsync(lg.data(), d_lg, (size_t)T.n_leaves * 4, cudaMemcpyDeviceToHost, s));
        CU_T(cudaStreamSynchronize(s));
        std::vector<int> order(T.n_leaves);
        std::iota(order.begin(), order.end(), 0);
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return lg[a] < lg[b]; });
        for (int i = 0; i < T.n_leaves; i++) rank[order[i]] = i;
    }
    CU_T(cudaStreamSynchronize(s));
    free_tmp();
#undef CU_T
#undef TRY_T
    CU_TRY(cudaMemcpy(const_cast<int*>(T.leaf_rank), rank.data(), (size_t)T.n_leaves * 4, cudaMemcpyHostToDevice));
    I.tcpages = tcp;
    I.slot_pos = slot_pos;
    I.srows = srows;
    I.pcnt = pcnt;
    I.pcen = pcen;
    I.prad = prad;
    I.scen = scen;
    I.srad = srad;
    I.spage = spage;
    I.scnt = scnt;
    I.n_slots = n_spages * kPage;
    I.n_spages = n_spages;
    I.n_supers = S;
    I.n_groups = C;
    I.mean = d_mean;
    I.kp = kp;
    return SCHT_OK;
}

}  // namespace

namespace scht {

void options_from_env(Options& o) {
    o.scan_variant = std::max(-1, std::min(3, env_int("SCHT_SCAN_VARIANT", o.scan_variant)));
    o.scan_stages = env_int("SCHT_SCAN_STAGES", o.scan_stages);
    o.traverse_sampling = env_int("SCHT_TRAVERSE_SAMPLING", o.traverse_sampling) != 0;
    o.sample_reuse = std::max(0, env_int("SCHT_SAMPLE_REUSE", o.sample_reuse));
    o.tc_mode = std::max(0, std::min(2, env_int("SCHT_TC", o.tc_mode)));
    o.tc_split = std::max(0, env_int("SCHT_TC_SPLIT", o.tc_split));
    o.tc_issuers = env_int("SCHT_TC_ISSUERS", o.tc_issuers);
    o.tc_n256 = env_int("SCHT_TC_N256", o.tc_n256);
    o.seed_points = std::max(1, env_int("SCHT_SEED_POINTS", o.seed_points));
    o.groups = env_int("SCHT_GROUPS", o.groups);
    o.page_select = env_int("SCHT_PAGE_SELECT", o.page_select) != 0;
    o.brute_tc = env_int("SCHT_BRUTE_TC", o.brute_tc) != 0;
    o.pipeline = env_int("SCHT_PIPELINE", o.pipeline) != 0;
    o.pipeline_split = std::max(0, env_int("SCHT_PIPELINE_SPLIT", o.pipeline_split));
    o.first_pass = std::max(-1, std::min(2, env_int("SCHT_FIRST_PASS", o.first_pass)));
    o.rescorer_sleep_ns = std::max(0, env_int("SCHT_RESCORER_SLEEP_NS", o.rescorer_sleep_ns));
    o.tc_recheck = env_int("SCHT_TC_RECHECK", o.tc_recheck) ? 1 : 0;
    o.lazy_seed = env_int("SCHT_LAZY_SEED", o.lazy_seed) ? 1 : 0;
    o.ring_wait_ns = std::max(0, env_int("SCHT_RING_WAIT_NS", o.ring_wait_ns));
    {
        const int ring = env_int("SCHT_TC_RING", o.tc_ring);
        o.tc_ring = (ring >= 32 && ring <= 1024 && (ring & (ring - 1)) == 0) ? ring : 0;
    }
}

// seed neighbourhood of every leaf: the smallest enclosing subtree with >= opt.seed_points points
// (512 points, about one tree level above a 500-point leaf.  Against 1024: 1Mx16 step 11.41 -> 11.29 ms, Gaussian clusters
// 2Mx32 47.3 -> 46.2 ms, SIFT-like 1Mx128 46.7 -> 44.0 ms; 256: 11.21 ms on 1Mx16, 2048: 11.74 ms)
int rebuild_seed_ranges(scht_handle* h) {
    const int L = h->T.n_leaves, nn = h->T.n_nodes;
    std::vector<int> ss(L), se(L);
    for (int i = 0; i < nn; i++) {
        const NodeRec& r = h->h_rec[i];
        if (r.leaf < 0) continue;
        int a = i;
        while (h->h_rec[a].pos1 - h->h_rec[a].pos0 < h->opt.seed_points && h->h_parent[a] >= 0) a = h->h_parent[a];
        ss[r.leaf] = h->h_rec[a].pos0;
        se[r.leaf] = h->h_rec[a].pos1;
    }
    CU_TRY(cudaSetDevice(h->device));
    CU_TRY(cudaStreamSynchronize(h->stream));
    CU_TRY(cudaMemcpy(const_cast<int*>(h->T.seed_start), ss.data(), (size_t)L * 4, cudaMemcpyHostToDevice));
    CU_TRY(cudaMemcpy(const_cast<int*>(h->T.seed_end), se.data(), (size_t)L * 4, cudaMemcpyHostToDevice));
    return SCHT_OK;
}

// per-leaf means of the whole data set (host, all cores): the bucket order of the leaves is derived from them
void compute_leaf_means(const scht_tree_desc* t, std::vector<float>& out) {
    const int L = t->n_leaves, dim = t->dim;
    out.assign((size_t)L * dim, 0.f);
    const int nthr = (int)std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    auto work = [&](int tid) {
        std::vector<double> acc(dim);
        for (int l = tid; l < L; l += nthr) {
            std::fill(acc.begin(), acc.end(), 0.0);
            const int b = t->leaf_start[l], c = t->leaf_count[l];
            for (int i = 0; i < c; i++) {
                const float* p = t->sorted_points + (size_t)(b + i) * dim;
                for (int j = 0; j < dim; j++) acc[j] += p[j];
            }
            for (int j = 0; j < dim; j++) out[(size_t)l * dim + j] = c > 0 ? (float)(acc[j] / c) : 0.f;
        }
    };
    if (nthr == 1 || (int64_t)t->n_points * dim < (1 << 22)) {
        for (int i = 0; i < nthr; i++) work(i);
        return;
    }
    std::vector<std::thread> th;
    for (int i = 0; i < nthr; i++) th.emplace_back(work, i);
    for (auto& x : th) x.join();
}

int create_handle(const scht_tree_desc* t, int device, int leaf_begin, int leaf_end, const Options* opt_in, const float* leaf_means_in,
                  scht_handle** out) {
    if (!out) return scht::fail(SCHT_ERR_INVALID_ARG, "scht_create: out is NULL");
    *out = nullptr;
    if (!t || t->dim < 1 || t->n_points < 1 || t->n_points > 0x7fffff00LL || t->n_nodes < 1 || t->n_leaves < 1 || !t->sorted_points ||
        !t->sorted_to_orig || !t->leaf_start || !t->leaf_count || !t->leaf_node_id || !t->node_left || !t->node_right ||
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
    if (leaf_begin < 0 || leaf_end > L || leaf_begin >= leaf_end) return scht::fail(SCHT_ERR_INVALID_ARG, "scht_create: bad leaf range");
    // ---- validate the structure and derive per-node records (host).  Parents are derived from the validated pre-order links;
    // the caller's node_parent is not trusted (a cyclic array must not hang or crash the library). ------------------------
    std::vector<NodeRec> rec((size_t)nn);
    std::vector<int> parent((size_t)nn, -1);
    int next_leaf = 0;
    for (int i = 0; i < nn; i++) {
        bool leaf = t->node_leaf_index[i] >= 0;
        if (leaf != (t->node_left[i] < 0) || leaf != (t->node_right[i] < 0)) return scht::fail(SCHT_ERR_INVALID_ARG, "scht_create: node child links inconsistent");
        if (!leaf && (t->node_left[i] != i + 1 || t->node_right[i] <= i + 1 || t->node_right[i] >= nn))
            return scht::fail(SCHT_ERR_INVALID_ARG, "scht_create: nodes must be numbered in DFS pre-order (left child = id+1)");
        if (leaf && t->node_leaf_index[i] != next_leaf++) return scht::fail(SCHT_ERR_INVALID_ARG, "scht_create: leaf ids must ascend in node order");
        if (t->node_beta_off[i + 1] < t->node_beta_off[i] || t->node_beta_off[i] < 0) return scht::fail(SCHT_ERR_INVALID_ARG, "scht_create: node_beta_off must ascend");
        rec[i].depth = t->node_beta_off[i + 1] - t->node_beta_off[i];
        rec[i].beta_off = t->node_beta_off[i];
        rec[i].leaf = leaf ? t->node_leaf_index[i] : -1;
        rec[i].right = leaf ? -1 : t->node_right[i];
        rec[i].pad = 0;
        if (!leaf) {
            parent[i + 1] = i;
            if (parent[t->node_right[i]] >= 0) return scht::fail(SCHT_ERR_INVALID_ARG, "scht_create: a node has two parents");
            parent[t->node_right[i]] = i;
        }
    }
    if (next_leaf != L) return scht::fail(SCHT_ERR_INVALID_ARG, "scht_create: n_leaves does not match the nodes");
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
    h->max_smem = (int)prop.sharedMemPerBlockOptin;
    h->n_sm = std::max(1, prop.multiProcessorCount);
    h->depth = depth;
    h->leaf_lo = leaf_begin;
    h->leaf_hi = leaf_end;
    h->h_leaf_start.assign(t->leaf_start, t->leaf_start + L);
    h->h_leaf_count.assign(t->leaf_count, t->leaf_count + L);
    h->h_rec = rec;
    h->h_parent = parent;
    if (opt_in) h->opt = *opt_in;
    else options_from_env(h->opt);
    auto body = [&]() -> int {
        CU_TRY(cudaSetDevice(device));
        CU_TRY(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
        CU_TRY(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
        CU_TRY(cudaStreamCreateWithFlags(&h->out_stream, cudaStreamNonBlocking));
        for (auto& e : h->ev) CU_TRY(cudaEventCreate(&e));
        for (int i = 0; i < 2; i++) {
            CU_TRY(cudaEventCreateWithFlags(&h->ev_in[i], cudaEventDisableTiming));
            CU_TRY(cudaEventCreateWithFlags(&h->ev_done[i], cudaEventDisableTiming));
            CU_TRY(cudaEventCreateWithFlags(&h->ev_out[i], cudaEventDisableTiming));
        }
        DevTree& T = h->T;
        T.dim = dim;
        T.dpad = (dim + 3) & ~3;
        T.n_points = (int)N;
        T.n_pages = (int)((N + kPage - 1) / kPage);
        T.n_nodes = nn;
        T.n_leaves = L;
        T.pos_lo = t->leaf_start[leaf_begin];
        T.pos_hi = t->leaf_start[leaf_end - 1] + t->leaf_count[leaf_end - 1];
        T.page_lo = T.pos_lo / kPage;
        T.page_hi = (T.pos_hi + kPage - 1) / kPage;
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
                for (int i = len - 1; i >= 0 && n >= 0; i--) {
                    paths[c * (size_t)(d0 + 1) + i] = n;
                    n = parent[n];
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
            std::vector<int> ident(L);
            std::iota(ident.begin(), ident.end(), 0);
            TRY_H(upload(h, ident.data(), ident.size(), &T.leaf_rank));
            TRY_H(upload(h, ident.data(), ident.size(), &T.seed_start));  // filled by rebuild_seed_ranges
            TRY_H(upload(h, ident.data(), ident.size(), &T.seed_end));
            CU_TRY(cudaStreamSynchronize(h->stream));
        }
        TRY_H(rebuild_seed_ranges(h));
        {
            // position -> original index, complete on every device (a merge maps positions of any shard), -1 on padding
            int* oi = nullptr;
            TRY_H(dev_alloc(h, (size_t)T.n_pages * kPage, &oi));
            CU_TRY(cudaMemsetAsync(oi, 0xff, (size_t)T.n_pages * kPage * 4, h->stream));
            CU_TRY(cudaMemcpyAsync(oi, t->sorted_to_orig, (size_t)N * 4, cudaMemcpyHostToDevice, h->stream));
            T.orig_idx = oi;
        }
        // pages of the resident positions: upload row-major slabs, transpose on the device (and keep a padded row-major copy)
        {
            const size_t n_res_pages = (size_t)(T.page_hi - T.page_lo);
            float* pages = nullptr;
            TRY_H(dev_alloc(h, n_res_pages * dpad * kPage, &pages));
            pages -= (size_t)T.page_lo * dpad * kPage;  // global page indices address the resident block
            T.pages = pages;
            // wide points (several 16 KB slices per page): the same padded values row-major, for the exact re-evaluation of single
            // points.  Narrow points are re-evaluated from the stage in shared memory (k_scan) or the L2-resident pages (k_scan_tc)
            float* rows = nullptr;
            if (dpad > kRowsPerStage) {
                TRY_H(dev_alloc(h, n_res_pages * dpad * kPage, &rows));
                rows -= (size_t)T.page_lo * dpad * kPage;
            }
            T.rows = rows;
            const int64_t slab = 4 << 20;  // rows per slab
            const int64_t row_lo = (int64_t)T.page_lo * kPage, row_hi = (int64_t)T.page_hi * kPage;
            DevBuf tmp;
            CU_TRY(tmp.reserve((size_t)std::min<int64_t>(slab, row_hi - row_lo) * dim * 4));
            for (int64_t r0 = row_lo; r0 < row_hi; r0 += slab) {
                int64_t nr = std::min<int64_t>(slab, row_hi - r0);
                int64_t real = std::max<int64_t>(0, std::min<int64_t>(nr, N - r0));
                cudaError_t e = cudaSuccess;
                if (real > 0) e = cudaMemcpyAsync(tmp.p, t->sorted_points + (size_t)r0 * dim, (size_t)real * dim * 4, cudaMemcpyHostToDevice, h->stream);
                if (e == cudaSuccess) {
                    int64_t work = nr * dpad;
                    k_build_pages<<<(unsigned)((work + 255) / 256), 256, 0, h->stream>>>(tmp.as<float>(), r0, nr, dim, dpad, N, pages, rows);
                    e = cudaGetLastError();
                }
                if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
                if (e != cudaSuccess) {
                    tmp.release();
                    CU_TRY(e);
                }
            }
            tmp.release();
        }
        // statistics of the resident points (device): sample mean, max |p_j - mean_j|, max |p - mean|^2, max |p|^2
        float* d_mean = nullptr;
        TRY_H(dev_alloc(h, (size_t)dpad, &d_mean));
        std::vector<float> pmax_h(dpad, 0.f);
        double n2max = 0;
        {
            const int n_res = T.pos_hi - T.pos_lo;
            DevBuf st;
            CU_TRY(st.reserve((size_t)dpad * 4 + 16));
            CU_TRY(cudaMemsetAsync(st.p, 0, (size_t)dpad * 4 + 16, h->stream));
            unsigned long long* d_n2 = reinterpret_cast<unsigned long long*>(st.p);
            unsigned int* d_amax = reinterpret_cast<unsigned int*>(d_n2 + 2);
            k_sample_mean<<<dpad, 256, 0, h->stream>>>(T, std::min(n_res, 65536), d_mean);
            k_shard_maxes<<<(unsigned)((n_res + 255) / 256), 256, dpad * sizeof(unsigned int), h->stream>>>(T, d_mean, d_amax, d_n2, d_n2 + 1);
            std::vector<unsigned int> ab(dpad);
            unsigned long long nb[2] = {0, 0};
            cudaError_t e = cudaGetLastError();
            if (e == cudaSuccess) e = cudaMemcpyAsync(ab.data(), d_amax, (size_t)dpad * 4, cudaMemcpyDeviceToHost, h->stream);
            if (e == cudaSuccess) e = cudaMemcpyAsync(nb, d_n2, 16, cudaMemcpyDeviceToHost, h->stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
            st.release();
            CU_TRY(e);
            for (int j = 0; j < dim; j++) {
                float a;
                memcpy(&a, &ab[j], 4);
                pmax_h[j] = a * 1.0001f;
            }
            double x2;
            memcpy(&n2max, &nb[0], 8);
            memcpy(&x2, &nb[1], 8);
            T.xnorm_max = (float)(std::sqrt(x2) * 1.0001);
            if (!(T.xnorm_max < 3.0e38f)) T.xnorm_max = 3.0e38f;
        }
        TRY_H(upload(h, pmax_h.data(), pmax_h.size(), &h->I.pmax));
        h->I.kp = 0;
        const int kp = (dim + 3 + 15) & ~15;
        if (h->opt.tc_mode != 0 && kp <= 1024) {
            std::vector<float> lm_own;
            const float* lm = leaf_means_in;
            int C = h->opt.groups;
            if (C != 1 && !lm && (T.pos_hi - T.pos_lo) >= 8192) {
                compute_leaf_means(t, lm_own);
                lm = lm_own.data();
            }
            TRY_H(build_scan_image(h, d_mean, pmax_h, n2max, lm));
        }
        CU_TRY(cudaStreamSynchronize(h->stream));
        return SCHT_OK;
    };
    int rc = body();
    if (rc != SCHT_OK) {
        scht_destroy(h);
        return rc;
    }
    *out = h;
    return SCHT_OK;
}

}  // namespace scht
