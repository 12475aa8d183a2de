// Tensor-core prefilter for the leaf scan (north_star: "Tensor cores are allowed only as a TF32 or BF16 GEMM-form
// prefilter followed by exact FP32 rescoring").
//
// For a tile of 128 queries and a page of 256 points a tcgen05.mma chain (kind::f16, FP16 operands -- the same 11-bit
// significand as TF32 at half the bytes -- FP32 accumulator in TMEM) evaluates
//     S[q][p] = -2 s^2 q'.p' + s^2 |p'|^2 = s^2 (|p-q|^2 - |q'|^2)
// with q' = q - mean, p' = p - mean and s a power of two that puts the data into FP16's range; the norm rides in three
// extra K slots as an exact hi+mid+lo FP16 split (33 bits) against a constant in the query operand, so the epilogue
// is a pure min tree over the accumulator.  The rounding error of S is bounded per query (margin_q, DESIGN.md
// "Tensor-core prefilter").  A pair survives when
//         S <= s^2 (kth_q (1 + 2^-20) - |q'|^2 + margin_q)
// and every survivor is re-evaluated with the reference's float/double arithmetic (the same exact path as k_scan)
// before it may enter the query's top-K list, so results stay bit-identical; the tensor core only decides which of
// the ~10^11 pairs are worth an exact evaluation (about 3 in 10^4).
//
// Roles (23 or 25 warps): warps 0-15 epilogue, warp 16 = producer lane (cp.async.bulk of page slices already
// stored in the UMMA K-major no-swizzle layout), then 2 or 4 MMA issuer warps (one per page parity or per buffer).  The unit of the
// MMA->epilogue pipeline is a HALF page (128 points, N=128): the 512 TMEM columns hold four 128-column accumulator
// buffers; half page g of page p lives in buffer 2(p&1)+g.  Epilogue warps 8g..8g+7 process the half pages g: warp w
// reads TMEM lanes 32(w%4)..+31 (thread = query) and 64 of the buffer's 128 columns; survivors go into the warp's ring.
// Two issue streams and two double-buffered epilogue groups hide the hand-off latencies (mbarrier waits, MMA issue,
// commits: ~900 cycles per page for a single issuer).  The last four warps are the RESCORERS: rescorer r consumes the rings of
// the four epilogue warps of lane quadrant r, re-evaluates the survivors exactly, owns the quadrant's 32 top-K lists
// (no locks) and publishes the tightened bounds, so the MMA->epilogue pipeline waits for an exact evaluation only when a ring is
// full.  The rings are deliberately SHALLOW for short lists (64 entries per epilogue warp for K <= 32): a survivor that waits goes
// stale while its query's bound tightens, and the back-pressure of a full ring is what keeps the number of exact evaluations down
// (DESIGN.md 4b).  For wide points the entries also carry the pair's accumulator value, and the rescorer repeats the epilogue's
// test against the bound of the moment before it spends a row fetch and an FP64 chain on an entry (TcScanParams::recheck).
#pragma once
#include <cuda_fp16.h>

#include "scht_kernels.cuh"

namespace scht {

#ifdef SCHT_TC_PROF
#define TC_STAMP(cond, pg, slot) do { if ((cond) && tile == 0 && lane == 0 && (pg) >= 2000 && (pg) < 2064) TP.trace[((pg) - 2000) * 8 + (slot)] = clock64(); } while (0)
#else
#define TC_STAMP(cond, pg, slot)
#endif

constexpr int kTcQ = 128;          // queries per tile (MMA M)
constexpr int kTcKSlice = 32;      // K elements (FP16) per shared-memory stage: 16 KB
constexpr int kTcMaxStages = 6;    // slice stages (fewer when K is large and the lists need the shared memory)
constexpr int kTcEpiWarps = 16;
constexpr int kTcMaxMmaWarps = 4;  // MMA issuers: 2 (one per page parity; narrow points, one K slice per page) or 4 (one per TMEM buffer)
__host__ __device__ constexpr int tc_threads(int n_iss) { return (kTcEpiWarps + 1 + n_iss + 4) * 32; }  // + producer, MMA issuers, 4 rescorers
constexpr int kTcQueueCap = 256;   // ring entries per epilogue warp for long lists (K > 32); TcScanParams::queue_cap is what a launch uses
constexpr int kTcQueueCapMax = 1024;
constexpr int kTcTake = 4;         // wide points: ring entries a rescorer lane takes per round (re-checked against the current bound first)
constexpr int kTcPassCap = 32 * kTcTake;
constexpr int kTcSlotBits = 27;    // ring entries that carry the prefilter value name the slot in 27 bits (and the lane in 5)

struct TcScanParams {
    ScanParams S;            // q, q_order, q_leaf, nq, K, write_mode, pos range, ranges (tile = 128 queries), state, outputs
    ScanImage I;             // the scan-order image: FP16 pages, slot -> position, centre / scale of the operands
    int brute;               // 1: keys carry the ORIGINAL index and no approximate-leaf bit (scht_brute_force)
    int single_issuer;       // 1: MMA issuer 0 issues every page (set when a page has at least as many K slices as there are stages: an
                             //    issuer that skips the other parity's pages would run a whole ring round ahead of the producer, and a
                             //    parity wait cannot tell round r from round r+2)
    int n_stages;            // slice stages in shared memory (3..kTcMaxStages)
    int queue_cap;           // ring entries per epilogue warp (power of two, 32..kTcQueueCapMax; the launcher picks 64 for K <= 32, else kTcQueueCap)
    int recheck;             // wide points: ring entries carry the pair's prefilter value and the rescorer drops an entry that the query's
                             //   bound has overtaken while it waited (needs n_slots <= 2^27)
    int n_seg;               // page segments per tile (>= 1): CTA b scans segment b / n_tiles of tile b % n_tiles and, when
    int seg_pages;           //   n_seg > 1, writes its key lists to S.keys_out[segment][query][K] (merged by k_merge_partial);
    int seg_page0;           //   segment s = pages [seg_page0 + s*seg_pages, +seg_pages)
    int rescorer_sleep_ns;   // an idle rescorer's pause between polls of its rings
    int ring_wait_ns;        // an epilogue warp's pause between polls of its rescorer's head while its ring is full
    int debug;               // SCHT_TC_PROF builds only: 1 = no pair passes the filter, 2 = also skip the min tree
    long long* trace;        // SCHT_TC_PROF builds: [64 pages][8] clock64 stamps of tile 0 (pages 2000..2063)
};

__host__ __device__ inline size_t tc_smem_bytes(int dpad, int kp, int K, int n_stages, int queue_cap = kTcQueueCap) {
    size_t b = 1024;                                                         // alignment slack
    b += (size_t)n_stages * (kp < kTcKSlice ? kp : kTcKSlice) * 512;         // B slice stages (<= 16 KB each)
    b += (size_t)kTcQ * kp * 2;                                              // A operand
    b += (size_t)kTcQ * K * 8;                                               // lists
    b += (size_t)kTcEpiWarps * queue_cap * 8;                                // candidate queues
    b += (size_t)4 * kTcPassCap * 8;                                         // rescorers: the entries of a round that passed the re-check
    b += (size_t)kTcQ * 3 * 4;                                               // a0, a1, qid
    b += 1536;                                                               // barriers, ring counters, bounds, tmem base
    return b;
}

__device__ __forceinline__ uint64_t umma_desc_kmajor(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fff);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;  // Blackwell descriptor version; no swizzle, base offset 0
    return d;
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}
// One lane of a fully active warp (the lowest), as a warp-uniform predicate: keeps the tcgen05 issue code on the
// uniform datapath instead of the per-thread ELECT loop the compiler emits inside a divergent `if (lane == 0)`.
__device__ __forceinline__ bool elect_one_sync() {
    uint32_t pred = 0;
    asm volatile("{\n.reg .b32 rx;\n.reg .pred px;\nelect.sync rx|px, 0xffffffff;\n@px mov.s32 %0, 1;\n}\n" : "+r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
          "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
          "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]),
          "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// -------------------------------------------------------------------------------------------------------------
// per-query rounding margin of the filter (unscaled units):
//   |S/s^2 + |q'|^2 - |p-q|^2| <= 2(2^-10+2^-21) B + a(kp) (2B + Nmax) + subnormal term,   B >= sum_j |q'_j p'_j|
// a(kp) covers the FP32 accumulation inside the tensor core (truncating adds, one per K16 step and a short tree inside a
// step): measured <= 1.6 2^-23 sum|a_i b_i| for kp <= 144 (tools/tc_probe.cu), bounded here by max(16, 4 kp/16) 2^-23, i.e.
// it grows with the number of K16 steps instead of trusting the measurement beyond the probed range.
// -------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float tc_margin(float abs_sum_pmax, float qnorm, float q_l1, float pnorm_max, float pabs_sum, float scale, int kp) {
    float B = fminf(abs_sum_pmax, qnorm * pnorm_max);
    const float acc_term = fmaxf(16.f, 4.f * (float)(kp >> 4)) * 1.1920929e-7f;
    float m = 2.0f * (9.765625e-4f + 4.7683716e-7f) * B + acc_term * (2.0f * B + pnorm_max * pnorm_max);
    m += 1.1920929e-7f * (q_l1 + pabs_sum) / scale;  // FP16 subnormals: absolute error 2^-25 per scaled operand
    return m * 1.01f;
}

// eligibility: the prefilter pays off when its error margin is small against the K-th distance after the seed scan.
// counters[6] += queries whose margin is below 1/4 of the seed K-th distance^2 (and which fit FP16's range): even then
// the survivors are a tiny fraction of the pairs, and a prefilter page costs ~1/10 of an exact page
__global__ void k_tc_eligibility(DevTree T, ScanImage I, const float* __restrict__ q, const int* __restrict__ q_order, int nq, int K,
                                 const uint64_t* __restrict__ state, const float* __restrict__ ext_bound, unsigned long long* __restrict__ counters) {
    const float* __restrict__ mean = I.mean;
    const float* __restrict__ pmax = I.pmax;
    const float pnorm_max = I.pnorm_max, pabs_sum = I.pabs_sum, scale = I.scale;
    int slot = blockIdx.x * blockDim.x + threadIdx.x;
    bool ok = false;
    if (slot < nq) {
        int qid = q_order[slot];
        float s1 = 0.f, n2 = 0.f, l1 = 0.f, amax = 0.f;
        for (int j = 0; j < T.dim; j++) {
            float c = q[(size_t)qid * T.dim + j] - mean[j];
            s1 = fmaf(fabsf(c), pmax[j], s1);
            n2 = fmaf(c, c, n2);
            l1 += fabsf(c);
            amax = fmaxf(amax, fabsf(c));
        }
        float margin = tc_margin(s1 * 1.0001f, sqrtf(n2) * 1.0001f, l1 * 1.0001f, pnorm_max, pabs_sum, scale, I.kp);
        float kth = __uint_as_float((uint32_t)(state[(size_t)slot * K + K - 1] >> 32));
        if (ext_bound) kth = fminf(kth, ext_bound[qid]);
        ok = (kth < 3.0e38f) && (margin * 4.f <= kth) && (2.f * amax * scale < 60000.f);
    }
    unsigned m = __ballot_sync(0xffffffffu, ok);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(&counters[6], (unsigned long long)__popc(m));
}

// -------------------------------------------------------------------------------------------------------------
// the prefilter scan kernel
// -------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float fmin3f(float a, float b, float c) {
    float d;
    asm("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}
// asm so that the compiler does not re-fuse pairs of 2-input minima into FMNMX3
__device__ __forceinline__ float fmin2f(float a, float b) {
    float d;
    asm("min.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b));
    return d;
}
// v[c] for a run-time c without indexing the register array (which would move it to local memory): 31 selects
__device__ __forceinline__ uint32_t select32(const uint32_t (&v)[32], int c) {
    const bool b0 = c & 1, b1 = c & 2, b2 = c & 4;
    uint32_t g[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const uint32_t y0 = b0 ? v[8 * i + 1] : v[8 * i], y1 = b0 ? v[8 * i + 3] : v[8 * i + 2];
        const uint32_t y2 = b0 ? v[8 * i + 5] : v[8 * i + 4], y3 = b0 ? v[8 * i + 7] : v[8 * i + 6];
        const uint32_t z0 = b1 ? y1 : y0, z1 = b1 ? y3 : y2;
        g[i] = b2 ? z1 : z0;
    }
    const uint32_t h0 = (c & 8) ? g[1] : g[0], h1 = (c & 8) ? g[3] : g[2];
    return (c & 16) ? h1 : h0;
}
// N_ISS = MMA issuer warps: 2 (23 warps: 80 registers per thread, the min trees keep their values in registers) or 4 (25 warps, 72)
// ROWS = wide points (more than one 16-dimension slice): the rescorers take two ring entries per lane and round instead of one
// (a compile-time choice because the pipeline is sensitive to the kernel's code size and register count: +2 % on 16-dimensional points)
// N256 (two issuers only) = one N=256 MMA chain per page instead of two N=128 chains (128 instead of 2 x 92 cycles per K16
// step): both half-page buffers of the page's parity are filled and committed together; the epilogue is unchanged
template <int N_ISS, bool ROWS, bool N256>
__global__ void __launch_bounds__(tc_threads(N_ISS), 1) k_scan_tc(const TcScanParams TP) {
    static_assert(N_ISS == 2 || N_ISS == 4, "issuers: one per page parity or one per TMEM buffer");
    static_assert(!N256 || N_ISS == 2, "whole-page chains: one issuer per page parity");
    const ScanParams& P = TP.S;
    const DevTree& T = P.T;
    const ScanImage& I = TP.I;
    const int dpad = T.dpad, K = P.K, kp = I.kp;
    // (no manual re-alignment: the no-swizzle UMMA layout and cp.async.bulk need 16 bytes only, and a pointer that is
    // visibly derived from the __shared__ array keeps every access below an LDS/STS/ATOMS instead of a generic LD/ST/ATOM)
    extern __shared__ __align__(128) unsigned char smem_dyn[];
    unsigned char* base = smem_dyn;
    const int stage_k = min(kp, kTcKSlice);                                    // K elements held by one stage
    const uint32_t stage_bytes = (uint32_t)stage_k * 512;
    unsigned char* stage_mem = base;                                           // [stages][stage_k/8][32][8][8] halves
    const int n_stages = TP.n_stages;
    __half* sA = reinterpret_cast<__half*>(stage_mem + (size_t)n_stages * stage_bytes);  // [kp/8][16][8][8]
    uint64_t* lists = reinterpret_cast<uint64_t*>(sA + (size_t)kTcQ * kp);     // [128][K]
    const int qcap = TP.queue_cap;
    uint64_t* queues = lists + (size_t)kTcQ * K;                               // [16][qcap]
    uint64_t* passed = queues + (size_t)kTcEpiWarps * qcap;                    // [4][kTcPassCap]
    int* a0s = reinterpret_cast<int*>(passed + (size_t)4 * kTcPassCap);
    int* a1s = a0s + kTcQ;
    int* qids = a1s + kTcQ;
    uint64_t* full_b = reinterpret_cast<uint64_t*>(qids + kTcQ);
    uint64_t* empty_b = full_b + kTcMaxStages;
    uint64_t* tfull_b = empty_b + kTcMaxStages;   // [4]
    uint64_t* tempty_b = tfull_b + 4;             // [4]
    int* q_tail = reinterpret_cast<int*>(tempty_b + 4);  // [16] entries pushed (written by the epilogue warp)
    int* q_head = q_tail + kTcEpiWarps;                  // [16] entries consumed (written by the rescorer)
    int* q_done = q_head + kTcEpiWarps;                  // [16] the epilogue warp has finished
    int* q_resv = q_done + kTcEpiWarps;                  // [16] ring slots reserved by the epilogue warp's lanes (>= tail)
    float* tq_s = reinterpret_cast<float*>(q_resv + kTcEpiWarps);  // [128] current filter bound of every query
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tq_s + kTcQ);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n_tiles = (P.nq + kTcQ - 1) / kTcQ;
    const int seg = (TP.n_seg > 1) ? (int)blockIdx.x / n_tiles : 0;  // segment-major: CTAs that run together scan the same pages (L2)
    const int tile = (int)blockIdx.x - seg * n_tiles;
    const int q_begin = tile * kTcQ;
    const int nq_tile = min(kTcQ, P.nq - q_begin);
    // Tile slot (= TMEM lane = operand row) i holds the tile's query number ord(i) = 4 (i % 32) + i / 32 in bucket order: queries
    // that are neighbours in bucket order (same approximate leaf, same cluster, hot on the same pages) are dealt round-robin to the
    // four lane quadrants, so a burst of survivors is shared by all four rescorers instead of filling one quadrant's rings
    auto ord = [](int slot) { return ((slot & 31) << 2) | (slot >> 5); };
    const float scale = I.scale;

    // ---- set-up ---------------------------------------------------------------------------------------------
    if (tid == 0) {
        for (int s = 0; s < n_stages; s++) {
            mbar_init(&full_b[s], 1);
            mbar_init(&empty_b[s], N_ISS / 2);  // one commit per issuer that reads the stage
        }
        for (int b = 0; b < 4; b++) {
            mbar_init(&tfull_b[b], 1);
            mbar_init(&tempty_b[b], 8);  // the 8 epilogue warps of the buffer's group
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kTcEpiWarps + 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid < 4 * kTcEpiWarps) q_tail[tid] = 0;  // ring counters, done flags, reservations (contiguous)
    for (int i = tid; i < kTcQ; i += blockDim.x) {
        int qid = (ord(i) < nq_tile) ? P.q_order[q_begin + ord(i)] : -1;
        qids[i] = qid;
        int a0 = 0, a1 = 0;
        if (qid >= 0) {
            int leaf = P.q_leaf[qid];
            a0 = T.leaf_start[leaf];
            a1 = a0 + T.leaf_count[leaf];
        }
        a0s[i] = a0;
        a1s[i] = a1;
    }
    __syncthreads();
    // queries: exact copy (dim-major) and the A operand: row q = fp16(-2 s q'_j); a query outside FP16's range gets a
    // zero row (its bound is set to "everything passes" below, so it stays exact, only slow)
    // (tq_s doubles as the per-query "inside FP16's range" flag until the real bounds are written below)
    for (int qi = tid; qi < kTcQ; qi += blockDim.x) {
        bool in_range = qids[qi] >= 0;
        if (in_range) {
            const float* qrow = P.q + (size_t)qids[qi] * T.dim;
            for (int j = 0; j < T.dim; j++) in_range &= fabsf(qrow[j] - I.mean[j]) * 2.f * scale < 60000.f;
        }
        tq_s[qi] = in_range ? 1.f : 0.f;
    }
    __syncthreads();
    for (int i = tid; i < kTcQ * kp; i += blockDim.x) {
        int qi = i / kp, k = i - qi * kp;
        float v = 0.f;
        if (qids[qi] >= 0 && k < T.dim) {
            const float* qrow = P.q + (size_t)qids[qi] * T.dim;
            if (tq_s[qi] != 0.f) v = (float)(-2.0 * ((double)qrow[k] - (double)I.mean[k]) * (double)scale);
        } else if (k >= T.dim && k < T.dim + 3) {
            v = I.norm_unit;  // multiplies the norm slots of the point operand (every row, also padded ones)
        }
        sA[(((size_t)(k >> 3) * (kTcQ / 8) + (qi >> 3)) * 8 + (qi & 7)) * 8 + (k & 7)] = __float2half_rn(v);
    }
    for (int i = tid; i < kTcQ * K; i += blockDim.x) {
        int qi = i / K;
        lists[i] = (ord(qi) < nq_tile) ? P.state[(size_t)(q_begin + ord(qi)) * K + (i - qi * K)] : kEmptyKey;
    }
    __syncthreads();  // the range flags in tq_s have been read
    for (int i = tid; i < kTcQ; i += blockDim.x) tq_s[i] = -3.0e38f;  // real bounds are published below, before the pipeline starts
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // sA was written through the generic proxy
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    // No seed-page exclusion here: the seed scan ran on 32-query tiles, so a page may have been seed-scanned for some of
    // these 128 queries only.  Re-scanning is cheap in this kernel and list insertion is idempotent.
    const int w_lo = (TP.n_seg > 1) ? TP.seg_page0 + seg * TP.seg_pages : 0;
    const int w_hi = (TP.n_seg > 1) ? ((seg == TP.n_seg - 1) ? 0x7fffffff : w_lo + TP.seg_pages) : 0x7fffffff;
    PageWalk walk(P.ranges + (size_t)tile * kRangeCap, P.n_ranges + (size_t)tile * kMaxCuts, P.range_groups, kJobRanges, nullptr, 0, w_lo, w_hi);
    const int n_slices = (kp + kTcKSlice - 1) / kTcKSlice;
    int page, run_lo, run_hi;

    if (warp == kTcEpiWarps) {
        // ================= producer: the whole warp walks, one elected lane issues the bulk copies =================
        {
            uint32_t st = 0, ph = 0, ppg = 0;  // stage ring position, kept incrementally (a division per page would sit on this warp's critical path)
            const size_t page_bytes = tc_page_bytes(kp);
            while (walk.next_range(run_lo, run_hi)) for (page = run_lo; page < run_hi; page++) {
                const unsigned char* src = I.tcpages + (size_t)page * page_bytes;
                for (int s = 0; s < n_slices; s++) {
                    mbar_wait(&empty_b[st], ph ^ 1);
                    TC_STAMP(s == 0, ppg, 0);
                    const uint32_t bytes = (uint32_t)min(kTcKSlice, kp - s * kTcKSlice) * 512;
                    if (elect_one_sync()) {
                        mbar_arrive_expect_tx(&full_b[st], bytes);
                        bulk_g2s(stage_mem + (size_t)st * stage_bytes, src + (size_t)s * kTcKSlice * 512, bytes, &full_b[st]);
                    }
                    __syncwarp();
                    if (++st == (uint32_t)n_stages) {
                        st = 0;
                        ph ^= 1;
                    }
                }
                TC_STAMP(true, ppg, 1);
                ppg++;
            }
        }
    } else if (warp > kTcEpiWarps && warp <= kTcEpiWarps + N_ISS) {
        // ================= MMA issuers: warp 17+b issues the pages of parity b (half page h -> TMEM buffer 2b+h); the whole warp walks the pipeline,
        // one elected lane issues (keeps tcgen05 on the uniform datapath).  Several issuers, so the hand-off latencies of
        // consecutive pages (mbarrier waits, MMA issue, commits) overlap =================
        // n_iss == 2: warp 17+b issues both half pages of the pages with parity b (3 % faster on one-slice pages);
        // n_iss == 4: warp 17+m issues half page m&1 of the pages with parity m>>1 (7 % faster when a page has several K slices)
        const int m_iss = warp - (kTcEpiWarps + 1);
        const bool single = (N_ISS == 2) && TP.single_issuer;  // issuer 0 takes every page, issuer 1 idles
        const int my_par = (N_ISS == 2) ? m_iss : (m_iss >> 1);
        const int h_lo = (N_ISS == 2) ? 0 : (m_iss & 1), h_hi = (N_ISS == 2) ? 1 : (m_iss & 1);
        constexpr int kHalf = kPage / 2;  // N of one MMA: a half page
        // kind::f16: D = F32 (1<<4), A = B = F16 (format 0), K-major both, N>>3, M>>4
        const uint32_t idesc = (1u << 4) | ((uint32_t)(kHalf >> 3) << 17) | ((uint32_t)(kTcQ >> 4) << 24);
        // Descriptors are affine in (slice, k16, stage, half): two bases built once plus 16-byte-unit offsets.
        constexpr uint32_t kAStep = 2 * (kTcQ / 8) * 128 / 16;   // two K chunks (16 halves) of the A image, in 16-byte units
        constexpr uint32_t kBStep = 2 * (kPage / 8) * 128 / 16;  // two K chunks of a B stage
        constexpr uint32_t kBHalf = (kHalf / 8) * 128 / 16;      // second half page inside a K chunk
        const uint32_t kBStage = stage_bytes / 16;
        const uint64_t da0 = umma_desc_kmajor(smem_u32(sA), (kTcQ / 8) * 128, 128);
        const uint64_t db0 = umma_desc_kmajor(smem_u32(stage_mem), (kPage / 8) * 128, 128);
        uint32_t st = 0, ph = 0, pg = 0;
        auto advance = [&](uint32_t k) {
            st += k;
            while (st >= (uint32_t)n_stages) {
                st -= (uint32_t)n_stages;
                ph ^= 1;
            }
        };
        if (single && m_iss != 0) run_hi = 0;  // (falls through to the teardown)
        else
        while (walk.next_range(run_lo, run_hi)) for (page = run_lo; page < run_hi; page++) {
            const int par = single ? (int)(pg & 1) : my_par;
            if ((int)(pg & 1) != par) {  // the other parity's issuer handles this page
                advance((uint32_t)n_slices);
                pg++;
                continue;
            }
            const uint32_t d_tmem = tmem + (uint32_t)par * kPage;  // buffers 2*par (half 0) and 2*par+1 (half 1)
            uint64_t* const tfull_p = &tfull_b[2 * par];
            uint64_t* const tempty_p = &tempty_b[2 * par];
            const uint32_t use_ph = ((pg >> 1) & 1) ^ 1;  // parity to wait for on tempty: the buffer pair's (pg/2)-th use (the first passes immediately)
            for (int s = 0; s < n_slices; s++) {
                mbar_wait(&full_b[st], ph);
                TC_STAMP(s == 0, pg, 3);
                const int n16 = min(kTcKSlice, kp - s * kTcKSlice) / 16;
                const uint64_t da_s = da0 + (uint64_t)((uint32_t)s * (kTcKSlice / 16) * kAStep);
                const uint64_t db_s = db0 + (uint64_t)(st * kBStage);
                if constexpr (N256) {
                    if (s == 0) {  // both halves' epilogue warps have read the previous page of this parity
                        mbar_wait(&tempty_p[0], use_ph);
                        mbar_wait(&tempty_p[1], use_ph);
                    }
                    TC_STAMP(s == 0, pg, 2);
                    tc_fence_after();
                    if (elect_one_sync()) {
                        const uint32_t idesc256 = (1u << 4) | ((uint32_t)(kPage >> 3) << 17) | ((uint32_t)(kTcQ >> 4) << 24);
#pragma unroll 2
                        for (int k16 = 0; k16 < n16; k16++)
                            umma_f16(d_tmem, da_s + (uint64_t)((uint32_t)k16 * kAStep), db_s + (uint64_t)((uint32_t)k16 * kBStep), idesc256,
                                     (s > 0 || k16 > 0) ? 1u : 0u);
                        if (s == n_slices - 1) {
                            umma_commit(&tfull_p[0]);
                            umma_commit(&tfull_p[1]);
                        }
                        umma_commit(&empty_b[st]);
                    }
                    __syncwarp();
                } else
                for (int h = h_lo; h <= h_hi; h++) {
                    if (s == 0) mbar_wait(&tempty_p[h], use_ph);  // the half's epilogue warps have read its previous page
                    TC_STAMP(s == 0 && h == 0, pg, 2);
                    tc_fence_after();
                    if (elect_one_sync()) {
#pragma unroll 2
                        for (int k16 = 0; k16 < n16; k16++)
                            umma_f16(d_tmem + h * kHalf, da_s + (uint64_t)((uint32_t)k16 * kAStep),
                                     db_s + (uint64_t)((uint32_t)k16 * kBStep + (uint32_t)h * kBHalf), idesc, (s > 0 || k16 > 0) ? 1u : 0u);
                        if (s == n_slices - 1) umma_commit(&tfull_p[h]);
                        if (h == h_hi) umma_commit(&empty_b[st]);  // the stage may be refilled once both half pages' MMAs have read it
                    }
                    __syncwarp();
                }
                advance(1);
            }
            TC_STAMP(true, pg, 4);
            pg++;
        }
    } else {
        // epilogue warps and rescorers share the per-query filter constants: thread = query of the lane quadrant
        const bool is_epi = warp < kTcEpiWarps;
        const int quad = is_epi ? (warp & 3) : (warp - (kTcEpiWarps + 1 + N_ISS));
        const int qi = quad * 32 + lane;  // tile-local query
        const bool valid_q = ord(qi) < nq_tile;
        float qn2 = 0.f, s1 = 0.f, l1 = 0.f, amax = 0.f;
        const float* my_qrow = P.q + (size_t)(valid_q ? qids[qi] : 0) * T.dim;
        for (int j = 0; j < T.dim; j++) {
            float c = valid_q ? (float)((double)my_qrow[j] - (double)I.mean[j]) : 0.f;
            qn2 = fmaf(c, c, qn2);
            s1 = fmaf(fabsf(c), I.pmax[j], s1);
            l1 += fabsf(c);
            amax = fmaxf(amax, fabsf(c));
        }
        const bool in_range = 2.f * amax * scale < 60000.f;
        const float margin = tc_margin(s1 * 1.0001f, sqrtf(qn2) * 1.0001f, l1 * 1.0001f, I.pnorm_max, I.pabs_sum, scale, kp) + qn2 * 1.0e-5f;
        const float s2 = scale * scale;
        // external upper bound of the K-th distance^2 (seed scan of another shard / of the tree for the brute-force entry)
        const float ext_kth = (P.ext_bound && valid_q) ? P.ext_bound[qids[qi]] : 3.402823466e+38f;
        auto bound_of = [&](float kth) {
            kth = fminf(kth, ext_kth);
            // S passes when S <= s^2 (kth(1+2^-20) - |q'|^2 + margin); padded query slots never pass; a query outside
            // FP16's range has a zero operand row and lets everything pass.  The bound is NOT clamped below the norm slots
            // of padded positions (60000 u): a real pair of a far query can lie above any such clamp, and a padded position
            // that passes (only the last page has them) is rejected by the exact evaluation (its distance is +inf).
            if (!valid_q) return -3.0e38f;
            if (!in_range) return 3.0e38f;
            return (fminf(kth * 1.000001f, 3.0e38f) - qn2 + margin) * s2;
        };

        if (!is_epi) {
            // ================= rescorer of lane quadrant `quad` =================
            // Consumes the rings of the quadrant's four epilogue warps (single producer / single consumer each), re-evaluates
            // the survivors with the reference arithmetic, inserts into the quadrant's lists (it is their only writer) and
            // publishes every query's tightened bound in tq_s.
            volatile int* v_tail = q_tail;
            volatile int* v_head = q_head;
            volatile int* v_done = q_done;
            volatile float* v_tq = tq_s;
            unsigned int n_exact = 0, n_ins = 0, rot = 0;
            v_tq[qi] = bound_of(__uint_as_float((uint32_t)(lists[(size_t)qi * K + K - 1] >> 32)));  // bounds from the seed scan
            __threadfence_block();
            asm volatile("bar.sync %0, 160;" ::"r"(1 + quad) : "memory");  // ... visible to the quadrant's 4 epilogue warps
            for (;;) {
                // One ROUND takes up to 32 kE entries from the quadrant's four rings together, kE per lane (entry x of the round ->
                // ring by prefix sums of the available counts): the latency of an exact evaluation (slot -> position lookup, row
                // loads, the sequential FP64 chain) is paid once per round whatever the number of entries, and a lane's two chains
                // are independent, so they overlap.  Clustered data makes the rescorers the bottleneck of the kernel (a hundred
                // or more survivors per query against ~20 on uniform data): 32 entries per round left them latency-bound.
                // Fair when the rings are busy (large K): every ring gets 16 of the 64 places first, the rest goes to whoever has
                // more, starting at a ring that rotates from round to round (a starved ring stalls its epilogue warp, hence the MMA).
                constexpr int kE = ROWS ? 2 : 1;  // (narrow points, i.e. the uniform 16-D benchmark, have ~20 survivors per query: one entry per lane, 2 % faster there)
                constexpr int kTake = ROWS ? kTcTake : 1;  // ring entries per lane and round
                bool all_done = true;
                int head_r[4], n_r[4], avail_r[4], total = 0;
#pragma unroll
                for (int i = 0; i < 4; i++) {  // slot i = ring (i + rot) % 4
                    const int w = quad + 4 * ((i + rot) & 3);  // epilogue warp ids of this quadrant: quad, quad+4, quad+8, quad+12
                    const int done = v_done[w];  // (read before the tail: the warp publishes its last tail first)
                    const int tail = v_tail[w];
                    head_r[i] = v_head[w];
                    avail_r[i] = tail - head_r[i];
                    all_done &= (done != 0) && (avail_r[i] == 0);
                    n_r[i] = min(avail_r[i], 8 * kTake);
                    total += n_r[i];
                }
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const int extra = min(avail_r[i] - n_r[i], 32 * kTake - total);
                    n_r[i] += extra;
                    total += extra;
                }
                if (total == 0) {
                    if (all_done) break;
                    __nanosleep((unsigned)TP.rescorer_sleep_ns);  // survivors are rare on uniform data (~1 per 10 pages and warp): poll seldom, the pipeline warps need the issue slots
                    continue;
                }
                if constexpr (ROWS) {
                    // ---- wide points: take up to 4 entries per lane, keep what the queries' CURRENT bounds still admit, evaluate ----
                    // A survivor waits in its ring while its query's bound tightens; on clustered data a third of the entries are
                    // stale by the time they are taken (147 evaluations per query with 256-entry rings against 120 with 32-entry
                    // ones).  With `recheck` an entry carries the pair's prefilter value, so the rescorer repeats the epilogue's
                    // test with the bound of NOW -- one shared-memory read -- before it spends a row fetch and an FP64 chain on it.
                    // What is left is compacted through shared memory, so the evaluation rounds below are full.
                    __threadfence_block();  // entries below a `tail` were written before it was published
                    const bool recheck = TP.recheck != 0;
                    uint64_t* mine = passed + (size_t)quad * kTcPassCap;
                    int n_pass = 0;
#pragma unroll
                    for (int e = 0; e < kTake; e++) {
                        int ri = 0, off = lane + 32 * e;
                        bool ok = off < total;
#pragma unroll
                        for (int i = 0; i < 3; i++)
                            if (ri == i && off >= n_r[i]) {
                                off -= n_r[i];
                                ri = i + 1;
                            }
                        const int head = ri == 0 ? head_r[0] : ri == 1 ? head_r[1] : ri == 2 ? head_r[2] : head_r[3];
                        const uint64_t* ring = queues + (size_t)(quad + 4 * ((ri + rot) & 3)) * qcap;
                        uint64_t en = 0;
                        if (ok) {
                            en = ring[(head + off) & (qcap - 1)];
                            if (recheck) {
                                const uint32_t lq = ((uint32_t)en >> kTcSlotBits) & 31u;
                                ok = __uint_as_float((uint32_t)(en >> 32)) <= v_tq[quad * 32 + lq];
                                en = ((uint64_t)lq << 32) | ((uint32_t)en & ((1u << kTcSlotBits) - 1u));
                            }
                        }
                        const unsigned m = __ballot_sync(0xffffffffu, ok);
                        if (ok) mine[n_pass + __popc(m & ((1u << lane) - 1u))] = en;
                        n_pass += __popc(m);
                    }
                    __syncwarp();
                    if (lane < 4) {  // the ring slots may be reused
                        const int nh = lane == 0 ? head_r[0] + n_r[0] : lane == 1 ? head_r[1] + n_r[1] : lane == 2 ? head_r[2] + n_r[2] : head_r[3] + n_r[3];
                        v_head[quad + 4 * ((lane + rot) & 3)] = nh;
                    }
                    rot++;
                    __syncwarp();  // ... and every lane reads the new heads in the next round; `mine` is complete
                    for (int c0 = 0; c0 < n_pass; c0 += 32 * kE) {
                        bool live[kE], cand[kE];
                        float d2[kE];
                        int pos[kE], cq[kE];
                        const float* qq[kE];
                        const float4* pr[kE];
#pragma unroll
                        for (int e = 0; e < kE; e++) {
                            const int x = c0 + lane + 32 * e;
                            live[e] = x < n_pass;
                            cand[e] = false;
                            d2[e] = 0.f;
                            pos[e] = -1;
                            cq[e] = 0;
                            qq[e] = P.q;
                            pr[e] = reinterpret_cast<const float4*>(I.srows);
                            if (live[e]) {
                                const uint64_t en = mine[x];
                                const uint32_t slot = (uint32_t)en;  // a SLOT of the scan-order image
                                // the survivor's exact coordinates: row-major in scan order, so the row loads do not wait for the
                                // slot -> position lookup (both go out together; the position is only needed for the key)
                                pr[e] = reinterpret_cast<const float4*>(I.srows + (size_t)slot * dpad);
                                prefetch_row(pr[e], dpad);
                                pos[e] = I.slot_pos[slot];  // -1: padding
                                cq[e] = (int)(en >> 32);
                                qq[e] = P.q + (size_t)qids[quad * 32 + cq[e]] * T.dim;  // the query's original coordinates
                                n_exact++;
                            }
                        }
                        // loads first (8 dims of every entry), then the sequential reference chains side by side
                        for (int j0 = 0; j0 < T.dim; j0 += 8) {
                            float pv[kE][8], qv[kE][8];
#pragma unroll
                            for (int e = 0; e < kE; e++) {
                                const float4 x = live[e] ? __ldg(pr[e] + (j0 >> 2)) : make_float4(0.f, 0.f, 0.f, 0.f);
                                const float4 y = (live[e] && j0 + 4 < dpad) ? __ldg(pr[e] + (j0 >> 2) + 1) : make_float4(0.f, 0.f, 0.f, 0.f);
                                pv[e][0] = x.x, pv[e][1] = x.y, pv[e][2] = x.z, pv[e][3] = x.w;
                                pv[e][4] = y.x, pv[e][5] = y.y, pv[e][6] = y.z, pv[e][7] = y.w;
#pragma unroll
                                for (int u = 0; u < 8; u++) qv[e][u] = (j0 + u < T.dim) ? qq[e][j0 + u] : 0.f;
                            }
#pragma unroll
                            for (int u = 0; u < 8; u++)
                                if (j0 + u < T.dim) {
#pragma unroll
                                    for (int e = 0; e < kE; e++) d2[e] = ref_acc_sq(d2[e], pv[e][u] - qv[e][u]);
                                }
                        }
#pragma unroll
                        for (int e = 0; e < kE; e++) {
                            cand[e] = live[e] && pos[e] >= 0 && (d2[e] < 3.402823466e+38f) && (pos[e] >= P.pos_lo) && (pos[e] < P.pos_hi);
                            if (cand[e]) cand[e] = ((uint64_t)__float_as_uint(d2[e]) << 32) <= lists[(size_t)(quad * 32 + cq[e]) * K + K - 1];
                        }
                        __syncwarp();
#pragma unroll
                        for (int e = 0; e < kE; e++) {
                            unsigned m = __ballot_sync(0xffffffffu, cand[e]);
                            while (m) {
                                int src = __ffs(m) - 1;
                                m &= m - 1;
                                float cd2 = __shfl_sync(0xffffffffu, d2[e], src);
                                int cpos = __shfl_sync(0xffffffffu, pos[e], src);
                                int ccq = __shfl_sync(0xffffffffu, cq[e], src);
                                const int a0 = a0s[quad * 32 + ccq], a1 = a1s[quad * 32 + ccq];
                                uint32_t low = TP.brute ? (uint32_t)T.orig_idx[cpos] : (((cpos >= a0 && cpos < a1) ? 0u : 0x80000000u) | (uint32_t)cpos);
                                uint64_t key = ((uint64_t)__float_as_uint(cd2) << 32) | low;
                                uint64_t* L = lists + (size_t)(quad * 32 + ccq) * K;
                                if (key < L[K - 1]) n_ins += list_insert(L, K, key, lane);
                            }
                        }
                        // publish the (possibly) tightened bound of this lane's query
                        v_tq[qi] = bound_of(__uint_as_float((uint32_t)(lists[(size_t)qi * K + K - 1] >> 32)));
                    }
                } else
                {
                    __threadfence_block();  // entries below a `tail` were written before it was published
                    bool live[kE], cand[kE];
                    float d2[kE];
                    int pos[kE], cq[kE];
                    const float* qq[kE];
                    const float4* pr[kE];
#pragma unroll
                    for (int e = 0; e < kE; e++) {
                        int ri = 0, off = lane + 32 * e;
                        live[e] = off < total;
#pragma unroll
                        for (int i = 0; i < 3; i++)
                            if (ri == i && off >= n_r[i]) {
                                off -= n_r[i];
                                ri = i + 1;
                            }
                        const int head = ri == 0 ? head_r[0] : ri == 1 ? head_r[1] : ri == 2 ? head_r[2] : head_r[3];
                        const uint64_t* ring = queues + (size_t)(quad + 4 * ((ri + rot) & 3)) * qcap;
                        cand[e] = false;
                        d2[e] = 0.f;
                        pos[e] = -1;
                        cq[e] = 0;
                        qq[e] = P.q;
                        pr[e] = reinterpret_cast<const float4*>(I.srows);
                        if (live[e]) {
                            const uint64_t en = ring[(head + off) & (qcap - 1)];
                            const uint32_t slot = (uint32_t)en;  // ring entries name a SLOT of the scan-order image
                            // the survivor's exact coordinates: row-major in scan order, so the row loads do not wait for the
                            // slot -> position lookup (both go out together; the position is only needed for the key)
                            pr[e] = reinterpret_cast<const float4*>(I.srows + (size_t)slot * dpad);
                            prefetch_row(pr[e], dpad);
                            pos[e] = I.slot_pos[slot];  // -1: padding
                            cq[e] = (int)(en >> 32);
                            qq[e] = P.q + (size_t)qids[quad * 32 + cq[e]] * T.dim;  // the query's original coordinates
                            n_exact++;
                        }
                    }
                    // loads first (8 dims of every entry), then the sequential reference chains side by side
                    for (int j0 = 0; j0 < T.dim; j0 += 8) {
                        float pv[kE][8], qv[kE][8];
#pragma unroll
                        for (int e = 0; e < kE; e++) {
                            const float4 x = live[e] ? __ldg(pr[e] + (j0 >> 2)) : make_float4(0.f, 0.f, 0.f, 0.f);
                            const float4 y = (live[e] && j0 + 4 < dpad) ? __ldg(pr[e] + (j0 >> 2) + 1) : make_float4(0.f, 0.f, 0.f, 0.f);
                            pv[e][0] = x.x, pv[e][1] = x.y, pv[e][2] = x.z, pv[e][3] = x.w;
                            pv[e][4] = y.x, pv[e][5] = y.y, pv[e][6] = y.z, pv[e][7] = y.w;
#pragma unroll
                            for (int u = 0; u < 8; u++) qv[e][u] = (j0 + u < T.dim) ? qq[e][j0 + u] : 0.f;
                        }
#pragma unroll
                        for (int u = 0; u < 8; u++)
                            if (j0 + u < T.dim) {
#pragma unroll
                                for (int e = 0; e < kE; e++) d2[e] = ref_acc_sq(d2[e], pv[e][u] - qv[e][u]);
                            }
                    }
#pragma unroll
                    for (int e = 0; e < kE; e++) {
                        cand[e] = live[e] && pos[e] >= 0 && (d2[e] < 3.402823466e+38f) && (pos[e] >= P.pos_lo) && (pos[e] < P.pos_hi);
                        if (cand[e]) cand[e] = ((uint64_t)__float_as_uint(d2[e]) << 32) <= lists[(size_t)(quad * 32 + cq[e]) * K + K - 1];
                    }
                    __syncwarp();
                    if (lane < 4) {  // the ring slots may be reused (values are in registers now)
                        const int nh = lane == 0 ? head_r[0] + n_r[0] : lane == 1 ? head_r[1] + n_r[1] : lane == 2 ? head_r[2] + n_r[2] : head_r[3] + n_r[3];
                        v_head[quad + 4 * ((lane + rot) & 3)] = nh;
                    }
                    rot++;
                    __syncwarp();  // ... and every lane reads the new heads in the next round
#pragma unroll
                    for (int e = 0; e < kE; e++) {
                        unsigned m = __ballot_sync(0xffffffffu, cand[e]);
                        while (m) {
                            int src = __ffs(m) - 1;
                            m &= m - 1;
                            float cd2 = __shfl_sync(0xffffffffu, d2[e], src);
                            int cpos = __shfl_sync(0xffffffffu, pos[e], src);
                            int ccq = __shfl_sync(0xffffffffu, cq[e], src);
                            const int a0 = a0s[quad * 32 + ccq], a1 = a1s[quad * 32 + ccq];
                            uint32_t low = TP.brute ? (uint32_t)T.orig_idx[cpos] : (((cpos >= a0 && cpos < a1) ? 0u : 0x80000000u) | (uint32_t)cpos);
                            uint64_t key = ((uint64_t)__float_as_uint(cd2) << 32) | low;
                            uint64_t* L = lists + (size_t)(quad * 32 + ccq) * K;
                            if (key < L[K - 1]) n_ins += list_insert(L, K, key, lane);
                        }
                    }
                    // publish the (possibly) tightened bound of this lane's query
                    v_tq[qi] = bound_of(__uint_as_float((uint32_t)(lists[(size_t)qi * K + K - 1] >> 32)));
                }
            }
            __syncwarp();
            // ---- results: the rescorer owns the lists of its 32 queries ----
            for (int qq = 0; qq < 32; qq++) {
                const int ql = quad * 32 + qq;
                if (ord(ql) >= nq_tile) continue;
                const uint64_t* L = lists + (size_t)ql * K;
                const int qid = qids[ql];
                if (TP.n_seg > 1) {  // a page segment of a split tile: partial lists, merged afterwards
                    uint64_t* dst = P.keys_out + ((size_t)seg * P.nq + qid) * K;
                    for (int k = lane; k < K; k += 32) dst[k] = L[k];
                } else if (P.write_mode == kWriteFinal) {
                    for (int k = lane; k < K; k += 32) {
                        uint64_t key = L[k];
                        int idx = -1;
                        float d2 = 3.402823466e+38f;
                        if (key != kEmptyKey) {
                            idx = TP.brute ? (int)(uint32_t)key : T.orig_idx[(uint32_t)key & 0x7fffffffu];
                            d2 = __uint_as_float((uint32_t)(key >> 32));
                        }
                        P.idx_out[(size_t)qid * K + k] = idx;
                        P.d2_out[(size_t)qid * K + k] = d2;
                        if (P.dist_out) P.dist_out[(size_t)qid * K + k] = __fsqrt_rn(d2);
                    }
                } else if (P.write_mode == kWriteState) {
                    for (int k = lane; k < K; k += 32) P.state[(size_t)(q_begin + ord(ql)) * K + k] = L[k];
                } else {
                    uint64_t* dst = keys_dst(P, qid);  // plain key array, or the owner device's gather buffer (peer memory)
                    for (int k = lane; k < K; k += 32) dst[k] = L[k];
                }
            }
            if (P.counters) {
                for (int o = 16; o; o >>= 1) n_exact += __shfl_xor_sync(0xffffffffu, n_exact, o);
                if (lane == 0) {
                    atomicAdd(&P.counters[1], (unsigned long long)n_exact);
                    atomicAdd(&P.counters[2], (unsigned long long)n_ins);
                }
            }
        } else {
            // ================= epilogue: thread = query (TMEM lane) =================
            // quad = TMEM lane quadrant (= warp id % 4, a hardware rule), sub = which 64 columns of the buffer,
            // half = pipeline group = which half page of every page
            const int sub = (warp >> 2) & 1, half = warp >> 3;
            uint64_t* myq = queues + (size_t)warp * qcap;
            volatile int* v_head = q_head + warp;
            volatile float* v_tq = tq_s + qi;
            int resv = 0;  // ring slots written by this warp so far (warp-uniform)
            int head_cache = 0;  // last value of the rescorer's head this warp has seen (warp-uniform; only re-read when the ring looks full)
            asm volatile("bar.sync %0, 160;" ::"r"(1 + quad) : "memory");  // the rescorer has published the initial bounds
            float Tq = *v_tq;

            // 32 columns: min tree over four groups of 8.  Rare path (some lane has a value under its bound): the lane
            // finds its surviving columns through the group minima, reserves ring slots with one shared-memory
            // atomic, writes them, and the warp publishes the new tail for the quadrant's rescorer.
            auto process = [&](uint32_t (&v)[32], int col0) {
#ifdef SCHT_TC_PROF
                if (TP.debug == 2) return;
#endif
                // four groups of 8: 2 FMNMX3 (ALU pipe, half rate) + 3 FMNMX (FMA pipe, full rate) each, so both pipes work
                float m[4];
#pragma unroll
                for (int g = 0; g < 4; g++) {
                    const float x = fmin3f(__uint_as_float(v[8 * g]), __uint_as_float(v[8 * g + 1]), __uint_as_float(v[8 * g + 2]));
                    const float y = fmin3f(__uint_as_float(v[8 * g + 3]), __uint_as_float(v[8 * g + 4]), __uint_as_float(v[8 * g + 5]));
                    const float z = fmin2f(__uint_as_float(v[8 * g + 6]), __uint_as_float(v[8 * g + 7]));
                    m[g] = fmin2f(fmin2f(x, y), z);
                }
                const float mn = fmin2f(fmin2f(m[0], m[1]), fmin2f(m[2], m[3]));
                const bool hit = mn <= Tq;
                if (__any_sync(0xffffffffu, hit)) {
                    uint32_t mask = 0;
                    if (hit) {
#pragma unroll
                        for (int g = 0; g < 4; g++) {
                            if (m[g] <= Tq) {  // rare: look at the group's columns
#pragma unroll
                                for (int u = 0; u < 8; u++)
                                    if (__uint_as_float(v[8 * g + u]) <= Tq) mask |= 1u << (8 * g + u);
                            }
                        }
                    }
                    // Ballot rounds: in every round each lane that still has a survivor pushes one; its ring slot is its rank among
                    // those lanes (no atomic: the warp is the ring's only producer and `resv` is warp-uniform).  A handful of survivors
                    // take one or two rounds.  Only published entries are consumed, so before the warp waits for room it publishes
                    // what it has written (a full ring can always drain).  On clustered data nearly every 32-column group of a cluster
                    // page has a survivor, and this branch -- it used to hold a shared-memory atomic with its round trip, a second
                    // warp-wide reduction and two more warp barriers -- set the pace of the whole kernel there.
                    for (;;) {
                        const unsigned b = __ballot_sync(0xffffffffu, mask != 0);
                        if (!b) break;
                        const int n = __popc(b);
                        if (resv + n - head_cache > qcap) {
                            __syncwarp();
                            __threadfence_block();
                            if (lane == 0) *reinterpret_cast<volatile int*>(q_tail + warp) = resv;
                            for (;;) {
                                head_cache = __shfl_sync(0xffffffffu, *v_head, 0);
                                if (resv + n - head_cache <= qcap) break;
                                __nanosleep((unsigned)TP.ring_wait_ns);
                            }
                        }
                        if (ROWS && TP.recheck) {  // (warp-uniform) the entry carries the pair's prefilter value
                            const int c = mask ? __ffs(mask) - 1 : 0;
                            const uint32_t val = select32(v, c);
                            if (mask) {
                                mask &= mask - 1;
                                const int slot = resv + __popc(b & ((1u << lane) - 1u));
                                myq[slot & (qcap - 1)] = ((uint64_t)val << 32) | ((uint32_t)lane << kTcSlotBits) | (uint32_t)(page * kPage + col0 + c);
                            }
                        } else if (mask) {
                            const int c = __ffs(mask) - 1;
                            mask &= mask - 1;
                            const int slot = resv + __popc(b & ((1u << lane) - 1u));
                            myq[slot & (qcap - 1)] = ((uint64_t)(uint32_t)lane << 32) | (uint32_t)(page * kPage + col0 + c);
                        }
                        resv += n;
                    }
                    __syncwarp();
                    __threadfence_block();
                    if (lane == 0) *reinterpret_cast<volatile int*>(q_tail + warp) = resv;  // publish for the rescorer
                }
            };

            uint32_t pg = 0;
            while (walk.next_range(run_lo, run_hi)) for (page = run_lo; page < run_hi; page++) {
                const uint32_t buf = 2 * (pg & 1) + half;
                mbar_wait(&tfull_b[buf], (pg >> 1) & 1);
                TC_STAMP(warp == 0, pg, 5);
                tc_fence_after();
                const uint32_t taddr = tmem + ((uint32_t)(quad * 32) << 16) + buf * 128 + sub * 64;
                const int col0 = half * 128 + sub * 64;
                Tq = *v_tq;  // bound as last published by the rescorer (only ever tightens)
                uint32_t va[32];
                tmem_ld32(taddr, va);
                tmem_ld_wait();
                process(va, col0);
                tmem_ld32(taddr + 32, va);
                tmem_ld_wait();
                // all of this warp's TMEM reads of the buffer are complete: hand it back to the MMA issuer
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty_b[buf]);
                TC_STAMP(warp == 0, pg, 6);
                process(va, col0 + 32);
                TC_STAMP(warp == 0, pg, 7);
                pg++;
            }
            __syncwarp();
            __threadfence_block();
            if (lane == 0) *reinterpret_cast<volatile int*>(q_done + warp) = 1;
            if (P.counters && half == 0 && sub == 0 && lane == 0) {
                int nvalid = max(0, min(32, (nq_tile - quad + 3) / 4));  // slots of the quadrant with ord(slot) < nq_tile
                atomicAdd(&P.counters[0], (unsigned long long)pg * kPage * (unsigned long long)nvalid);  // every scan page is 256 slots (padding included)
            }
        }
    }

    // ---- teardown: everyone is done with TMEM ----
    tc_fence_before();
    __syncthreads();
    if (warp == kTcEpiWarps + 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

}  // namespace scht
