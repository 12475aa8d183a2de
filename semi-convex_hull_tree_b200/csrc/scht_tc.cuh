// Tensor-core prefilter for the leaf scan (north_star: "Tensor cores are allowed only as a TF32 or BF16 GEMM-form
// prefilter followed by exact FP32 rescoring").
//
// For a tile of 128 queries and a page of 256 points one tcgen05.mma chain (kind::tf32, M=128, N=256, accumulator in
// TMEM) evaluates   S[q][p] = -2 q'.p' + |p'|^2   with q' = q - mean, p' = p - mean (operands rounded to TF32,
// |p'|^2 carried as a hi+lo pair in two extra K slots).  |p-q|^2 = S + |q'|^2 up to a rounding error that is bounded
// per query (margin_q, DESIGN.md section "Tensor-core prefilter").  A pair survives the filter when
//         S <= kth_q * (1 + 2^-20) - |q'|^2 + margin_q
// and every survivor is re-evaluated with the reference's float/double arithmetic (the same exact path as k_scan)
// before it may enter the query's top-K list, so results stay bit-identical; the tensor core only decides which of
// the ~10^11 pairs are worth an exact evaluation (about 1 in 10^4).
//
// Roles: warps 0-15 epilogue, warp 16 = producer lane (cp.async.bulk of page slices already stored in the UMMA
// K-major no-swizzle layout), warps 17/18 = MMA issuers.  The unit of the MMA->epilogue pipeline is a HALF page (128
// points, N=128): the 512 TMEM columns hold four 128-column accumulator buffers.  Group g (MMA warp 17+g and epilogue
// warps 8g..8g+7) owns half page g of every page and the buffers g and g+2, so two double-buffered chains run
// interleaved and hide each other's hand-off latencies (measured: a single issuer and 4 epilogue warps spent half their
// time waiting on each other).  Epilogue warp w reads TMEM lanes 32(w%4)..+31 (thread = query) and 64 of the 128
// columns of its group's buffer.
#pragma once
#include "scht_kernels.cuh"

namespace scht {

#ifdef SCHT_TC_PROF
#define TCP_T0() long long tcp_t0 = clock64()
#define TCP_ADD(var) do { long long t1 = clock64(); var += t1 - tcp_t0; tcp_t0 = t1; } while (0)
#else
#define TCP_T0()
#define TCP_ADD(var)
#endif

constexpr int kTcQ = 128;          // queries per tile (MMA M)
constexpr int kTcStages = 4;
constexpr int kTcKSlice = 32;      // K elements per shared-memory stage (32 KB)
constexpr int kTcEpiWarps = 16;
constexpr int kTcThreads = (kTcEpiWarps + 3) * 32;
constexpr int kTcQueueCap = 384;   // candidate queue entries per epilogue warp (an 8-column quarter chunk adds <= 256)

struct TcScanParams {
    ScanParams S;            // q, q_order, q_leaf, nq, K, write_mode, pos range, ranges (tile = 128 queries), state, outputs
    const float* tcpages;    // [n_pages][kp/4][32][8][4]  tf32-rounded centred coordinates + norm hi/lo
    const float* mean;       // [dpad]
    const float* pmax;       // [dpad] max |p_j - mean_j|
    float pnorm_max;         // max |p - mean|_2
    int kp;                  // padded K (multiple of 8) = ceil8(dim + 2)
    long long* trace;        // debug: [64 pages][8] timestamps of tile 0 (SCHT_TC_PROF builds)
};

__host__ __device__ inline size_t tc_smem_bytes(int dpad, int kp, int K) {
    size_t b = 1024;                                        // alignment slack
    b += (size_t)kTcStages * (kp < kTcKSlice ? kp : kTcKSlice) * kPage * 4;  // B stages (<= 32 KB each)
    b += (size_t)kTcQ * kp * 4;                             // A operand
    b += (size_t)dpad * kTcQ * 4;                           // original queries, dim-major (exact path)
    b += (size_t)kTcQ * K * 8;                              // lists
    b += (size_t)kTcEpiWarps * kTcQueueCap * 8;             // candidate queues
    b += (size_t)kTcQ * 8 + (size_t)kTcQ * 3 * 4;           // seed ranges, a0, a1, qid
    b += 256;                                               // barriers, counters, tmem base
    return b;
}

__device__ __forceinline__ float to_tf32(float x) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
    return __uint_as_float(u);
}

__device__ __forceinline__ uint64_t umma_desc_kmajor(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fff);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;  // Blackwell descriptor version; no swizzle, base offset 0
    return d;
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}
// One lane of a fully active warp (the lowest), as a warp-uniform predicate: keeps the tcgen05 issue code on the
// uniform datapath instead of the per-thread ELECT loop the compiler emits inside a divergent `if (lane == 0)`.
__device__ __forceinline__ bool elect_one_sync() {
    uint32_t pred = 0;
    asm volatile(
        "{\n.reg .b32 rx;\n.reg .pred px;\nelect.sync rx|px, 0xffffffff;\n@px mov.s32 %0, 1;\n}\n"
        : "+r"(pred));
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
// builder of the tensor-core page image from the exact dim-major pages (one thread per sorted position)
// -------------------------------------------------------------------------------------------------------------
__global__ void k_build_tcpages(DevTree T, const float* __restrict__ mean, int kp, float* __restrict__ tcpages) {
    int64_t pos = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (pos >= (int64_t)T.n_pages * kPage) return;
    int page = (int)(pos / kPage), in_page = (int)(pos % kPage);
    const float* src = T.pages + (size_t)page * T.dpad * kPage + in_page;
    float* dst = tcpages + (size_t)page * kp * kPage;
    const int g = in_page >> 3, r = in_page & 7;
    const bool real = pos < T.n_points;
    double n2 = 0.0;
    for (int k = 0; k < kp; k++) {
        float v = 0.f;
        if (k < T.dim) {
            if (real) {
                float p = src[(size_t)k * kPage];
                double c = (double)p - (double)mean[k];
                n2 += c * c;
                v = to_tf32((float)c);
            }
        } else if (k == T.dim) {
            v = real ? to_tf32((float)n2) : 1.0e30f;  // padded positions can never pass the filter
        } else if (k == T.dim + 1) {
            v = real ? to_tf32((float)(n2 - (double)to_tf32((float)n2))) : 0.f;
        }
        dst[(((size_t)(k >> 2) * 32 + g) * 8 + r) * 4 + (k & 3)] = v;
    }
}

// -------------------------------------------------------------------------------------------------------------
// eligibility: the prefilter pays off when its error margin is small against the K-th distance after the seed scan
// counters[6] += queries whose margin is below 1/16 of the seed K-th distance^2
// -------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float tc_margin(const float* qc_abs_sum_pmax, float qnorm, float pnorm_max) {
    // |S_mma + |q'|^2 - |p-q|^2| <= 2(2^-10+2^-20) B + 2^-19 (2B + Nmax),  B >= sum_j |q'_j p'_j|  
    float B = fminf(*qc_abs_sum_pmax, qnorm * pnorm_max);
    return (2.0f * (9.765625e-4f + 9.5367432e-7f) * B + 1.9073486e-6f * (2.0f * B + pnorm_max * pnorm_max)) * 1.01f;
}

__global__ void k_tc_eligibility(DevTree T, const float* __restrict__ q, const int* __restrict__ q_order, int nq, int K,
                                 const uint64_t* __restrict__ state, const float* __restrict__ mean, const float* __restrict__ pmax,
                                 float pnorm_max, unsigned long long* __restrict__ counters) {
    int slot = blockIdx.x * blockDim.x + threadIdx.x;
    bool ok = false;
    if (slot < nq) {
        int qid = q_order[slot];
        float s1 = 0.f, n2 = 0.f;
        for (int j = 0; j < T.dim; j++) {
            float c = q[(size_t)qid * T.dim + j] - mean[j];
            s1 = fmaf(fabsf(c), pmax[j], s1);
            n2 = fmaf(c, c, n2);
        }
        s1 *= 1.0001f;
        float margin = tc_margin(&s1, sqrtf(n2) * 1.0001f, pnorm_max);
        float kth = __uint_as_float((uint32_t)(state[(size_t)slot * K + K - 1] >> 32));
        ok = (kth < 3.0e38f) && (margin * 16.f <= kth);
    }
    unsigned m = __ballot_sync(0xffffffffu, ok);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(&counters[6], (unsigned long long)__popc(m));
}

// -------------------------------------------------------------------------------------------------------------
// the prefilter scan kernel
// -------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kTcThreads, 1) k_scan_tc(const TcScanParams TP) {
    const ScanParams& P = TP.S;
    const DevTree& T = P.T;
    const int dpad = T.dpad, K = P.K, kp = TP.kp;
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* base = reinterpret_cast<unsigned char*>(((uintptr_t)smem_dyn + 1023) & ~(uintptr_t)1023);
    float* stage_mem = reinterpret_cast<float*>(base);                              // [stages][kTcKSlice/4][32][8][4]
    const int stage_k = min(kp, kTcKSlice);                                         // K elements held by one stage
    float* sA = stage_mem + (size_t)kTcStages * stage_k * kPage;                    // [kp/4][16][8][4]
    float* qs = sA + (size_t)kTcQ * kp;                                             // [dpad][128] original queries
    uint64_t* lists = reinterpret_cast<uint64_t*>(qs + (size_t)dpad * kTcQ);        // [128][K]
    uint64_t* queues = lists + (size_t)kTcQ * K;                                    // [16][kTcQueueCap]
    int2* seed_rg = reinterpret_cast<int2*>(queues + (size_t)kTcEpiWarps * kTcQueueCap);  // [128]
    int* a0s = reinterpret_cast<int*>(seed_rg + kTcQ);
    int* a1s = a0s + kTcQ;
    int* qids = a1s + kTcQ;
    uint64_t* full_b = reinterpret_cast<uint64_t*>(qids + kTcQ);
    uint64_t* empty_b = full_b + kTcStages;
    uint64_t* tfull_b = empty_b + kTcStages;   // [4]
    uint64_t* tempty_b = tfull_b + 4;          // [4]
    int* qcount = reinterpret_cast<int*>(tempty_b + 4);  // [16]
    int* qlock = qcount + kTcEpiWarps;                   // [4] one lock per lane quadrant (four warps share its lists)
    int* misc = qlock + 4;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(misc + 2);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile = blockIdx.x;
    const int q_begin = tile * kTcQ;
    const int nq_tile = min(kTcQ, P.nq - q_begin);

    // ---- set-up ---------------------------------------------------------------------------------------------
    if (tid == 0) {
        for (int s = 0; s < kTcStages; s++) {
            mbar_init(&full_b[s], 1);
            mbar_init(&empty_b[s], 2);  // one commit per MMA warp
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
    if (tid < kTcEpiWarps + 4) qcount[tid] = 0;  // queue counters and quadrant locks
    for (int i = tid; i < kTcQ; i += blockDim.x) {
        int qid = (i < nq_tile) ? P.q_order[q_begin + i] : -1;
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
    // queries: exact copy (dim-major) and the A operand: row q = (-2 tf32(q'_j) ..., 1, 1, 0 ...)
    for (int i = tid; i < kTcQ * dpad; i += blockDim.x) {
        int qi = i / dpad, j = i - qi * dpad;
        int qid = qids[qi];
        qs[j * kTcQ + qi] = (qid >= 0 && j < T.dim) ? P.q[(size_t)qid * T.dim + j] : 0.f;
    }
    for (int i = tid; i < kTcQ * kp; i += blockDim.x) {
        int qi = i / kp, k = i - qi * kp;
        int qid = qids[qi];
        float v = 0.f;
        if (qid >= 0) {
            if (k < T.dim) v = -2.0f * to_tf32((float)((double)P.q[(size_t)qid * T.dim + k] - (double)TP.mean[k]));
            else if (k <= T.dim + 1) v = 1.0f;
        }
        sA[(((size_t)(k >> 2) * (kTcQ / 8) + (qi >> 3)) * 8 + (qi & 7)) * 4 + (k & 3)] = v;
    }
    for (int i = tid; i < kTcQ * K; i += blockDim.x) {
        int qi = i / K;
        lists[i] = (qi < nq_tile) ? P.state[(size_t)(q_begin + qi) * K + (i - qi * K)] : kEmptyKey;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // sA was written through the generic proxy
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    // No seed-page exclusion here: the seed scan ran on 32-query tiles, so a page may have been seed-scanned for some of
    // these 128 queries only.  Re-scanning is cheap in this kernel and list insertion is idempotent.
    PageWalk walk(P.ranges + (size_t)tile * kRangeCap, P.n_ranges + (size_t)tile * kMaxCuts, T.n_cuts, kJobRanges, seed_rg, 0);
    const int n_slices = (kp + kTcKSlice - 1) / kTcKSlice;
    int page;

    if (warp == kTcEpiWarps) {
        // ================= producer =================
        if (lane == 0) {
            uint32_t it = 0;
            [[maybe_unused]] long long w_empty = 0, w_other = 0;
            TCP_T0();
            while (walk.next(page)) {
                for (int s = 0; s < n_slices; s++, it++) {
                    int st = it % kTcStages;
                    uint32_t ph = (it / kTcStages) & 1;
                    TCP_ADD(w_other);
                    mbar_wait_backoff(&empty_b[st], ph ^ 1);
                    TCP_ADD(w_empty);
                    int kk = min(kTcKSlice, kp - s * kTcKSlice);
                    uint32_t bytes = (uint32_t)kk * kPage * 4;
                    mbar_arrive_expect_tx(&full_b[st], bytes);
                    bulk_g2s(stage_mem + (size_t)st * stage_k * kPage, TP.tcpages + ((size_t)page * kp + (size_t)s * kTcKSlice) * kPage, bytes,
                             &full_b[st]);
                }
            }
#ifdef SCHT_TC_PROF
            atomicAdd(&P.counters[8], (unsigned long long)w_empty);
            atomicAdd(&P.counters[9], (unsigned long long)w_other);
#endif
        }
    } else if (warp > kTcEpiWarps) {
        // ================= MMA issuers: warp 17+g issues the MMAs of half page g; the whole warp walks the pipeline,
        // one elected lane issues (keeps tcgen05 on the uniform datapath) =================
        {
            const int g = warp - (kTcEpiWarps + 1);
            constexpr int kHalf = kPage / 2;  // N of one MMA: a half page
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kHalf >> 3) << 17) | ((uint32_t)(kTcQ >> 4) << 24);
            // Descriptors are affine in (slice, k8, stage): two bases built once plus 16-byte-unit offsets.
            constexpr uint32_t kAStep = 2 * (kTcQ / 8) * 128 / 16;   // two K chunks of the A image, in 16-byte units
            constexpr uint32_t kBStep = 2 * (kPage / 8) * 128 / 16;  // two K chunks of a B stage
            constexpr uint32_t kBHalf = (kHalf / 8) * 128 / 16;      // second half page inside a K chunk
            const uint32_t kBStage = (uint32_t)stage_k * kPage * 4 / 16;
            const uint64_t da0 = umma_desc_kmajor(smem_u32(sA), (kTcQ / 8) * 128, 128);
            const uint64_t db0 = umma_desc_kmajor(smem_u32(stage_mem), (kPage / 8) * 128, 128) + (uint64_t)((uint32_t)g * kBHalf);
            uint32_t it = 0, pg = 0;
            [[maybe_unused]] long long m_full = 0, m_tempty = 0, m_issue = 0;
            TCP_T0();
            while (walk.next(page)) {
                // half page g of page pg -> buffer 2*(pg&1)+g, used for the (pg>>1)-th time
                const uint32_t buf = 2 * (pg & 1) + g;
                const uint32_t d_tmem = tmem + buf * kHalf;
                TCP_ADD(m_issue);
#ifdef SCHT_TC_PROF
                const bool tr = TP.trace && tile == 0 && g == 0 && lane == 0 && pg >= 2000 && pg < 2064;
                if (tr) TP.trace[(pg - 2000) * 8 + 0] = clock64();
#endif
                mbar_wait(&tempty_b[buf], ((pg >> 1) & 1) ^ 1);
#ifdef SCHT_TC_PROF
                if (tr) TP.trace[(pg - 2000) * 8 + 1] = clock64();
#endif
                TCP_ADD(m_tempty);
                for (int s = 0; s < n_slices; s++, it++) {
                    const uint32_t st = it % kTcStages;
                    const uint32_t ph = (it / kTcStages) & 1;
                    TCP_ADD(m_issue);
                    mbar_wait(&full_b[st], ph);
#ifdef SCHT_TC_PROF
                    if (tr) TP.trace[(pg - 2000) * 8 + 2] = clock64();
#endif
                    TCP_ADD(m_full);
                    tc_fence_after();
                    const int n8 = min(kTcKSlice, kp - s * kTcKSlice) / 8;
                    const uint64_t da_s = da0 + (uint64_t)((uint32_t)s * (kTcKSlice / 8) * kAStep);
                    const uint64_t db_s = db0 + (uint64_t)(st * kBStage);
                    if (elect_one_sync()) {
#pragma unroll 4
                        for (int k8 = 0; k8 < n8; k8++)
                            umma_tf32(d_tmem, da_s + (uint64_t)((uint32_t)k8 * kAStep), db_s + (uint64_t)((uint32_t)k8 * kBStep), idesc,
                                      (s > 0 || k8 > 0) ? 1u : 0u);
                        if (s == n_slices - 1) umma_commit(&tfull_b[buf]);
                        umma_commit(&empty_b[st]);  // the stage may be refilled once both issuers' MMAs have read it
                    }
                    __syncwarp();
#ifdef SCHT_TC_PROF
                    if (tr) TP.trace[(pg - 2000) * 8 + 3] = clock64();
#endif
                }
                pg++;
            }
#ifdef SCHT_TC_PROF
            if (lane == 0 && g == 0) {
                atomicAdd(&P.counters[10], (unsigned long long)m_full);
                atomicAdd(&P.counters[11], (unsigned long long)m_tempty);
                atomicAdd(&P.counters[12], (unsigned long long)m_issue);
            }
#endif
        }
    } else {
        // ================= epilogue: thread = query (TMEM lane), this warp's 128 columns of every page =================
        // quad = TMEM lane quadrant (= warp id % 4, a hardware rule), sub = which 64 columns of the buffer,
        // half = pipeline group = which half page of every page
        const int quad = warp & 3, sub = (warp >> 2) & 1, half = warp >> 3;
        const int qi = quad * 32 + lane;  // tile-local query
        uint64_t* myq = queues + (size_t)warp * kTcQueueCap;
        int* mycount = &qcount[warp];
        int* lock = &qlock[quad];
        const bool valid_q = qi < nq_tile;
        // per-query constants of the filter
        float qn2 = 0.f, s1 = 0.f;
        for (int j = 0; j < T.dim; j++) {
            float c = (float)((double)qs[j * kTcQ + qi] - (double)TP.mean[j]);
            qn2 = fmaf(c, c, qn2);
            s1 = fmaf(fabsf(c), TP.pmax[j], s1);
        }
        s1 *= 1.0001f;
        const float margin = tc_margin(&s1, sqrtf(qn2) * 1.0001f, TP.pnorm_max) + qn2 * 1.0e-5f;  // + slack for qn2's own rounding
        auto bound_of = [&](float kth) {
            // S passes when S <= kth(1+2^-20) - |q'|^2 + margin ; padded query slots never pass
            return valid_q ? fminf(kth * 1.000001f, 3.0e38f) - qn2 + margin : -3.0e38f;
        };
        float Tq = bound_of(__uint_as_float((uint32_t)(lists[(size_t)qi * K + K - 1] >> 32)));
        unsigned long long n_pairs = 0;
        unsigned int n_exact = 0, n_ins = 0;

        // Exact re-evaluation of the queued candidates (reference arithmetic, 32 at a time) and insertion into the lists.
        // The four warps of a lane quadrant share the 32 lists, so the insert phase runs under the quadrant's lock.
        auto drain = [&](int n) {
            __syncwarp();
            for (int b0 = 0; b0 < n; b0 += 32) {
                bool cand = false;
                float d2 = 0.f;
                int pos = 0, cq = 0;
                if (b0 + lane < n) {
                    uint64_t e = myq[b0 + lane];
                    pos = (int)(uint32_t)e;
                    cq = (int)(e >> 32);
                    const int pgi = pos / kPage;
                    const float* pg = T.pages + (size_t)pgi * dpad * kPage + (pos - pgi * kPage);
                    const float* qq = qs + quad * 32 + cq;
                    for (int j = 0; j < T.dim; j++) d2 = ref_acc_sq(d2, pg[(size_t)j * kPage] - qq[j * kTcQ]);
                    n_exact++;
                    cand = (d2 < 3.402823466e+38f) && (pos >= P.pos_lo) && (pos < P.pos_hi);
                    // cheap pre-check against the (possibly stale) list tail, before taking the lock
                    if (cand) cand = ((uint64_t)__float_as_uint(d2) << 32) <= lists[(size_t)(quad * 32 + cq) * K + K - 1];
                }
                unsigned m = __ballot_sync(0xffffffffu, cand);
                if (m) {
                    if (lane == 0) {
                        while (atomicCAS(lock, 0, 1) != 0) __nanosleep(50);
                    }
                    __syncwarp();
                    __threadfence_block();
                    while (m) {
                        int src = __ffs(m) - 1;
                        m &= m - 1;
                        float cd2 = __shfl_sync(0xffffffffu, d2, src);
                        int cpos = __shfl_sync(0xffffffffu, pos, src);
                        int ccq = __shfl_sync(0xffffffffu, cq, src);
                        const int a0 = a0s[quad * 32 + ccq], a1 = a1s[quad * 32 + ccq];
                        uint32_t low = ((cpos >= a0 && cpos < a1) ? 0u : 0x80000000u) | (uint32_t)cpos;
                        uint64_t key = ((uint64_t)__float_as_uint(cd2) << 32) | low;
                        volatile uint64_t* L = lists + (size_t)(quad * 32 + ccq) * K;
                        if (key < L[K - 1]) n_ins += list_insert(const_cast<uint64_t*>(L), K, key, lane);
                    }
                    __threadfence_block();
                    __syncwarp();
                    if (lane == 0) atomicExch(lock, 0);
                }
            }
            __syncwarp();
            if (lane == 0) *mycount = 0;
            Tq = bound_of(__uint_as_float((uint32_t)(reinterpret_cast<volatile uint64_t*>(lists)[(size_t)qi * K + K - 1] >> 32)));
            __syncwarp();
        };

        // Balanced min tree over 32 accumulator values (FMNMX3, depth 4).  Rare path (some lane has a value under its
        // bound): each such lane builds the bit mask of its surviving columns, reserves that many queue slots with ONE
        // shared-memory atomic and writes them; lanes without a hit only wait at the reconvergence point.
        auto process = [&](const uint32_t (&v)[32], int col0) {
            float m[11];
#pragma unroll
            for (int i = 0; i < 10; i++) m[i] = fminf(fminf(__uint_as_float(v[3 * i]), __uint_as_float(v[3 * i + 1])), __uint_as_float(v[3 * i + 2]));
            m[10] = fminf(__uint_as_float(v[30]), __uint_as_float(v[31]));
            float n0 = fminf(fminf(m[0], m[1]), m[2]), n1 = fminf(fminf(m[3], m[4]), m[5]), n2 = fminf(fminf(m[6], m[7]), m[8]),
                  n3 = fminf(m[9], m[10]);
            const float mn = fminf(fminf(fminf(n0, n1), n2), n3);
            const bool hit = mn <= Tq;
            if (__any_sync(0xffffffffu, hit)) {
                uint32_t mask = 0;
                if (hit) {
#pragma unroll
                    for (int c = 0; c < 32; c++) mask |= (__uint_as_float(v[c]) <= Tq ? 1u : 0u) << c;
                }
                for (;;) {
                    bool fits = true;
                    int slot = 0;
                    const int cnt = __popc(mask);
                    if (cnt) {
                        slot = atomicAdd(mycount, cnt);
                        fits = slot + cnt <= kTcQueueCap;
                        if (fits) {
                            while (mask) {
                                const int c = __ffs(mask) - 1;
                                mask &= mask - 1;
                                myq[slot++] = ((uint64_t)(uint32_t)lane << 32) | (uint32_t)(page * kPage + col0 + c);
                            }
                        }
                    }
                    __syncwarp();
                    if (!__any_sync(0xffffffffu, !fits)) break;
                    // queue overflow (only while bounds are still loose): reservations are made in slot order, so the
                    // entries below the first failed reservation are complete; drain them, then the failed lanes retry
                    int first_bad = fits ? 0x7fffffff : slot;
                    for (int o = 16; o; o >>= 1) first_bad = min(first_bad, __shfl_xor_sync(0xffffffffu, first_bad, o));
                    drain(first_bad);
                    if (fits) mask = 0;
                    if (mask) {
                        // the bound may have tightened: drop survivors that no longer pass
                        uint32_t keep = 0;
#pragma unroll
                        for (int c = 0; c < 32; c++) keep |= (__uint_as_float(v[c]) <= Tq ? 1u : 0u) << c;
                        mask &= keep;
                    }
                }
            }
        };

        uint32_t pg = 0;
        [[maybe_unused]] long long e_wait = 0, e_work = 0, e_drain = 0;
        TCP_T0();
        while (walk.next(page)) {
            const uint32_t buf = 2 * (pg & 1) + half;
            TCP_ADD(e_work);
#ifdef SCHT_TC_PROF
            const bool tr = TP.trace && tile == 0 && warp == 0 && lane == 0 && pg >= 2000 && pg < 2064;
            if (tr) TP.trace[(pg - 2000) * 8 + 4] = clock64();
#endif
            mbar_wait(&tfull_b[buf], (pg >> 1) & 1);
#ifdef SCHT_TC_PROF
            if (tr) TP.trace[(pg - 2000) * 8 + 5] = clock64();
#endif
            TCP_ADD(e_wait);
            tc_fence_after();
            const uint32_t taddr = tmem + ((uint32_t)(quad * 32) << 16) + buf * 128 + sub * 64;
            const int col0 = half * 128 + sub * 64;
            // (four epilogue warps per scheduler hide the TMEM-load latency; one 32-column register buffer keeps the
            // kernel inside the 96-register budget of a 608-thread CTA)
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
#ifdef SCHT_TC_PROF
            if (tr) TP.trace[(pg - 2000) * 8 + 6] = clock64();
#endif
            process(va, col0 + 32);
#ifdef SCHT_TC_PROF
            if (tr) TP.trace[(pg - 2000) * 8 + 7] = clock64();
#endif
            if (half == 0 && sub == 0) n_pairs += (unsigned long long)min(kPage, T.n_points - page * kPage);
            TCP_ADD(e_work);
            {
                const int qn = *reinterpret_cast<volatile int*>(mycount);  // warp-uniform: only this warp changes it
                if (qn >= 128) drain(qn);
            }
            TCP_ADD(e_drain);
            pg++;
        }
        drain(*reinterpret_cast<volatile int*>(mycount));
#ifdef SCHT_TC_PROF
        if (lane == 0 && warp == 0) {
            atomicAdd(&P.counters[13], (unsigned long long)e_wait);
            atomicAdd(&P.counters[14], (unsigned long long)e_work);
            atomicAdd(&P.counters[15], (unsigned long long)e_drain);
        }
#endif
        // all four warps of the quadrant must be done before the lists are final
        asm volatile("bar.sync %0, 128;" ::"r"(1 + quad) : "memory");

        // ---- results: the first warp of each quadrant writes its 32 queries ----
        if (half == 0 && sub == 0) {
            for (int qq = 0; qq < 32; qq++) {
                const int ql = quad * 32 + qq;
                if (ql >= nq_tile) break;
                const uint64_t* L = lists + (size_t)ql * K;
                const int qid = qids[ql];
                if (P.write_mode == kWriteFinal) {
                    for (int k = lane; k < K; k += 32) {
                        uint64_t key = L[k];
                        int idx = -1;
                        float d2 = 3.402823466e+38f;
                        if (key != kEmptyKey) {
                            idx = T.orig_idx[(uint32_t)key & 0x7fffffffu];
                            d2 = __uint_as_float((uint32_t)(key >> 32));
                        }
                        P.idx_out[(size_t)qid * K + k] = idx;
                        P.d2_out[(size_t)qid * K + k] = d2;
                        if (P.dist_out) P.dist_out[(size_t)qid * K + k] = __fsqrt_rn(d2);
                    }
                } else if (P.write_mode == kWriteState) {
                    for (int k = lane; k < K; k += 32) P.state[(size_t)(q_begin + ql) * K + k] = L[k];
                } else {
                    for (int k = lane; k < K; k += 32) P.keys_out[(size_t)qid * K + k] = L[k];
                }
            }
        }
        if (P.counters) {
            int nvalid = max(0, min(32, nq_tile - quad * 32));
            for (int o = 16; o; o >>= 1) n_exact += __shfl_xor_sync(0xffffffffu, n_exact, o);
            if (lane == 0) {
                atomicAdd(&P.counters[0], n_pairs * (unsigned long long)nvalid);
                atomicAdd(&P.counters[1], (unsigned long long)n_exact);
                atomicAdd(&P.counters[2], (unsigned long long)n_ins);
            }
        }
    }

    // ---- teardown: everyone is done with TMEM ----
    tc_fence_before();
    __syncthreads();
    if (warp == kTcEpiWarps + 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

}  // namespace scht
