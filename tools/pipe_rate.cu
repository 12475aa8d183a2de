// Micro-benchmark: page rate of the prefilter pipeline SKELETON (producer -> tcgen05.mma -> TMEM -> epilogue min tree) for
// different structures, without the survivor path: how many cycles per 128-query x 256-point page can one SM sustain?
//   NBUF     TMEM accumulator buffers (512/NBUF columns each = N of one MMA chain)
//   SPLIT    warp groups (4 warps = 4 lane quadrants) sharing the columns of one buffer fill
//   GROUPS   epilogue warp groups in total; TEAMS = GROUPS/SPLIT teams take buffer fills round-robin
//   ISS/PROD issuer / producer warps
//   TREE     0: epilogue only reads TMEM; 1: FMNMX3 tree; 2: FMNMX3 + FMNMX mix (ALU pipe + FMA pipe)
//   SPIN     1: test_wait spinning instead of try_wait
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/pipe_rate tools/pipe_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fff);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d), "l"(da),
                 "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
template <int SPIN>
__device__ __forceinline__ void wait(uint32_t bar, uint32_t parity) {
    if (SPIN) {
        uint32_t done;
        do {
            asm volatile("{\n.reg .pred p;\nmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        } while (!done);
    } else {
        asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(bar), "r"(parity), "r"(20000u) : "memory");
    }
}
__device__ __forceinline__ void arrive(uint32_t bar) { asm volatile("{ .reg .b64 st; mbarrier.arrive.shared::cta.b64 st, [%0]; }" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void arrive_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("{ .reg .b64 st; mbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1; }" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ bool elect() {
    uint32_t pred = 0;
    asm volatile("{\n.reg .b32 rx;\n.reg .pred px;\nelect.sync rx|px, 0xffffffff;\n@px mov.s32 %0, 1;\n}\n" : "+r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
          "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
          "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]),
          "=r"(v[31])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ float fmin3(float a, float b, float c) { float d; asm("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }

template <int TREE>
__device__ __forceinline__ float tree(const uint32_t (&v)[32]) {
    if (TREE == 0) return __uint_as_float(v[0] ^ v[16]);
    float f[32];
#pragma unroll
    for (int i = 0; i < 32; i++) f[i] = __uint_as_float(v[i]);
    if (TREE == 1) {
        float m[11];
#pragma unroll
        for (int i = 0; i < 10; i++) m[i] = fmin3(f[3 * i], f[3 * i + 1], f[3 * i + 2]);
        m[10] = fminf(f[30], f[31]);
        return fminf(fmin3(fmin3(m[0], m[1], m[2]), fmin3(m[3], m[4], m[5]), fmin3(m[6], m[7], m[8])), fminf(m[9], m[10]));
    } else {
        // 10 FMNMX3 (ALU pipe) + 11 FMNMX (FMA pipe): 20 + 11 values
        float a[7], b[6];
#pragma unroll
        for (int i = 0; i < 7; i++) a[i] = fmin3(f[3 * i], f[3 * i + 1], f[3 * i + 2]);      // 21 values
#pragma unroll
        for (int i = 0; i < 5; i++) b[i] = fminf(f[21 + 2 * i], f[22 + 2 * i]);             // 10 values
        b[5] = f[31];
        float c0 = fmin3(a[0], a[1], a[2]), c1 = fmin3(a[3], a[4], a[5]);                   // 2 more FMNMX3
        float d0 = fminf(b[0], b[1]), d1 = fminf(b[2], b[3]), d2 = fminf(b[4], b[5]);       // 3 FMNMX
        float e0 = fmin3(c0, c1, a[6]);                                                     // FMNMX3
        float e1 = fminf(d0, d1);
        return fminf(fminf(e0, e1), d2);
    }
}

constexpr int kStages = 6;

template <int NBUF, int GROUPS, int SPLIT, int ISS, int PROD, int TREE, int SPIN, int KP, int FEAT>
__global__ void __launch_bounds__((GROUPS * 4 + ISS + PROD + ((FEAT & 2) ? 4 : 0)) * 32, 1) k(const unsigned char* __restrict__ pages, int n_pages, int iters, float thr,
                                                                      float* out, unsigned long long* cycles) {
    constexpr int N = 512 / NBUF;          // columns per buffer fill (= MMA N)
    constexpr int UPP = 256 / N;           // buffer fills (units) per page
    constexpr int TEAMS = GROUPS / SPLIT;
    constexpr int COLS = N / SPLIT;        // columns per epilogue warp per unit
    static_assert(COLS % 64 == 0 || COLS == 32, "cols");
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char* stage_mem = smem;                               // [kStages][KP*512]
    __half* sA = reinterpret_cast<__half*>(smem + kStages * KP * 512);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sA + 128 * KP);   // full[6], empty[6], tfull[NBUF], tempty[NBUF]
    __shared__ uint32_t slot;
    __shared__ float tq_s[128];
    __shared__ unsigned int ring_resv[4], ring_head[4], done_flag[16];
    __shared__ unsigned long long ring[4][256];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid < 128) tq_s[tid] = thr;
    if (tid < 4) { ring_resv[tid] = 0; ring_head[tid] = 0; }
    if (tid < 16) done_flag[tid] = 0;
    const uint32_t full0 = smem_u32(bars), empty0 = full0 + 8 * kStages, tfull0 = empty0 + 8 * kStages, tempty0 = tfull0 + 8 * NBUF;
    for (int i = tid; i < 128 * KP; i += blockDim.x) sA[i] = __float2half(0.001f * (i % 97));
    if (tid == 0) {
        for (int s = 0; s < kStages; s++) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(full0 + 8 * s), "r"(1));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(empty0 + 8 * s), "r"(UPP));
        }
        for (int b = 0; b < NBUF; b++) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tfull0 + 8 * b), "r"(1));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tempty0 + 8 * b), "r"(4 * SPLIT));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = slot;
    const int page0 = (blockIdx.x * 97) % n_pages;
    long long t0 = clock64();
    float acc = 0.f;

    if (warp < GROUPS * 4) {
        // ---- epilogue ----
        const int quad = warp & 3, grp = warp >> 2, team = grp / SPLIT, mem = grp % SPLIT;
        const uint32_t tbase = tmem + ((uint32_t)(quad * 32) << 16) + mem * COLS;
        const int n_units = iters * UPP;
        // unit u -> buffer u % NBUF, team u % TEAMS; use count of a buffer: u / NBUF
        for (int u = team; u < n_units; u += TEAMS) {
            const uint32_t buf = u % NBUF, use = (u / NBUF) & 1;
            wait<SPIN>(tfull0 + 8 * buf, use);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t ta = tbase + buf * N;
            if (COLS == 32) {
                uint32_t va[32];
                ld32(ta, va);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                if (lane == 0) arrive(tempty0 + 8 * buf);
                float m = tree<TREE>(va);
                if (m <= thr) acc += m;
            } else {
#pragma unroll
                for (int c = 0; c < COLS; c += 64) {
                    uint32_t va[32], vb[32];
                    ld32(ta + c, va);
                    ld32(ta + c + 32, vb);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    if (c + 64 >= COLS) {
                        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                        if (lane == 0) arrive(tempty0 + 8 * buf);
                    }
                    if (FEAT & 1) {
                        // like the product kernel: per-query bound re-read from shared memory, one vote + branch per 32 columns,
                        // survivors reserved with a shared-memory atomic
                        const float tq = *reinterpret_cast<volatile float*>(&tq_s[quad * 32 + lane]);
                        const float ma = tree<TREE>(va);
                        bool hit = ma <= tq;
                        if (FEAT & 4) hit |= ((u * 2654435761u + warp * 40503u + lane * 97u + c) & 0x3fff) < 9;  // ~5 % of the warp calls
                        if (__any_sync(0xffffffffu, hit)) {
                            if (hit) {
                                unsigned int sl = atomicAdd(&ring_resv[quad], 1u);
                                reinterpret_cast<volatile unsigned long long*>(ring[quad])[sl & 255] = ((unsigned long long)lane << 32) | u;
                            }
                            __syncwarp();
                        }
                        const float mb = tree<TREE>(vb);
                        hit = mb <= tq;
                        if (FEAT & 4) hit |= ((u * 2246822519u + warp * 40503u + lane * 97u + c) & 0x3fff) < 9;
                        if (__any_sync(0xffffffffu, hit)) {
                            if (hit) {
                                unsigned int sl = atomicAdd(&ring_resv[quad], 1u);
                                reinterpret_cast<volatile unsigned long long*>(ring[quad])[sl & 255] = ((unsigned long long)lane << 32) | u;
                            }
                            __syncwarp();
                        }
                    } else {
                        float m = fminf(tree<TREE>(va), tree<TREE>(vb));
                        if (m <= thr) acc += m;
                    }
                }
            }
        }
        if (lane == 0) *reinterpret_cast<volatile unsigned int*>(&done_flag[warp]) = 1;
    } else if (warp >= GROUPS * 4 + ISS + PROD) {
        // ---- rescorer stand-ins: poll the quadrant's ring counter like the product's rescorers ----
        const int quad = warp - (GROUPS * 4 + ISS + PROD);
        unsigned int head = 0;
        for (;;) {
            bool fin = true;
            for (int g = 0; g < GROUPS; g++) fin = fin && (*reinterpret_cast<volatile unsigned int*>(&done_flag[quad + 4 * g]) != 0);
            const unsigned int r = *reinterpret_cast<volatile unsigned int*>(&ring_resv[quad]);
            if (r != head) {
                unsigned long long e = reinterpret_cast<volatile unsigned long long*>(ring[quad])[(head + lane) & 255];
                acc += (float)(e & 0xff);
                head = r;
                *reinterpret_cast<volatile unsigned int*>(&ring_head[quad]) = head;
            } else {
                if (fin) break;
                __nanosleep(4000);
            }
        }
    } else if (warp < GROUPS * 4 + ISS) {
        // ---- MMA issuers: unit u handled by issuer u % ISS ----
        const int me = warp - GROUPS * 4;
        const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint64_t da0 = make_desc(smem_u32(sA), 16 * 128, 128);
        const uint64_t db0 = make_desc(smem_u32(stage_mem), 32 * 128, 128);
        const int n_units = iters * UPP;
        for (int u = me; u < n_units; u += ISS) {
            const int pg = u / UPP, h = u % UPP;
            const uint32_t st = pg % kStages, ph = (pg / kStages) & 1;
            const uint32_t buf = u % NBUF, use = (u / NBUF) & 1;
            wait<SPIN>(full0 + 8 * st, ph);
            wait<SPIN>(tempty0 + 8 * buf, use ^ 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (elect()) {
                const uint64_t db = db0 + (uint64_t)(st * (KP * 512 / 16) + h * (N / 8) * 128 / 16);
#pragma unroll
                for (int k16 = 0; k16 < KP / 16; k16++) umma(tmem + buf * N, da0 + (uint64_t)(k16 * 2 * 16 * 128 / 16), db + (uint64_t)(k16 * 2 * 32 * 128 / 16), idesc, k16 > 0);
                commit(tfull0 + 8 * buf);
                commit(empty0 + 8 * st);
            }
            __syncwarp();
        }
    } else if (warp < GROUPS * 4 + ISS + PROD) {
        // ---- producers: page p handled by producer p % PROD ----
        const int me = warp - GROUPS * 4 - ISS;
        for (int pg = me; pg < iters; pg += PROD) {
            const uint32_t st = pg % kStages, ph = (pg / kStages) & 1;
            wait<SPIN>(empty0 + 8 * st, ph ^ 1);
            if (elect()) {
                arrive_tx(full0 + 8 * st, KP * 512);
                bulk(smem_u32(stage_mem) + st * KP * 512, pages + (size_t)((page0 + pg) % n_pages) * KP * 512, KP * 512, full0 + 8 * st);
            }
            __syncwarp();
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + tid] = acc;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    long long t2 = clock64();
    if (tid == 0) atomicAdd(cycles, (unsigned long long)(t2 - t0));
    (void)t1;
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

template <int NBUF, int GROUPS, int SPLIT, int ISS, int PROD, int TREE, int SPIN, int KP, int FEAT = 0>
void run(int sms, const unsigned char* pages, int n_pages, float* out, unsigned long long* cyc) {
    const int iters = 6000;
    auto kern = k<NBUF, GROUPS, SPLIT, ISS, PROD, TREE, SPIN, KP, FEAT>;
    size_t smem = (size_t)kStages * KP * 512 + 128 * KP * 2 + 512;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int threads = (GROUPS * 4 + ISS + PROD + ((FEAT & 2) ? 4 : 0)) * 32;
    kern<<<sms, threads, smem>>>(pages, n_pages, 50, -1.f, out, cyc);
    cudaMemset(cyc, 0, 8);
    kern<<<sms, threads, smem>>>(pages, n_pages, iters, -1.f, out, cyc);
    cudaError_t e = cudaDeviceSynchronize();
    unsigned long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("NBUF=%d N=%3d groups=%d split=%d teams=%d iss=%d prod=%d tree=%d spin=%d K'=%d feat=%d : %7.1f cycles/page  %s\n", NBUF, 512 / NBUF, GROUPS, SPLIT,
           GROUPS / SPLIT, ISS, PROD, TREE, SPIN, KP, FEAT, (double)c / sms / iters, cudaGetErrorString(e));
    fflush(stdout);
}

int main(int argc, char** argv) {
    cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
    int sms = pr.multiProcessorCount;
    printf("%s, %d SMs\n", pr.name, sms);
    const int n_pages = 3907;
    unsigned char* pages; cudaMalloc(&pages, (size_t)n_pages * 32 * 512); cudaMemset(pages, 0x11, (size_t)n_pages * 32 * 512);
    float* out; unsigned long long* cyc;
    cudaMalloc(&out, 4 * 1024 * 1024 * 4); cudaMalloc(&cyc, 8);
    const int set = argc > 1 ? atoi(argv[1]) : 0;
    if (set == 0) {
        //   NBUF GROUPS SPLIT ISS PROD TREE SPIN KP
        run<2, 4, 4, 2, 2, 0, 0, 32>(sms, pages, n_pages, out, cyc);   // v2 structure, no tree
        run<2, 4, 4, 2, 2, 1, 0, 32>(sms, pages, n_pages, out, cyc);   // v2 + tree
        run<2, 4, 4, 2, 2, 2, 0, 32>(sms, pages, n_pages, out, cyc);   // v2 + mixed tree
        run<2, 4, 4, 2, 2, 0, 1, 32>(sms, pages, n_pages, out, cyc);   // v2 spin
        run<2, 4, 4, 1, 1, 0, 0, 32>(sms, pages, n_pages, out, cyc);   // single issuer/producer
        run<4, 4, 2, 2, 2, 0, 0, 32>(sms, pages, n_pages, out, cyc);   // old structure (half pages), 2 issuers
        run<4, 4, 2, 4, 2, 0, 0, 32>(sms, pages, n_pages, out, cyc);   // old structure, 4 issuers
        run<4, 4, 2, 4, 2, 1, 0, 32>(sms, pages, n_pages, out, cyc);
        run<4, 4, 2, 4, 2, 2, 0, 32>(sms, pages, n_pages, out, cyc);
        run<4, 4, 1, 4, 2, 0, 0, 32>(sms, pages, n_pages, out, cyc);   // one team per buffer, 128 cols per warp
        run<4, 4, 1, 4, 2, 1, 0, 32>(sms, pages, n_pages, out, cyc);
        run<4, 4, 1, 4, 2, 2, 0, 32>(sms, pages, n_pages, out, cyc);
        run<4, 4, 4, 4, 2, 0, 0, 32>(sms, pages, n_pages, out, cyc);   // 16 warps on every half page, 32 cols each
        run<4, 2, 2, 4, 2, 0, 0, 32>(sms, pages, n_pages, out, cyc);   // 8 epilogue warps
        run<4, 2, 1, 4, 2, 2, 0, 32>(sms, pages, n_pages, out, cyc);
        run<8, 4, 1, 4, 2, 0, 0, 32>(sms, pages, n_pages, out, cyc);   // 8 buffers of 64 columns
        run<8, 4, 1, 4, 2, 2, 0, 32>(sms, pages, n_pages, out, cyc);
    }
    if (set == 1) {
        // the product kernel's structure (4 buffers of 128 columns, 2 teams x 2 column halves, 4 issuers, 1 producer) with the
        // product's extra features switched on one by one: 1 = per-query bound from smem + vote + branch (+ atomic push code),
        // 2 = four polling rescorer warps, 4 = ~5 % of the warp calls really push
        run<4, 4, 2, 4, 1, 2, 0, 32, 0>(sms, pages, n_pages, out, cyc);
        run<4, 4, 2, 4, 1, 2, 0, 32, 1>(sms, pages, n_pages, out, cyc);
        run<4, 4, 2, 4, 1, 2, 0, 32, 2>(sms, pages, n_pages, out, cyc);
        run<4, 4, 2, 4, 1, 2, 0, 32, 3>(sms, pages, n_pages, out, cyc);
        run<4, 4, 2, 4, 1, 2, 0, 32, 7>(sms, pages, n_pages, out, cyc);
        run<4, 4, 2, 4, 1, 1, 0, 32, 7>(sms, pages, n_pages, out, cyc);
        run<4, 4, 2, 4, 1, 0, 0, 32, 7>(sms, pages, n_pages, out, cyc);
    }
    return 0;
}
