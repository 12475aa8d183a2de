// Micro-benchmark: issue-to-completion throughput of tcgen05.mma kind::f16 (FP16 operands, FP32 accumulate) in the exact
// configuration of the prefilter scan: cta_group::1, M=128, A and B in shared memory in the K-major NO-swizzle layout,
// K' = 32 per accumulator (two K16 instructions), accumulators rotating through TMEM.  One CTA per SM, one issuing thread.
// Prints cycles per 128x256 "page" for N=128 (two MMAs chains per page, as k_scan_tc issues them) and N=256.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/mma_rate tools/mma_rate.cu
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

// KP: K per accumulator chain; NW: N of one MMA (128 or 256)
template <int NW>
__global__ void k(int pages, int KP, unsigned long long* cycles) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __half* sA = reinterpret_cast<__half*>(smem);            // [KP/8][16][8][8]
    __half* sB = sA + 128 * KP;                              // [KP/8][32][8][8]
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < (128 + 256) * KP; i += blockDim.x) sA[i] = __float2half(0.001f * (i % 97));
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
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
    if (tid == 0) {
        const uint32_t idesc = (1u << 4) | ((uint32_t)(NW >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint64_t da0 = make_desc(smem_u32(sA), 16 * 128, 128);
        const uint64_t db0 = make_desc(smem_u32(sB), 32 * 128, 128);
        long long t0 = clock64();
        for (int p = 0; p < pages; p++) {
            for (int h = 0; h < 256 / NW; h++) {
                const uint32_t d = tmem + (uint32_t)(((p & 1) * 256 + h * NW) & 511);
                for (int k16 = 0; k16 < KP / 16; k16++)
                    umma(d, da0 + (uint64_t)(k16 * 2 * 16 * 128 / 16), db0 + (uint64_t)(k16 * 2 * 32 * 128 / 16 + h * (NW / 8) * 128 / 16), idesc, k16 > 0);
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        uint32_t done = 0;
        while (!done)
            asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(smem_u32(&bar)), "r"(0) : "memory");
        long long t1 = clock64();
        atomicAdd(cycles, (unsigned long long)(t1 - t0));
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

template <int NW>
void run(int KP, int sms, unsigned long long* cyc) {
    const int pages = 4000;
    size_t smem = (size_t)(128 + 256) * KP * 2 + 1024;
    cudaFuncSetAttribute(k<NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k<NW><<<sms, 128, smem>>>(10, KP, cyc);
    cudaMemset(cyc, 0, 8);
    k<NW><<<sms, 128, smem>>>(pages, KP, cyc);
    cudaError_t e = cudaDeviceSynchronize();
    unsigned long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    double per_page = (double)c / sms / pages;
    printf("N=%3d K'=%3d : %.1f cycles per 128x256 page  (%.0f FP16 FMA/clk/SM)  %s\n", NW, KP, per_page, 128.0 * 256 * KP / per_page, cudaGetErrorString(e));
}

int main() {
    cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
    int sms = pr.multiProcessorCount;
    printf("%s, %d SMs\n", pr.name, sms);
    unsigned long long* cyc; cudaMalloc(&cyc, 8);
    for (int KP : {16, 32, 48, 144}) { run<128>(KP, sms, cyc); run<256>(KP, sms, cyc); }
    return 0;
}
