// Micro-benchmark: TMEM read throughput of tcgen05.ld.32x32b (the accumulator read of the prefilter epilogue) on one B200.
// One CTA per SM allocates all 512 TMEM columns; W warps each read their lane quadrant (warp%4) repeatedly.
// Prints bytes per clock per SM.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/tmem_rate tools/tmem_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int X>
__device__ __forceinline__ uint32_t ld(uint32_t taddr);
template <>
__device__ __forceinline__ uint32_t ld<32>(uint32_t taddr) {
    uint32_t v[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
          "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
          "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]),
          "=r"(v[31])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 32; i += 8) s ^= v[i];
    return s;
}
template <>
__device__ __forceinline__ uint32_t ld<16>(uint32_t taddr) {
    uint32_t v[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
          "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    return v[0] ^ v[8];
}

template <int X, bool TWO_IN_FLIGHT>
__global__ void k(int iters, uint32_t* out, unsigned long long* cycles) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = slot;
    const uint32_t base = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t acc = 0;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int c = 0; c < 512; c += X) acc ^= ld<X>(base + ((c + (warp >> 2) * X) & 511));
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) atomicAdd(cycles, (unsigned long long)(t1 - t0));
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

template <int X>
void run(int warps, int sms, uint32_t* out, unsigned long long* cyc) {
    const int iters = 2000;
    k<X, false><<<sms, warps * 32>>>(10, out, cyc);
    cudaMemset(cyc, 0, 8);
    k<X, false><<<sms, warps * 32>>>(iters, out, cyc);
    cudaError_t e = cudaDeviceSynchronize();
    unsigned long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    double avg = (double)c / sms;
    double bytes = (double)warps * iters * 512 * 32 * 4;  // every warp reads 512 columns x 32 lanes x 4 B per iteration
    printf("x%-3d warps=%2d : %.1f B/clk/SM  (%.1f cycles per 32x%d load per warp)  %s\n", X, warps, bytes / avg, avg / (iters * (512 / X)), X, cudaGetErrorString(e));
}

int main() {
    cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
    int sms = pr.multiProcessorCount;
    printf("%s, %d SMs\n", pr.name, sms);
    uint32_t* out; unsigned long long* cyc;
    cudaMalloc(&out, 4 * 1024 * 1024 * 4); cudaMalloc(&cyc, 8);
    for (int w : {4, 8, 16, 32}) { run<32>(w, sms, out, cyc); run<16>(w, sms, out, cyc); }
    return 0;
}
