// Micro-benchmark: issue rate of the FP32 instructions the leaf-scan inner loop is made of, on one B200.
// Prints lane-operations per clock per SM (clock64 deltas inside the kernel) for:
//   ffma      scalar FFMA  acc = fma(d, d, acc)                 (2 distinct source registers)
//   ffma3     scalar FFMA  acc = fma(a, b, acc)                 (3 distinct source registers)
//   pair      scalar       d = p - q; acc = fma(d, d, acc)      (the distance step, scalar form)
//   ffma2     packed FFMA2 acc2 = fma2(d2, d2, acc2)
//   pair2     packed       d2 = p2 - {q,q}; acc2 = fma2(d2,d2,acc2)   (the distance step as the kernel issues it)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/fp32_rate tools/fp32_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t pack2(float lo, float hi) { uint64_t d; asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi)); return d; }
__device__ __forceinline__ uint64_t sub2(uint64_t a, uint64_t b) { uint64_t d; asm volatile("sub.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ float ffma(float a, float b, float c) { float d; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
__device__ __forceinline__ float fsub(float a, float b) { float d; asm volatile("sub.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b)); return d; }

constexpr int CH = 16;  // independent chains per thread

template <int MODE>
__global__ void k(int iters, float seed, float* out, unsigned long long* cycles) {
    float a[CH], p[CH];
    uint64_t a2[CH], p2[CH];
    float q = seed * 0.5f, b = seed + 1.f;
#pragma unroll
    for (int i = 0; i < CH; i++) { a[i] = seed * i; p[i] = seed + i; a2[i] = pack2(seed * i, seed); p2[i] = pack2(seed + i, seed - i); }
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < CH; i++) {
            if (MODE == 0) a[i] = ffma(p[i], p[i], a[i]);
            if (MODE == 1) a[i] = ffma(p[i], b, a[i]);
            if (MODE == 2) { float d = fsub(p[i], q); a[i] = ffma(d, d, a[i]); }
            if (MODE == 3) a2[i] = fma2(p2[i], p2[i], a2[i]);
            if (MODE == 4) { uint64_t d = sub2(p2[i], pack2(q, q)); a2[i] = fma2(d, d, a2[i]); }
        }
    }
    long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < CH; i++) s += a[i] + (float)(a2[i] >> 40);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) atomicAdd(cycles, (unsigned long long)(t1 - t0));
}

template <int MODE>
void run(const char* name, int lanes_per_instr_pair, int ctas_per_sm, int threads, int sms, float* out, unsigned long long* cyc) {
    const int iters = 20000;
    int grid = sms * ctas_per_sm;
    cudaMemset(cyc, 0, 8);
    k<MODE><<<grid, threads>>>(10, 1.0f, out, cyc);
    cudaMemset(cyc, 0, 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<grid, threads>>>(iters, 1.0f, out, cyc);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    unsigned long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    double avg_cycles = (double)c / grid;
    // "lane ops" = scalar fp32 operations (an FFMA2 counts 2 per lane)
    double ops = (double)grid * threads * iters * CH * lanes_per_instr_pair;
    printf("%-6s ctas/sm=%d thr=%d : %.1f fp32 lane-ops/clk/SM, %.2f T lane-ops/s, %.3f ms, clk %.0f MHz\n", name, ctas_per_sm, threads,
           ops / sms / avg_cycles, ops / ms * 1e-9, ms, avg_cycles / ms * 1e-3);
}

int main() {
    cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
    int sms = pr.multiProcessorCount;
    printf("%s, %d SMs\n", pr.name, sms);
    float* out; unsigned long long* cyc;
    cudaMalloc(&out, 4 * 1024 * 1024 * 4); cudaMalloc(&cyc, 8);
    for (int cps : {1, 2, 4}) {
        run<0>("ffma", 1, cps, 256, sms, out, cyc);
        run<1>("ffma3", 1, cps, 256, sms, out, cyc);
        run<2>("pair", 2, cps, 256, sms, out, cyc);
        run<3>("ffma2", 2, cps, 256, sms, out, cyc);
        run<4>("pair2", 4, cps, 256, sms, out, cyc);
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
