// Micro-benchmark: issue rate of the instructions an accumulator epilogue can be built from (min trees, compares,
// logic, packed half min), on one B200.  Prints warp-instructions per clock per SM sub-partition x 32 = lanes/clk/SM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/alu_rate tools/alu_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

constexpr int CH = 12;  // independent chains per thread

__device__ __forceinline__ float fmin2(float a, float b) { float d; asm volatile("min.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b)); return d; }
__device__ __forceinline__ float fmin3(float a, float b, float c) { float d; asm volatile("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
__device__ __forceinline__ int imin3(int a, int b, int c) { return min(min(a, b), c); }
__device__ __forceinline__ uint32_t lop3_or(uint32_t a, uint32_t b, uint32_t c) { uint32_t d; asm volatile("lop3.b32 %0, %1, %2, %3, 0xFE;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }
__device__ __forceinline__ uint32_t hmin2(uint32_t a, uint32_t b) { uint32_t d; asm volatile("min.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }
__device__ __forceinline__ uint32_t hmin2_3(uint32_t a, uint32_t b, uint32_t c) { return hmin2(hmin2(a, b), c); }
__device__ __forceinline__ uint32_t smin16x2_3(uint32_t a, uint32_t b, uint32_t c) { return __vmins2(__vmins2(a, b), c); }
__device__ __forceinline__ float ffma(float a, float b, float c) { float d; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
__device__ __forceinline__ float fadd(float a, float b) { float d; asm volatile("add.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b)); return d; }
__device__ __forceinline__ uint32_t iadd3(uint32_t a, uint32_t b, uint32_t c) { uint32_t d; asm volatile("{.reg .u32 t; add.u32 t, %1, %2; add.u32 %0, t, %3;}" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }

template <int MODE>
__global__ void k(int iters, float seed, float* out, unsigned long long* cycles) {
    float a[CH], b[CH];
    uint32_t u[CH], w[CH];
    float c = seed * 3.f;
#pragma unroll
    for (int i = 0; i < CH; i++) { a[i] = seed * (i + 1); b[i] = seed + i + threadIdx.x; u[i] = __float_as_uint(a[i]) + i; w[i] = __float_as_uint(b[i]) * 7 + i; }
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < CH; i++) {
            if (MODE == 0) a[i] = fmin2(a[i], b[i]);
            if (MODE == 1) a[i] = fmin3(a[i], b[i], c);
            if (MODE == 2) u[i] = (uint32_t)imin3((int)u[i], (int)w[i], it);
            if (MODE == 3) u[i] = lop3_or(u[i], w[i], (uint32_t)it);
            if (MODE == 4) u[i] = hmin2(u[i], w[i]);
            if (MODE == 5) u[i] = hmin2_3(u[i], w[i], (uint32_t)it);
            if (MODE == 6) u[i] = smin16x2_3(u[i], w[i], (uint32_t)it);
            if (MODE == 7) a[i] = ffma(a[i], b[i], c);
            if (MODE == 8) { if (i & 1) a[i] = ffma(a[i], b[i], c); else a[i] = fmin3(a[i], b[i], c); }  // 1:1 mix FMA pipe / ALU pipe
            if (MODE == 9) a[i] = fadd(a[i], b[i]);
            if (MODE == 10) { if (i % 3 == 2) a[i] = fmin3(a[i], b[i], c); else a[i] = ffma(a[i], b[i], c); }  // 2 FFMA : 1 FMNMX3
            if (MODE == 11) u[i] = iadd3(u[i], w[i], (uint32_t)it);
        }
    }
    long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < CH; i++) s += a[i] + (float)u[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) atomicAdd(cycles, (unsigned long long)(t1 - t0));
}

template <int MODE>
void run(const char* name, int threads, int sms, float* out, unsigned long long* cyc) {
    const int iters = 20000;
    int grid = sms;
    k<MODE><<<grid, threads>>>(10, 1.0f, out, cyc);
    cudaMemset(cyc, 0, 8);
    k<MODE><<<grid, threads>>>(iters, 1.0f, out, cyc);
    cudaDeviceSynchronize();
    unsigned long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    double avg_cycles = (double)c / grid;
    double instr = (double)threads * iters * CH;  // lane-instructions per SM
    printf("%-14s thr=%4d : %.1f lane-instr/clk/SM\n", name, threads, instr / avg_cycles);
}

int main() {
    cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
    int sms = pr.multiProcessorCount;
    printf("%s, %d SMs\n", pr.name, sms);
    float* out; unsigned long long* cyc;
    cudaMalloc(&out, 4 * 1024 * 1024 * 4); cudaMalloc(&cyc, 8);
    for (int thr : {512, 1024}) {
        run<0>("fmnmx", thr, sms, out, cyc);
        run<1>("fmnmx3", thr, sms, out, cyc);
        run<2>("vimnmx3.s32", thr, sms, out, cyc);
        run<3>("lop3", thr, sms, out, cyc);
        run<4>("hmnmx2", thr, sms, out, cyc);
        run<5>("hmnmx2_3in", thr, sms, out, cyc);
        run<6>("vimnmx3.s16x2", thr, sms, out, cyc);
        run<7>("ffma", thr, sms, out, cyc);
        run<8>("ffma+fmnmx3", thr, sms, out, cyc);
        run<9>("fadd", thr, sms, out, cyc);
        run<10>("2ffma+fmnmx3", thr, sms, out, cyc);
        run<11>("iadd x2", thr, sms, out, cyc);
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
