// Micro-benchmark: the epilogue's 32-value min tree built from float minima (FMNMX3 + FMNMX, k_scan_tc's tree) against a
// signed-integer tree (VIMNMX3.S32 + VIMNMX) on the same bit patterns, and the raw rate of VIMNMX3 with three register
// operands.  Values come from shared memory (8 LDS.128 per 32 values, the stand-in for one tcgen05.ld.x32).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/min_tree_rate tools/min_tree_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ float fmin3f(float a, float b, float c) { float d; asm("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
__device__ __forceinline__ float fmin2f(float a, float b) { float d; asm("min.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b)); return d; }
__device__ __forceinline__ int imin3(int a, int b, int c) { return min(min(a, b), c); }

// MODE 0: float tree (2 FMNMX3 + 3 FMNMX per 8, then 3 FMNMX); 1: int tree (3 VIMNMX3 + 1 VIMNMX per 8, then VIMNMX3 + VIMNMX);
// 2: loads only (xor fold of 4 of the 32 values, so the loads stay)
template <int MODE>
__global__ void k_tree(int iters, int* out, unsigned long long* cycles) {
    __shared__ uint4 sm[8 * 64];
    for (int i = threadIdx.x; i < 8 * 64; i += blockDim.x) sm[i] = make_uint4(0x3f800000u + i * 977u, 0x3f900000u + i * 31u, 0x40000000u + i, 0x3fa00000u ^ (i * 7u));
    __syncthreads();
    const int tq_i = 0x3f800100;
    const float tq_f = __int_as_float(tq_i);
    int hits = 0;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
        uint32_t v[32];
        const uint4* src = sm + ((it + (threadIdx.x >> 5)) & 63) * 8;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            uint4 x = src[j];
            v[4 * j] = x.x; v[4 * j + 1] = x.y; v[4 * j + 2] = x.z; v[4 * j + 3] = x.w;
        }
        bool hit;
        if (MODE == 0) {
            float m[4];
#pragma unroll
            for (int g = 0; g < 4; g++) {
                const float x = fmin3f(__uint_as_float(v[8 * g]), __uint_as_float(v[8 * g + 1]), __uint_as_float(v[8 * g + 2]));
                const float y = fmin3f(__uint_as_float(v[8 * g + 3]), __uint_as_float(v[8 * g + 4]), __uint_as_float(v[8 * g + 5]));
                const float z = fmin2f(__uint_as_float(v[8 * g + 6]), __uint_as_float(v[8 * g + 7]));
                m[g] = fmin2f(fmin2f(x, y), z);
            }
            hit = fmin2f(fmin2f(m[0], m[1]), fmin2f(m[2], m[3])) <= tq_f;
        } else if (MODE == 1) {
            int m[4];
#pragma unroll
            for (int g = 0; g < 4; g++) {
                const int x = imin3((int)v[8 * g], (int)v[8 * g + 1], (int)v[8 * g + 2]);
                const int y = imin3((int)v[8 * g + 3], (int)v[8 * g + 4], (int)v[8 * g + 5]);
                const int z = imin3((int)v[8 * g + 6], (int)v[8 * g + 7], x);
                m[g] = min(y, z);
            }
            hit = min(imin3(m[0], m[1], m[2]), m[3]) <= tq_i;
        } else {
            hit = (int)(v[0] ^ v[9] ^ v[18] ^ v[27]) <= tq_i;
        }
        if (__any_sync(0xffffffffu, hit)) hits += hit;
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = hits;
    if (threadIdx.x == 0) atomicAdd(cycles, (unsigned long long)(t1 - t0));
}

// raw rate: 12 independent chains, every operand a register that changes
template <int MODE>
__global__ void k_chain(int iters, int seed, int* out, unsigned long long* cycles) {
    constexpr int CH = 12;
    int u[CH], w[CH];
#pragma unroll
    for (int i = 0; i < CH; i++) { u[i] = seed * (i + 3) + threadIdx.x; w[i] = seed * 77 + i * threadIdx.x; }
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < CH; i++) {
            if (MODE == 0) u[i] = imin3(u[i], u[(i + 5) % CH], w[i]);
            if (MODE == 1) u[i] = __float_as_int(fmin3f(__int_as_float(u[i]), __int_as_float(u[(i + 5) % CH]), __int_as_float(w[i])));
            if (MODE == 2) u[i] = min(u[i], u[(i + 5) % CH]);
            if (MODE == 3) u[i] = __float_as_int(fmin2f(__int_as_float(u[i]), __int_as_float(u[(i + 5) % CH])));
        }
#pragma unroll
        for (int i = 0; i < CH; i++) w[i] ^= u[(i + 1) % CH];  // keeps the values moving (1 LOP per min, same in every mode)
    }
    long long t1 = clock64();
    int s = 0;
#pragma unroll
    for (int i = 0; i < CH; i++) s += u[i] + w[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) atomicAdd(cycles, (unsigned long long)(t1 - t0));
}

template <typename F>
double timed(F launch, unsigned long long* cyc, int grid) {
    launch(10);
    cudaMemset(cyc, 0, 8);
    launch(20000);
    cudaDeviceSynchronize();
    unsigned long long c;
    cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    return (double)c / grid;
}

int main() {
    cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
    const int sms = pr.multiProcessorCount;
    printf("%s, %d SMs\n", pr.name, sms);
    int* out; unsigned long long* cyc;
    cudaMalloc(&out, 4 * 1024 * 1024); cudaMalloc(&cyc, 8);
    const int thr = 512;
    const char* tn[3] = {"float tree (FMNMX3+FMNMX)", "int tree (VIMNMX3+VIMNMX)", "loads only"};
    double c;
    c = timed([&](int it) { k_tree<0><<<sms, thr>>>(it, out, cyc); }, cyc, sms); printf("%-28s %.1f cycles per 32 values per warp-scheduler (4 warps)\n", tn[0], c / 20000 * 1.0);
    c = timed([&](int it) { k_tree<1><<<sms, thr>>>(it, out, cyc); }, cyc, sms); printf("%-28s %.1f\n", tn[1], c / 20000);
    c = timed([&](int it) { k_tree<2><<<sms, thr>>>(it, out, cyc); }, cyc, sms); printf("%-28s %.1f\n", tn[2], c / 20000);
    const char* cn[4] = {"vimnmx3 (3 regs) + lop", "fmnmx3 (3 regs) + lop", "vimnmx (2 regs) + lop", "fmnmx (2 regs) + lop"};
    for (int m = 0; m < 4; m++) {
        if (m == 0) c = timed([&](int it) { k_chain<0><<<sms, thr>>>(it, 3, out, cyc); }, cyc, sms);
        if (m == 1) c = timed([&](int it) { k_chain<1><<<sms, thr>>>(it, 3, out, cyc); }, cyc, sms);
        if (m == 2) c = timed([&](int it) { k_chain<2><<<sms, thr>>>(it, 3, out, cyc); }, cyc, sms);
        if (m == 3) c = timed([&](int it) { k_chain<3><<<sms, thr>>>(it, 3, out, cyc); }, cyc, sms);
        printf("%-28s %.1f lane-pairs(min+lop)/clk/SM\n", cn[m], (double)thr * 20000 * 12 / c);
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
