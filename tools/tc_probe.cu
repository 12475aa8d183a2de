// Probe for the tensor-core prefilter: one CTA, D[128x256] = A[128xK] * B[256xK]^T with tcgen05.mma kind::tf32,
// operands in the K-major no-swizzle ("interleave") shared-memory layout, accumulator in TMEM, read back with
// tcgen05.ld 32x32b.x32.  Checks (a) descriptor/layout correctness against a CPU product and (b) how the hardware
// treats the 13 low mantissa bits of fp32 operands (truncate vs round) and how exact the fp32 accumulation is.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/tc_probe tools/tc_probe.cu
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>

constexpr int M = 128, N = 256;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fff);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
    return d;                // base_offset 0, lbo_mode 0, layout_type 0 (no swizzle)
}

// smem image: [kc][g][8 rows][4 floats]: element (row, k) at ((k/4)*G + row/8)*32 + (row%8)*4 + k%4 floats
__global__ void probe(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ C, int K) {
    extern __shared__ __align__(1024) unsigned char smem[];
    float* sA = reinterpret_cast<float*>(smem);
    float* sB = sA + M * K;
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < M * K; i += blockDim.x) {
        int row = i / K, k = i % K;
        sA[((k / 4) * (M / 8) + row / 8) * 32 + (row % 8) * 4 + k % 4] = A[i];
    }
    for (int i = tid; i < N * K; i += blockDim.x) {
        int row = i / K, k = i % K;
        sB[((k / 4) * (N / 8) + row / 8) * 32 + (row % 8) * 4 + k % 4] = B[i];
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(256));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    // make the generic-proxy smem writes visible to the async proxy (tensor core reads smem through it)
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;

    if (tid == 0) {
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
        for (int k8 = 0; k8 < K / 8; k8++) {
            uint64_t da = make_desc(smem_u32(sA) + (uint32_t)k8 * 2 * (M / 8) * 128, (M / 8) * 128, 128);
            uint64_t db = make_desc(smem_u32(sB) + (uint32_t)k8 * 2 * (N / 8) * 128, (N / 8) * 128, 128);
            uint32_t acc = k8 > 0;
            asm volatile(
                "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem),
                "l"(da), "l"(db), "r"(idesc), "r"(acc)
                : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    // wait for the MMAs
    {
        uint32_t done = 0;
        while (!done) {
            asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                         : "=r"(done)
                         : "r"(smem_u32(&bar)), "r"(0)
                         : "memory");
        }
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // warp w reads lanes [32w, 32w+32): thread = row
    const int row = warp * 32 + lane;
    for (int c0 = 0; c0 < N; c0 += 32) {
        uint32_t v[32];
        uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
              "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
              "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
              "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int c = 0; c < 32; c++) C[(size_t)row * N + c0 + c] = __uint_as_float(v[c]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256));
}

static float tf32_trunc(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xffffe000u; memcpy(&x, &u, 4); return x; }
static float tf32_rn(float x) {
    uint32_t u; memcpy(&u, &x, 4);
    u += 0x00000fffu + ((u >> 13) & 1u);
    u &= 0xffffe000u;
    memcpy(&x, &u, 4);
    return x;
}

int run(int K, int mode) {
    std::vector<float> A((size_t)M * K), B((size_t)N * K), C((size_t)M * N);
    srand(7 + K + mode);
    for (auto& x : A) x = (float)rand() / RAND_MAX - 0.5f;
    for (auto& x : B) x = (float)rand() / RAND_MAX - 0.5f;
    if (mode == 1) {  // operands already on the tf32 grid: result should be (nearly) exact
        for (auto& x : A) x = tf32_rn(x);
        for (auto& x : B) x = tf32_rn(x);
    }
    if (mode == 2) {  // small integers: everything exact
        for (auto& x : A) x = (float)(rand() % 219);
        for (auto& x : B) x = (float)(rand() % 219);
    }
    float *dA, *dB, *dC;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dC, C.size() * 4);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
    cudaMemset(dC, 0xff, C.size() * 4);
    size_t smem = (size_t)(M + N) * K * 4;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    probe<<<1, 128, smem>>>(dA, dB, dC, K);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("K=%d mode=%d CUDA error: %s\n", K, mode, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(C.data(), dC, C.size() * 4, cudaMemcpyDeviceToHost);
    double err_exact = 0, err_trunc = 0, err_rn = 0, mag = 0;
    for (int i = 0; i < M; i++)
        for (int j = 0; j < N; j++) {
            double se = 0, st = 0, sr = 0, sa = 0;
            for (int k = 0; k < K; k++) {
                float a = A[(size_t)i * K + k], b = B[(size_t)j * K + k];
                se += (double)a * b;
                st += (double)tf32_trunc(a) * tf32_trunc(b);
                sr += (double)tf32_rn(a) * tf32_rn(b);
                sa += fabs((double)a * b);
            }
            double c = C[(size_t)i * N + j];
            err_exact = fmax(err_exact, fabs(c - se));
            err_trunc = fmax(err_trunc, fabs(c - st));
            err_rn = fmax(err_rn, fabs(c - sr));
            mag = fmax(mag, sa);
        }
    printf("K=%3d mode=%d: max|C-exact|=%.3e  max|C-trunc model|=%.3e  max|C-rn model|=%.3e  (max sum|a*b|=%.3e; 2^-10*mag=%.3e, 2^-23*mag=%.3e)\n",
           K, mode, err_exact, err_trunc, err_rn, mag, mag / 1024, mag / 8388608);
    cudaFree(dA); cudaFree(dB); cudaFree(dC);
    return 0;
}

int main() {
    int rc = 0;
    for (int K : {8, 16, 24, 32, 136})
        for (int mode : {0, 1, 2}) rc |= run(K, mode);
    printf("status %s\n", rc ? "FAILED" : "ok");
    return rc;
}
