// TEST INFRASTRUCTURE — NOT PRODUCT CODE.
// Compiles the drop-in subclass integration/B200_Convex_Tree.h against the UNMODIFIED reference headers
// (from /root/reference/Code, never copied) and links libscht_b200.so: the reference's own constructor builds the
// tree on the host (srand(seed) first), B200_Convex_Tree flattens it and the CUDA library answers kNN().
// Output file format == ref_knn's (oracle/io_formats.py), so the GPU tests compare it with the CPU_Convex_Tree
// fixtures bit for bit.  Built into oracle/_ref/ref_b200_knn by oracle/Makefile (only where the reference exists);
// the binary travels to the GPU box with the snapshot.
#include <CL/cl.h>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>
#include <windows.h>
#include "basic_functions.h"
#include "ConvexNode.h"
#include "Array.hh"
#include "Convex_Tree.h"
#include "openCL_tool.h"
#include "cyw_types.h"
#include "B200_Convex_Tree.h"

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>

extern "C" {
cl_int clGetPlatformIDs(cl_uint, cl_platform_id*, cl_uint*) { return -1; }
cl_int clGetPlatformInfo(cl_platform_id, cl_platform_info, size_t, void*, size_t*) { return -1; }
cl_int clGetDeviceIDs(cl_platform_id, cl_device_type, cl_uint, cl_device_id*, cl_uint*) { return -1; }
cl_int clGetDeviceInfo(cl_device_id, cl_device_info, size_t, void*, size_t*) { return -1; }
cl_context clCreateContext(const cl_context_properties*, cl_uint, const cl_device_id*, void (*)(const char*, const void*, size_t, void*),
                           void*, cl_int*) { return 0; }
cl_command_queue clCreateCommandQueue(cl_context, cl_device_id, cl_command_queue_properties, cl_int*) { return 0; }
cl_int clReleaseMemObject(cl_mem) { return -1; }
cl_int clReleaseProgram(cl_program) { return -1; }
cl_program clCreateProgramWithSource(cl_context, cl_uint, const char**, const size_t*, cl_int*) { return 0; }
cl_int clBuildProgram(cl_program, cl_uint, const cl_device_id*, const char*, void (*)(cl_program, void*), void*) { return -1; }
cl_int clGetProgramBuildInfo(cl_program, cl_device_id, cl_program_build_info, size_t, void*, size_t*) { return -1; }
cl_kernel clCreateKernel(cl_program, const char*, cl_int*) { return 0; }
}

static Matrix<float> read_fmat(const char* path) {
    FILE* f = fopen(path, "rb");
    if (!f) { fprintf(stderr, "cannot open %s\n", path); _Exit(2); }
    int32_t hdr[2];
    if (fread(hdr, 4, 2, f) != 2) _Exit(2);
    Matrix<float> m(hdr[0], hdr[1]);
    if (fread(m.get_matrix_raw_data(), 4, (size_t)hdr[0] * hdr[1], f) != (size_t)hdr[0] * hdr[1]) _Exit(2);
    fclose(f);
    return m;
}
static void put_rec(FILE* f, const char* name, int dtype, int64_t count, const void* data) {
    int32_t nl = (int32_t)strlen(name);
    static const int esz[4] = {4, 4, 8, 8};
    fwrite(&nl, 4, 1, f); fwrite(name, 1, nl, f);
    int32_t dt = dtype; fwrite(&dt, 4, 1, f); fwrite(&count, 8, 1, f);
    fwrite(data, esz[dtype], count, f);
}

int main(int argc, char** argv) {
    if (argc < 8 || argc > 11) {
        fprintf(stderr, "ref_b200_knn <data.fmat> <pct> <seed> <queries.fmat> <K> <brute 0|1> <out.knn> [n_gpus (0 = device 0 only) [mode 0|1 [reps]]]\n");
        return 1;
    }
    const int n_gpus = argc > 8 ? atoi(argv[8]) : 0, mode = argc > 9 ? atoi(argv[9]) : 0, reps = argc > 10 ? atoi(argv[10]) : 1;
    std::cout.setstate(std::ios_base::failbit);
    Matrix<float> data = read_fmat(argv[1]);
    float pct = (float)atof(argv[2]);
    Matrix<float> q = read_fmat(argv[4]);
    int K = atoi(argv[5]), brute = atoi(argv[6]);
    srand((unsigned)atoi(argv[3]));
    // used through the reference's base-class interface; n_gpus > 0: that many GPUs behind the same class
    std::vector<int> devs;
    for (int i = 0; i < n_gpus; i++) devs.push_back(i);
    Convex_Tree<float>* t = n_gpus > 0 ? new B200_Convex_Tree<float>(data, pct, devs, mode) : new B200_Convex_Tree<float>(data, pct);
    t->set_batch_size(1000000);
    // reps > 1: the same call repeated (the results of a call are owned by the tree and replaced by the next one); the time of
    // the LAST call goes to stderr -- end to end through the drop-in: pageable Matrix<T> in, std::vector results out
    for (int r = 0; r < reps; r++) {
        struct timespec a, b;
        clock_gettime(CLOCK_MONOTONIC, &a);
        if (brute) t->brute_force_kNN(q, K); else t->kNN(q, K);
        clock_gettime(CLOCK_MONOTONIC, &b);
        if (r == reps - 1) fprintf(stderr, "ref_b200_knn: call_ms=%.3f queries=%u K=%d n_gpus=%d mode=%d\n", (b.tv_sec - a.tv_sec) * 1e3 + (b.tv_nsec - a.tv_nsec) * 1e-6, q.nrows(), K, n_gpus, mode);
    }
    int nq = (int)q.nrows();
    std::vector<int32_t> idx((size_t)nq * K);
    std::vector<float> d2((size_t)nq * K), d((size_t)nq * K);
    for (int i = 0; i < nq; i++) {
        memcpy(&idx[(size_t)i * K], t->get_kNN_indexes(i), K * 4);
        memcpy(&d2[(size_t)i * K], t->get_kNN_dists_squre(i), K * 4);
        memcpy(&d[(size_t)i * K], t->get_kNN_dists(i), K * 4);
    }
    FILE* f = fopen(argv[7], "wb");
    put_rec(f, "idx", 0, (int64_t)nq * K, idx.data());
    put_rec(f, "dist2", 1, (int64_t)nq * K, d2.data());
    put_rec(f, "dist", 1, (int64_t)nq * K, d.data());
    fclose(f);
    fflush(NULL);
    _Exit(0);
}
