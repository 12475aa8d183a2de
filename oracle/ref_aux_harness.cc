// TEST INFRASTRUCTURE — NOT PRODUCT CODE.
// Second driver over the UNMODIFIED reference sources (included from /root/reference/Code, never copied), for the two
// "next" rows of SURVEY §8 that are plain host code:
//   ref_aux readdata <file> <delims> <out.fmat>     read_data(filename, delims, &dim, &size)   (Code/data_processor.h:75-99)
//   ref_aux treeinfo <data.fmat> <pct> <seed> <out.txt>   Convex_Tree<T>::save_tree_info(FILE*)   (Code/Convex_Tree.h:964-991)
// tests/test_io_and_logs.py compares scht_read_text / scht_format_tree_info with these outputs byte for byte.
// Built into oracle/_ref/ref_aux by oracle/Makefile (only where the reference exists).
#include <CL/cl.h>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>
#include <windows.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
extern "C" char* gets(char*);  // removed from C11/C++14 headers; glibc still exports it (data_processor.h:36,52,80 call it)
#include "basic_functions.h"
#include "ConvexNode.h"
#include "Array.hh"
#include "Convex_Tree.h"
#include "openCL_tool.h"
#include "CPU_Convex_Tree.h"
#include "cyw_types.h"
#include "data_processor.h"

extern "C" {
cl_int clGetPlatformIDs(cl_uint, cl_platform_id*, cl_uint*) { return -1; }
cl_int clGetPlatformInfo(cl_platform_id, cl_platform_info, size_t, void*, size_t*) { return -1; }
cl_int clGetDeviceIDs(cl_platform_id, cl_device_type, cl_uint, cl_device_id*, cl_uint*) { return -1; }
cl_int clGetDeviceInfo(cl_device_id, cl_device_info, size_t, void*, size_t*) { return -1; }
cl_context clCreateContext(const cl_context_properties*, cl_uint, const cl_device_id*, void (*)(const char*, const void*, size_t, void*), void*,
                           cl_int*) { return 0; }
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

int main(int argc, char** argv) {
    if (argc == 5 && std::string(argv[1]) == "readdata") {
        int dim = 0, size = 0;
        float* data = read_data(argv[2], argv[3], &dim, &size);
        FILE* f = fopen(argv[4], "wb");
        int32_t hdr[2] = {size, dim};
        fwrite(hdr, 4, 2, f);
        fwrite(data, 4, (size_t)size * dim, f);
        fclose(f);
        fflush(NULL);
        _Exit(0);
    }
    if (argc == 6 && std::string(argv[1]) == "treeinfo") {
        std::cout.setstate(std::ios_base::failbit);
        Matrix<float> data = read_fmat(argv[2]);
        float pct = (float)atof(argv[3]);
        srand((unsigned)atoi(argv[4]));
        Convex_Tree<float>* t = new CPU_Convex_Tree<float>(data, pct, USE_CPU_LEAF_APPRXIMATE_QP);
        FILE* f = fopen(argv[5], "w");
        t->save_tree_info(f);
        fclose(f);
        fflush(NULL);
        _Exit(0);
    }
    fprintf(stderr, "ref_aux readdata <file> <delims> <out.fmat> | treeinfo <data.fmat> <pct> <seed> <out.txt>\n");
    return 1;
}
