// TEST INFRASTRUCTURE — NOT PRODUCT CODE.
//
// Driver that compiles the UNMODIFIED reference (headers included from where they lie under
// /root/reference/Code, never copied) into oracle/_ref/ref_knn.  It is used
//   (1) to pin oracle/scht_oracle.c and the product's host tree builder against the reference itself,
//   (2) to generate tests/golden/*.npz (tests/golden/make_golden.py),
//   (3) as the "reference" CPU baseline of bench.py (kind = "reference").
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may run it.
//
// Include order mirrors Code/main.cc:10-24 minus GPU_Convex_Tree.h / data_processor.h (OpenCL, gets()).
//
// Reference quirks honoured here: kNN() may be called once per tree (a second call
// or the destructor free()s interior pointers, CPU_Convex_Tree.h:119-122) -> one kNN per process,
// leave with _Exit(); do_kNN allocates Q*nodes floats (CPU_Convex_Tree.h:136-143) -> callers chunk
// queries; rand() drives the tree (Convex_Tree.h:789) -> srand(seed) right before construction.
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
#include "CPU_Convex_Tree.h"
#include "cyw_types.h"

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>

// Dummy definitions of the OpenCL entry points openCL_tool.h's non-template functions reference.
// They are never called (no GPU_Convex_Tree is instantiated).
extern "C" {
cl_int clGetPlatformIDs(cl_uint, cl_platform_id*, cl_uint*) { return -1; }
cl_int clGetPlatformInfo(cl_platform_id, cl_platform_info, size_t, void*, size_t*) { return -1; }
cl_int clGetDeviceIDs(cl_platform_id, cl_device_type, cl_uint, cl_device_id*, cl_uint*) { return -1; }
cl_int clGetDeviceInfo(cl_device_id, cl_device_info, size_t, void*, size_t*) { return -1; }
cl_context clCreateContext(const cl_context_properties*, cl_uint, const cl_device_id*,
                           void (*)(const char*, const void*, size_t, void*), void*, cl_int*) { return 0; }
cl_command_queue clCreateCommandQueue(cl_context, cl_device_id, cl_command_queue_properties, cl_int*) { return 0; }
cl_int clReleaseMemObject(cl_mem) { return -1; }
cl_int clReleaseProgram(cl_program) { return -1; }
cl_program clCreateProgramWithSource(cl_context, cl_uint, const char**, const size_t*, cl_int*) { return 0; }
cl_int clBuildProgram(cl_program, cl_uint, const cl_device_id*, const char*, void (*)(cl_program, void*), void*) { return -1; }
cl_int clGetProgramBuildInfo(cl_program, cl_device_id, cl_program_build_info, size_t, void*, size_t*) { return -1; }
cl_kernel clCreateKernel(cl_program, const char*, cl_int*) { return 0; }
}

static double now_s() {
    struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

// ---- tiny binary formats shared with the python side (see oracle/io_formats.py) -----------------
// .fmat : int32 rows, int32 cols, float32[rows*cols]
// record file: repeated { int32 name_len, name, int32 dtype(0=i32,1=f32,2=i64,3=f64), int64 count, data }
static Matrix<float> read_fmat(const char* path, int row_begin = 0, int row_end = -1) {
    FILE* f = fopen(path, "rb");
    if (!f) { fprintf(stderr, "cannot open %s\n", path); _Exit(2); }
    int32_t hdr[2];
    if (fread(hdr, 4, 2, f) != 2) _Exit(2);
    int rows = hdr[0], cols = hdr[1];
    if (row_end < 0 || row_end > rows) row_end = rows;
    int n = row_end - row_begin;
    Matrix<float> m(n, cols);
    fseek(f, 8 + (long)row_begin * cols * 4, SEEK_SET);
    if (fread(m.get_matrix_raw_data(), 4, (size_t)n * cols, f) != (size_t)n * cols) _Exit(2);
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

// Probe subclass: only exposes protected state of the reference tree for dumping.
struct Probe : public CPU_Convex_Tree<float> {
    Probe(Matrix<float>& d, float pct, int alg) : CPU_Convex_Tree<float>(d, pct, alg) {}

    void dump_tree(const char* path) {
        FILE* f = fopen(path, "wb");
        int dim = this->m_data.ncols();
        int n = this->m_data.nrows();
        int nn = this->nodes_num, L = this->leaf_nodes_num;
        int32_t meta[6] = {dim, n, nn, L, this->depth, this->leaf_pts_num_threshold};
        put_rec(f, "meta", 0, 6, meta);
        put_rec(f, "sorted_points", 1, (int64_t)n * dim, this->m_sorted_data.get_matrix_raw_data());
        put_rec(f, "sorted_to_orig", 0, n, this->m_sorted_data_ori_indexes.get_matrix_raw_data());
        std::vector<int32_t> leaf_start(L), leaf_count(L), leaf_node(L);
        for (int i = 0; i < L; i++) {
            leaf_start[i] = this->leaf_nodes_start_pos_in_data_set[i];
            leaf_node[i] = this->leaf_nodes_ori_indexes[i];
            leaf_count[i] = this->nodes[leaf_node[i]]->pts_number;
        }
        put_rec(f, "leaf_start", 0, L, leaf_start.data());
        put_rec(f, "leaf_count", 0, L, leaf_count.data());
        put_rec(f, "leaf_node_id", 0, L, leaf_node.data());
        std::vector<int32_t> left(nn), right(nn), parent(nn), leaf_index(nn), beta_off(nn + 1);
        std::vector<float> lc((size_t)nn * dim, 0.f), rc((size_t)nn * dim, 0.f), normal((size_t)nn * dim, 0.f), beta;
        for (int i = 0; i < nn; i++) {
            ConvexNode<float>* nd = this->nodes[i];
            left[i] = nd->left; right[i] = nd->right; parent[i] = nd->parent_index;
            leaf_index[i] = nd->isLeaf ? nd->leaf_index : -1;
            if (!nd->isLeaf) {
                for (int j = 0; j < dim; j++) { lc[(size_t)i * dim + j] = nd->left_center[j]; rc[(size_t)i * dim + j] = nd->right_center[j]; }
            }
            int c = (i == 0) ? 0 : (int)nd->ALPHA.nrows();
            beta_off[i] = (int32_t)beta.size();
            for (int k = 0; k < c; k++) beta.push_back(nd->BETA[k]);
            if (c > 0) {
                for (int j = 0; j < dim; j++) normal[(size_t)i * dim + j] = nd->ALPHA[c - 1][j];
                // sanity: inherited rows are bit-copies of the parent's rows (Convex_Tree.h:763)
                ConvexNode<float>* pn = this->nodes[nd->parent_index];
                int pc = (nd->parent_index == 0) ? 0 : (int)pn->ALPHA.nrows();
                if (pc != c - 1) { fprintf(stderr, "constraint count mismatch at node %d\n", i); _Exit(3); }
                for (int k = 0; k < pc; k++)
                    if (memcmp(pn->ALPHA[k], nd->ALPHA[k], dim * 4) != 0) { fprintf(stderr, "alpha mismatch at node %d\n", i); _Exit(3); }
            }
        }
        beta_off[nn] = (int32_t)beta.size();
        put_rec(f, "node_left", 0, nn, left.data());
        put_rec(f, "node_right", 0, nn, right.data());
        put_rec(f, "node_parent", 0, nn, parent.data());
        put_rec(f, "node_leaf_index", 0, nn, leaf_index.data());
        put_rec(f, "node_left_center", 1, (int64_t)nn * dim, lc.data());
        put_rec(f, "node_right_center", 1, (int64_t)nn * dim, rc.data());
        put_rec(f, "node_normal", 1, (int64_t)nn * dim, normal.data());
        put_rec(f, "node_beta_off", 0, nn + 1, beta_off.data());
        put_rec(f, "node_beta", 1, (int64_t)beta.size(), beta.data());
        fclose(f);
    }

    void dump_knn(const char* path, int nq, int K, double t_build, double t_knn) {
        FILE* f = fopen(path, "wb");
        std::vector<int32_t> idx((size_t)nq * K), appr(nq);
        std::vector<float> d2((size_t)nq * K), d((size_t)nq * K);
        for (int i = 0; i < nq; i++) {
            memcpy(&idx[(size_t)i * K], this->get_kNN_indexes(i), K * 4);
            memcpy(&d2[(size_t)i * K], this->get_kNN_dists_squre(i), K * 4);
            memcpy(&d[(size_t)i * K], this->get_kNN_dists(i), K * 4);
            appr[i] = this->appr_leaf_node_indexes[i];   // NODE index on the CPU path (CPU_Convex_Tree.h:215)
        }
        put_rec(f, "idx", 0, (int64_t)nq * K, idx.data());
        put_rec(f, "dist2", 1, (int64_t)nq * K, d2.data());
        put_rec(f, "dist", 1, (int64_t)nq * K, d.data());
        put_rec(f, "approx_leaf_node", 0, nq, appr.data());
        int64_t cnt[3] = {(int64_t)this->dist_computation_times_in_host, (int64_t)this->quadprog_cal_num,
                          (int64_t)this->scalar_product_num};
        put_rec(f, "counters", 2, 3, cnt);   // dist computations (incl. 2/level descent), hyperplane tests, dot products
        double tm[2] = {t_build, t_knn};
        put_rec(f, "seconds", 3, 2, tm);     // tree build, kNN() wall time
        fclose(f);
    }
};

static int usage() {
    fprintf(stderr,
            "ref_knn tree <data.fmat> <pct> <seed> <out.tree>\n"
            "ref_knn knn  <data.fmat> <pct> <seed> <queries.fmat> <K> <alg 0|2> <q_begin> <q_end> <out.knn> [tree_out]\n");
    return 1;
}

int main(int argc, char** argv) {
    if (argc < 2) return usage();
    std::string cmd = argv[1];
    std::cout.setstate(std::ios_base::failbit);   // silence the reference's progress prints
    if (cmd == "tree" && argc == 6) {
        Matrix<float> data = read_fmat(argv[2]);
        float pct = (float)atof(argv[3]);
        srand((unsigned)atoi(argv[4]));
        Probe* t = new Probe(data, pct, USE_CPU_LEAF_APPRXIMATE_QP);
        t->dump_tree(argv[5]);
        fflush(NULL); _Exit(0);
    }
    if (cmd == "knn" && (argc == 11 || argc == 12)) {
        Matrix<float> data = read_fmat(argv[2]);
        float pct = (float)atof(argv[3]);
        unsigned seed = (unsigned)atoi(argv[4]);
        int K = atoi(argv[6]), alg = atoi(argv[7]);
        int qb = atoi(argv[8]), qe = atoi(argv[9]);
        Matrix<float> q = read_fmat(argv[5], qb, qe);
        srand(seed);
        double t0 = now_s();
        Probe* t = new Probe(data, pct, alg);
        double t1 = now_s();
        t->kNN(q, K);
        double t2 = now_s();
        t->dump_knn(argv[10], (int)q.nrows(), K, t1 - t0, t2 - t1);
        if (argc == 12) t->dump_tree(argv[11]);
        fprintf(stderr, "ref_knn: n=%u dim=%u q=%u K=%d alg=%d build=%.3fs knn=%.3fs\n", data.nrows(), data.ncols(),
                q.nrows(), K, alg, t1 - t0, t2 - t1);
        fflush(NULL); _Exit(0);
    }
    return usage();
}
