// Host-side C++ mirror of the reference's tree interface for the batched exact k-NN path, on top of the C ABI
// (include/scht.h).  Same method names, argument meaning and error behaviour as the reference's
// Convex_Tree<T> / GPU_Convex_Tree<T> (reference: Code/Convex_Tree.h:201-272,456-486, Code/GPU_Convex_Tree.h:28-42,
// 197-204,950-1005), so drivers written against the reference (Code/main.cc:300-356) read the same:
//
//     scht::Matrix<float> data(n, dim), Q(nq, dim);  ... fill ...
//     scht::B200_Convex_Tree<float> tree(data, 0.0005f);      // throws std::runtime_error on a bad percent
//     tree.set_batch_size(1000000);
//     tree.kNN(Q, 10);
//     int*   idx = tree.get_kNN_indexes(i);                   // K original point indices of query i
//     float* d2  = tree.get_kNN_dists_squre(i);               // K ascending squared distances
//
// This header is standalone (it does not need the reference's sources); integration/B200_Convex_Tree.h is the
// variant that derives from the reference's own Convex_Tree<T> and reuses the reference's host build.
// Result arrays are owned by the tree and stay valid until the next kNN()/brute_force_kNN() or destruction
// (like Code/GPU_Convex_Tree.h:19-20,937-946).  Not thread-safe (neither is the reference).
#ifndef SCHT_B200_CONVEX_TREE_H
#define SCHT_B200_CONVEX_TREE_H

#include <chrono>
#include <cstdio>
#include <iostream>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>

#include "../../include/scht.h"

namespace scht {

// Row-major dense matrix with the accessors the reference's Matrix<T> offers on this path
// (Code/Array.hh:1175-1183 storage, :1908-1910 get_matrix_raw_data).
template <typename T>
class Matrix {
public:
    Matrix() : n_(0), m_(0) {}
    Matrix(unsigned n, unsigned m) : n_(n), m_(m), v_((size_t)n * m) {}
    unsigned nrows() const { return n_; }
    unsigned ncols() const { return m_; }
    T* operator[](unsigned i) { return v_.data() + (size_t)i * m_; }
    const T* operator[](unsigned i) const { return v_.data() + (size_t)i * m_; }
    T* get_matrix_raw_data() { return v_.data(); }
    const T* get_matrix_raw_data() const { return v_.data(); }

private:
    unsigned n_, m_;
    std::vector<T> v_;
};

// algorithm ids next to the reference's (Code/Convex_Tree.h:22-33); this back-end has one algorithm.
#ifndef USE_B200_LEAF_APPRXIMATE_QP
#define USE_B200_LEAF_APPRXIMATE_QP 7
#endif

template <typename T>
class B200_Convex_Tree {
    static_assert(std::is_same<T, float>::value, "FLOAT_TYPE is float (Code/cyw_types.h:5-16)");

public:
    // ctor == Convex_Tree<T>::Convex_Tree(Matrix<T>&, FLOAT_TYPE, int) + GPU_Convex_Tree ctor tail
    // (Code/Convex_Tree.h:518-548, Code/GPU_Convex_Tree.h:150-163).  rand_seed 1 == srand never called.
    B200_Convex_Tree(Matrix<T>& data, float leaf_pts_percent, int alg_type = USE_B200_LEAF_APPRXIMATE_QP, int device = 0,
                     unsigned rand_seed = 1)
        : ALG_TYPE(alg_type), dim_((int)data.ncols()), pct_(leaf_pts_percent) {
        if (leaf_pts_percent <= 0 || leaf_pts_percent >= 0.5f)  // Code/Convex_Tree.h:522-523
            throw std::runtime_error("illegal percent, it should be in (0,0.5)");
        if (scht_build_tree(data.get_matrix_raw_data(), data.nrows(), (int32_t)data.ncols(), leaf_pts_percent, rand_seed, &host_) != SCHT_OK)
            throw std::runtime_error(std::string("scht_build_tree: ") + scht_last_error());
        if (scht_create(scht_host_tree_desc(host_), device, &dev_) != SCHT_OK) {
            std::string msg = scht_last_error();
            scht_host_tree_free(host_);
            host_ = nullptr;
            throw std::runtime_error("scht_create: " + msg);  // no GPU => no tree: this back-end has no CPU fallback
        }
    }
    // several GPUs behind the same interface (scht_create_multi): `devices` = CUDA ordinals (empty = every visible GPU),
    // mode = SCHT_SHARD_QUERIES (tree replicated, each GPU answers a slice of the queries of a call) or SCHT_SHARD_DATASET
    // (each GPU holds 1/n of the leaves and scans it for all queries; the K-lists are merged over NVLink)
    B200_Convex_Tree(Matrix<T>& data, float leaf_pts_percent, const std::vector<int>& devices, int mode,
                     int alg_type = USE_B200_LEAF_APPRXIMATE_QP, unsigned rand_seed = 1)
        : ALG_TYPE(alg_type), dim_((int)data.ncols()), pct_(leaf_pts_percent) {
        if (leaf_pts_percent <= 0 || leaf_pts_percent >= 0.5f) throw std::runtime_error("illegal percent, it should be in (0,0.5)");
        if (scht_build_tree(data.get_matrix_raw_data(), data.nrows(), (int32_t)data.ncols(), leaf_pts_percent, rand_seed, &host_) != SCHT_OK)
            throw std::runtime_error(std::string("scht_build_tree: ") + scht_last_error());
        if (scht_create_multi(scht_host_tree_desc(host_), devices.empty() ? nullptr : devices.data(), (int32_t)devices.size(), mode, &dev_) != SCHT_OK) {
            std::string msg = scht_last_error();
            scht_host_tree_free(host_);
            host_ = nullptr;
            throw std::runtime_error("scht_create_multi: " + msg);
        }
    }
    ~B200_Convex_Tree() {
        scht_destroy(dev_);
        scht_host_tree_free(host_);
    }
    B200_Convex_Tree(const B200_Convex_Tree&) = delete;
    B200_Convex_Tree& operator=(const B200_Convex_Tree&) = delete;

    // the entry for perform kNN (Code/Convex_Tree.h:456-469)
    void kNN(Matrix<T>& query_points_mat, int K) { run(query_points_mat, K, false); }
    // Code/Convex_Tree.h:473-486
    void brute_force_kNN(Matrix<T>& query_points_mat, int K) { run(query_points_mat, K, true); }

    // after kNN processing, get the result by the following three entries (Code/Convex_Tree.h:231-238)
    float* get_kNN_dists(int query_point_index) { return dist_.data() + (size_t)query_point_index * K_; }
    float* get_kNN_dists_squre(int query_point_index) { return dist2_.data() + (size_t)query_point_index * K_; }
    int* get_kNN_indexes(int query_point_index) { return idx_.data() + (size_t)query_point_index * K_; }

    // Code/GPU_Convex_Tree.h:197-204
    void set_batch_size(int batch_size) {
        if (batch_size > 0 && scht_set_batch_size(dev_, batch_size) == SCHT_OK) batch_size_ = batch_size;
    }

    int get_nodes_num() const { return scht_host_tree_desc(host_)->n_nodes; }
    int get_leaf_nodes_num() const { return scht_host_tree_desc(host_)->n_leaves; }
    const scht_stats& stats() const { return stats_; }

    // Code/Convex_Tree.h:946-958 / :964-991 -- the reference's text block, formatted by the library (scht_format_tree_info)
    void print_tree_info() { std::cout << tree_info_text(); }
    void save_tree_info(FILE* log_file) { write_log(tree_info_text(), log_file); }
    // Code/GPU_Convex_Tree.h:950-971 / :973-1005 (scht_format_knn_log)
    void print_kNN_running_time_info() { std::cout << knn_log_text(); }
    void save_KNN_running_time_info(FILE* log_file) { write_log(knn_log_text(), log_file); }
    std::string tree_info_text() const {
        size_t need = 0;
        scht_format_tree_info(host_, pct_, nullptr, 0, &need);
        std::string s(need, '\0');
        scht_format_tree_info(host_, pct_, &s[0], need, nullptr);
        s.resize(need ? need - 1 : 0);
        return s;
    }
    std::string knn_log_text() const {
        size_t need = 0;
        scht_format_knn_log(&stats_, n_queries_, K_, batch_size_, nullptr, 0, &need);
        std::string s(need, '\0');
        scht_format_knn_log(&stats_, n_queries_, K_, batch_size_, &s[0], need, nullptr);
        s.resize(need ? need - 1 : 0);
        return s;
    }

    int ALG_TYPE;

private:
    void run(Matrix<T>& q, int K, bool brute) {
        int dim = (int)q.ncols();
        if (dim != dim_) {  // Code/Convex_Tree.h:458-461: message, then return without results
            std::cout << "the dim of query points is " << dim << ", while the dim of data set is " << dim_;
            return;
        }
        K_ = K;
        n_queries_ = q.nrows();
        size_t n = (size_t)q.nrows() * (size_t)(K > 0 ? K : 0);
        idx_.assign(n, -1);
        dist2_.assign(n, 3.402823466e+38f);
        dist_.assign(n, 3.402823466e+38f);
        int rc = (brute ? scht_brute_force : scht_query)(dev_, q.get_matrix_raw_data(), q.nrows(), dim, K, idx_.data(), dist2_.data(),
                                                        dist_.data(), &stats_);
        // the reference prints device errors and carries on (Code/openCL_tool.h:245-254); we print and leave the
        // result arrays at their "nothing found" values instead of returning garbage
        if (rc != SCHT_OK) std::cout << "B200_Convex_Tree: " << scht_last_error() << std::endl;
    }

    static void write_log(const std::string& text, FILE* fp) {  // write_file_log, Code/basic_functions.h:338-348
        if (!fp) {
            std::cout << "can't open file\n";
            return;
        }
        fseek(fp, 0, SEEK_END);
        fwrite(text.data(), text.size(), 1, fp);
    }

    int dim_, K_ = 0;
    float pct_;
    long long n_queries_ = 0, batch_size_ = 1000000;  // default batch: Code/main.cc:670
    scht_host_tree* host_ = nullptr;
    scht_handle* dev_ = nullptr;
    scht_stats stats_{};
    std::vector<int> idx_;
    std::vector<float> dist2_, dist_;
};

}  // namespace scht
#endif
