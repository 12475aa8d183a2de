// DROP-IN for the reference tree: B200_Convex_Tree<T> : Convex_Tree<T>.
//
// Add this header next to Code/GPU_Convex_Tree.h in XiaoLHu/Semi-convex_Hull_Tree, include it from Code/main.cc after
// CPU_Convex_Tree.h, link libscht_b200.so, and add the factory line shown in INTEGRATION.md.  The tree is built on the
// host by the reference's own constructor (Code/Convex_Tree.h:518-548); this class only flattens `nodes` the way
// GPU_Convex_Tree<T>::init_openCL_sorted_data_set does (Code/GPU_Convex_Tree.h:522-748) and overrides the virtual hooks
// the reference defines for back-ends (Code/Convex_Tree.h:231-272).  It needs the reference's headers and therefore
// compiles only inside the reference tree (or with -I/root/reference/Code, which is how oracle/ref_b200_harness.cc
// exercises it).
#ifndef B200_CONVEX_TREE_REF_H
#define B200_CONVEX_TREE_REF_H

#include <cstring>
#include <iostream>
#include <vector>

#include <string>
#include "Convex_Tree.h"
#include "scht.h"

#define USE_B200_LEAF_APPRXIMATE_QP 7  // next to the alg-type macros of Code/Convex_Tree.h:22-33

template <typename T>
class B200_Convex_Tree : public Convex_Tree<T> {
public:
    // device >= 0: that GPU; device == -1: every visible GPU, the queries of a kNN() call are sharded across them
    B200_Convex_Tree(Matrix<T>& data, FLOAT_TYPE leaf_pts_percent, int alg_type = USE_B200_LEAF_APPRXIMATE_QP, int device = 0)
        : Convex_Tree<T>(data, leaf_pts_percent, alg_type) {
        flatten_and_upload(std::vector<int>(1, device), SCHT_SHARD_QUERIES);
    }
    // several GPUs behind the reference's interface: `devices` = CUDA ordinals (empty = all), mode = SCHT_SHARD_QUERIES (tree
    // replicated, query slices) or SCHT_SHARD_DATASET (each GPU holds 1/n of the leaves; K-lists merged over NVLink)
    B200_Convex_Tree(Matrix<T>& data, FLOAT_TYPE leaf_pts_percent, const std::vector<int>& devices, int mode, int alg_type = USE_B200_LEAF_APPRXIMATE_QP)
        : Convex_Tree<T>(data, leaf_pts_percent, alg_type) {
        flatten_and_upload(devices, mode);
    }
    virtual ~B200_Convex_Tree() { scht_destroy(dev_); }

    virtual FLOAT_TYPE* get_kNN_dists(int i) { return dist_.data() + (size_t)i * this->K; }
    virtual FLOAT_TYPE* get_kNN_dists_squre(int i) { return dist2_.data() + (size_t)i * this->K; }
    virtual int* get_kNN_indexes(int i) { return idx_.data() + (size_t)i * this->K; }
    virtual void set_batch_size(int batch_size) {
        this->BATCH_SIZE_OF_QUERY_POINTS_PUSCH_TO_DEVICE = batch_size;
        if (dev_ && batch_size > 0) scht_set_batch_size(dev_, batch_size);
    }
    // the reference's log block (Code/GPU_Convex_Tree.h:950-1005), formatted by the library from this call's counters
    virtual void print_kNN_running_time_info() { std::cout << knn_log_text(); }
    virtual void save_KNN_running_time_info(FILE* log_file) { write_file_log(knn_log_text().c_str(), log_file); }
    std::string knn_log_text() const {
        size_t need = 0;
        const long long nq = (long long)this->query_points_vec.size();
        scht_format_knn_log(&stats_, nq, this->K, this->BATCH_SIZE_OF_QUERY_POINTS_PUSCH_TO_DEVICE, nullptr, 0, &need);
        std::string s(need, '\0');
        scht_format_knn_log(&stats_, nq, this->K, this->BATCH_SIZE_OF_QUERY_POINTS_PUSCH_TO_DEVICE, &s[0], need, nullptr);
        s.resize(need ? need - 1 : 0);
        return s;
    }
    const scht_stats& stats() const { return stats_; }

protected:
    virtual void init_kNN_result() {
        size_t n = (size_t)this->query_points_vec.size() * this->K;
        idx_.assign(n, -1);
        dist2_.assign(n, INFINITE);
        dist_.assign(n, INFINITE);
    }
    virtual void do_kNN(Matrix<T>& q) { run(q, false); }
    virtual void do_brute_force_kNN(Matrix<T>& q) { run(q, true); }

private:
    void run(Matrix<T>& q, bool brute) {
        int rc = (brute ? scht_brute_force : scht_query)(dev_, q.get_matrix_raw_data(), q.nrows(), q.ncols(), this->K, idx_.data(),
                                                        dist2_.data(), dist_.data(), &stats_);
        if (rc != SCHT_OK) std::cout << "B200_Convex_Tree: " << scht_last_error() << std::endl;  // like check_cl_error: print, go on
    }

    void flatten_and_upload(const std::vector<int>& devices, int mode) {
        const int dim = this->m_data.ncols(), n = this->m_data.nrows(), nn = this->nodes_num, L = this->leaf_nodes_num;
        std::vector<int> leaf_start(L), leaf_count(L), leaf_node(L), left(nn), right(nn), parent(nn), leaf_index(nn), beta_off(nn + 1);
        std::vector<float> lc((size_t)nn * dim, 0.f), rc((size_t)nn * dim, 0.f), normal((size_t)nn * dim, 0.f), beta;
        for (int i = 0; i < L; i++) {
            leaf_start[i] = this->leaf_nodes_start_pos_in_data_set[i];
            leaf_node[i] = this->leaf_nodes_ori_indexes[i];
            leaf_count[i] = this->nodes[leaf_node[i]]->pts_number;
        }
        for (int i = 0; i < nn; i++) {
            ConvexNode<T>* nd = this->nodes[i];
            left[i] = nd->left;
            right[i] = nd->right;
            parent[i] = nd->parent_index;
            leaf_index[i] = nd->isLeaf ? nd->leaf_index : -1;
            if (!nd->isLeaf)
                for (int j = 0; j < dim; j++) {
                    lc[(size_t)i * dim + j] = nd->left_center[j];
                    rc[(size_t)i * dim + j] = nd->right_center[j];
                }
            int c = (i == 0) ? 0 : (int)nd->ALPHA.nrows();
            beta_off[i] = (int)beta.size();
            for (int k = 0; k < c; k++) beta.push_back(nd->BETA[k]);
            if (c > 0)
                for (int j = 0; j < dim; j++) normal[(size_t)i * dim + j] = nd->ALPHA[c - 1][j];
        }
        beta_off[nn] = (int)beta.size();
        if (beta.empty()) beta.push_back(0.f);
        scht_tree_desc d;
        memset(&d, 0, sizeof(d));
        d.dim = dim; d.n_points = n; d.n_nodes = nn; d.n_leaves = L; d.depth = this->depth; d.leaf_threshold = this->leaf_pts_num_threshold;
        d.sorted_points = this->m_sorted_data.get_matrix_raw_data();
        d.sorted_to_orig = this->m_sorted_data_ori_indexes.get_matrix_raw_data();
        d.leaf_start = leaf_start.data(); d.leaf_count = leaf_count.data(); d.leaf_node_id = leaf_node.data();
        d.node_left = left.data(); d.node_right = right.data(); d.node_parent = parent.data(); d.node_leaf_index = leaf_index.data();
        d.node_left_center = lc.data(); d.node_right_center = rc.data(); d.node_normal = normal.data();
        d.node_beta_off = beta_off.data(); d.node_beta = beta.data();
        const bool one = devices.size() == 1 && mode == SCHT_SHARD_QUERIES;
        int status = one ? scht_create(&d, devices[0], &dev_) : scht_create_multi(&d, devices.empty() ? 0 : devices.data(), (int)devices.size(), mode, &dev_);
        if (status != SCHT_OK) {
            std::cout << "B200_Convex_Tree: " << scht_last_error() << std::endl;
            dev_ = 0;
        }
    }

    scht_handle* dev_ = 0;
    scht_stats stats_;
    std::vector<int> idx_;
    std::vector<float> dist2_, dist_;
};
#endif
