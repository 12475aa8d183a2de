/*
 * scht_oracle.c — CPU ORACLE.  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C restatement of the reference's CPU k-NN query path (XiaoLHu/Semi-convex_Hull_Tree,
 * Code/CPU_Convex_Tree.h) over the flat tree arrays of include/scht.h.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it; the
 * product (libscht_b200.so) never links, loads or calls anything in oracle/.
 *
 * PARITY PINNING: pinned against the reference itself — tests/test_oracle_golden.py runs
 * oracle/_ref/ref_knn (the unmodified reference compiled from /root/reference, oracle/Makefile)
 * on the same inputs and requires bit-identical indices, dist2, dist, approximate leaves and
 * work counters; tests/golden/ holds committed fixtures produced by that binary
 * (tests/golden/make_golden.py).  The reference ships no golden vectors of its own.
 *
 * Build with: gcc -O2 -ffp-contract=off -fopenmp (no -ffast-math): the reference is x86-64 SSE2
 * code with separately rounded mul/add.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/scht.h"
#include "scht_oracle.h"

#define ORACLE_INFINITE 3.402823466e+38f /* INFINITE, Code/cyw_types.h:18 */

/* result[i] += pow(x, 2): float accumulator, square and sum evaluated in double
 * (Code/basic_functions.h:73,177,194; pow(float,int) promotes to double, GCC folds to x*x). */
static inline float acc_sq(float acc, float d) { return (float)((double)acc + (double)d * (double)d); }

/* pdist2_squre(Vector a, Vector b) via norm_v_squre — Code/basic_functions.h:159-164,69-75 */
static float dist2_vec(const float* a, const float* b, int dim) {
    float acc = 0.f;
    for (int j = 0; j < dim; j++) acc = acc_sq(acc, a[j] - b[j]);
    return acc;
}

/* NNResult<T>::update — Code/Convex_Tree.h:121-151 */
static void nn_update(float* d2, float* d, int* idx, int K, float dist_square, int index) {
    if (d2[K - 1] > dist_square) {
        d2[K - 1] = dist_square;
        d[K - 1] = sqrtf(dist_square);
        idx[K - 1] = index;
        for (int j = K - 1; j > 0; j--) {
            if (d2[j - 1] > d2[j]) {
                float t = d2[j - 1]; d2[j - 1] = d2[j]; d2[j] = t;
                t = d[j - 1]; d[j - 1] = d[j]; d[j] = t;
                int ti = idx[j - 1]; idx[j - 1] = idx[j]; idx[j] = ti;
            }
        }
    }
}

/* is_appr_min_dist_from_q_larger_by_hyper_plane — Code/ConvexNode.h:113-144.
 * path[0..c-1] are the nodes on the root->node chain (root excluded); row i of the node's ALPHA is
 * node_normal[path[i]] (Code/Convex_Tree.h:755-771). Returns 1 when the node can be skipped. */
static int hyperplane_prune(const scht_tree_desc* t, const int* path, int c, const float* beta, const float* q,
                            float dist_compare, int64_t* n_products) {
    int dim = t->dim;
    for (int i = 0; i < c; i++) {
        const float* alpha = t->node_normal + (size_t)path[i] * dim;
        float tmp = beta[i];
        for (int j = 0; j < dim; j++) {
            float prod = alpha[j] * q[j];
            tmp = tmp + prod;
        }
        (*n_products)++;
        if (tmp < 0) {
            if (dist_compare < fabsf(tmp)) return 1;
        }
    }
    return 0;
}

/* is_appr_min_dist_from_q_larger_by_hyper_plane_pro — Code/ConvexNode.h:151-206: the dot product of q with
 * each edge normal is computed once per query (memo keyed by the path node id; 0 means "not yet computed",
 * :180) starting from 0, and beta is added afterwards (:172,190,193) — a different rounding from the
 * non-memoised test above, reproduced here. */
static int hyperplane_prune_pro(const scht_tree_desc* t, const int* path, int c, const float* beta, const float* q,
                                float dist_compare, float* memo, int64_t* n_products) {
    int dim = t->dim;
    for (int i = 0; i < c; i++) {
        float tmp = beta[i];
        int key = path[i];
        if (memo[key] == 0) {
            const float* alpha = t->node_normal + (size_t)key * dim;
            float prod = 0;
            for (int j = 0; j < dim; j++) {
                float m = alpha[j] * q[j];
                prod = prod + m;
            }
            (*n_products)++;
            memo[key] = prod;
            tmp = tmp + prod;
        } else {
            tmp = tmp + memo[key];
        }
        if (tmp < 0) {
            if (dist_compare < fabsf(tmp)) return 1;
        }
    }
    return 0;
}

static int node_path(const scht_tree_desc* t, int node, int* path, int cap) {
    int c = t->node_beta_off[node + 1] - t->node_beta_off[node];
    if (c > cap) return -1;
    int cur = node;
    for (int i = c - 1; i >= 0; i--) { path[i] = cur; cur = t->node_parent[cur]; }
    return c;
}

/* scan one leaf: pdist2_squre (Code/basic_functions.h:167-198) + update loop
 * (Code/CPU_Convex_Tree.h:209-212,420-423) */
static void scan_leaf(const scht_tree_desc* t, int leaf, const float* q, float* d2, float* d, int* idx, int K) {
    int dim = t->dim;
    int start = t->leaf_start[leaf], cnt = t->leaf_count[leaf];
    for (int i = 0; i < cnt; i++) {
        const float* p = t->sorted_points + (size_t)(start + i) * dim;
        float acc = 0.f;
        for (int j = 0; j < dim; j++) acc = acc_sq(acc, p[j] - q[j]);
        nn_update(d2, d, idx, K, acc, t->sorted_to_orig[start + i]);
    }
}

#define MAX_DEPTH 4096

static void walk_recursive(const scht_tree_desc* t, int node, int approx_node, const float* q, int K, float* d2,
                           float* d, int* idx, float* memo, int64_t* cnt);


/* one query of NN_a_point_by_leaf (alg 2, Code/CPU_Convex_Tree.h:380-430) or of
 * do_kNN_recurisive_by_loop (alg 0, :294-376); both start with approximate_kNN (:181-217). */
static void query_one(const scht_tree_desc* t, int alg, const float* q, int K, float* d2, float* d, int* idx,
                      int* approx_node_out, int64_t* cnt /* [4] */) {
    int dim = t->dim;
    for (int k = 0; k < K; k++) { d2[k] = ORACLE_INFINITE; d[k] = ORACLE_INFINITE; idx[k] = -1; }
    /* approximate_kNN: descend by nearer child centre; ties go right (:191) */
    int cur = 0;
    while (t->node_leaf_index[cur] < 0) {
        float l = dist2_vec(q, t->node_left_center + (size_t)cur * dim, dim);
        float r = dist2_vec(q, t->node_right_center + (size_t)cur * dim, dim);
        cnt[0] += 2;
        cur = (l >= r) ? t->node_right[cur] : t->node_left[cur];
    }
    int approx_node = cur;
    *approx_node_out = approx_node;
    scan_leaf(t, t->node_leaf_index[approx_node], q, d2, d, idx, K);
    cnt[0] += t->leaf_count[t->node_leaf_index[approx_node]];
    cnt[3] += t->leaf_count[t->node_leaf_index[approx_node]];

    int path[MAX_DEPTH];
    if (alg == 2) {
        for (int leaf = 0; leaf < t->n_leaves; leaf++) {
            int node = t->leaf_node_id[leaf];
            if (node == approx_node) continue;
            int c = node_path(t, node, path, MAX_DEPTH);
            float kth = d[K - 1];
            cnt[1]++;
            if (!hyperplane_prune(t, path, c, t->node_beta + t->node_beta_off[node], q, kth, &cnt[2])) {
                scan_leaf(t, leaf, q, d2, d, idx, K);
                cnt[0] += t->leaf_count[leaf];
                cnt[3] += t->leaf_count[leaf];
            }
        }
    } else {
        float* memo = (float*)calloc((size_t)t->n_nodes, sizeof(float));
        walk_recursive(t, 0, approx_node, q, K, d2, d, idx, memo, cnt);
        free(memo);
    }
}

/* do_kNN_recurisive_by_loop (Code/CPU_Convex_Tree.h:294-376) == do_kNN_recurisive (:220-290): a child is
 * tested against the k-th distance current at the moment it is about to be entered, left before right. */
static void walk_recursive(const scht_tree_desc* t, int node, int approx_node, const float* q, int K, float* d2,
                           float* d, int* idx, float* memo, int64_t* cnt) {
    if (t->node_leaf_index[node] >= 0) {
        if (node != approx_node) {
            int leaf = t->node_leaf_index[node];
            scan_leaf(t, leaf, q, d2, d, idx, K);
            cnt[0] += t->leaf_count[leaf];
            cnt[3] += t->leaf_count[leaf];
        }
        return;
    }
    int path[MAX_DEPTH];
    for (int side = 0; side < 2; side++) {
        int child = side == 0 ? t->node_left[node] : t->node_right[node];
        int c = node_path(t, child, path, MAX_DEPTH);
        cnt[1]++;
        if (!hyperplane_prune_pro(t, path, c, t->node_beta + t->node_beta_off[child], q, d[K - 1], memo, &cnt[2]))
            walk_recursive(t, child, approx_node, q, K, d2, d, idx, memo, cnt);
    }
}

int scht_oracle_knn(const scht_tree_desc* t, const float* queries, int64_t nq, int K, int alg, int* idx_out,
                    float* dist2_out, float* dist_out, int* approx_node_out, int64_t* counters) {
    if (!t || !queries || K < 1 || (alg != 0 && alg != 2)) return -1;
    int64_t c0 = 0, c1 = 0, c2 = 0, c3 = 0;
    int dim = t->dim;
#pragma omp parallel for schedule(dynamic, 8) reduction(+ : c0, c1, c2, c3)
    for (int64_t i = 0; i < nq; i++) {
        int64_t cnt[4] = {0, 0, 0, 0};
        int an;
        float* dtmp = (float*)malloc(sizeof(float) * K);
        query_one(t, alg, queries + (size_t)i * dim, K, dist2_out + (size_t)i * K, dist_out ? dist_out + (size_t)i * K : dtmp,
                  idx_out + (size_t)i * K, &an, cnt);
        free(dtmp);
        if (approx_node_out) approx_node_out[i] = an;
        c0 += cnt[0]; c1 += cnt[1]; c2 += cnt[2]; c3 += cnt[3];
    }
    if (counters) { counters[0] = c0; counters[1] = c1; counters[2] = c2; counters[3] = c3; }
    return 0;
}

/* brute_force_check's search (Code/main.cc:59-68): every row in ORIGINAL order through update(). */
int scht_oracle_brute_force(const float* data, int64_t n, int dim, const float* queries, int64_t nq, int K,
                            int* idx_out, float* dist2_out, float* dist_out) {
    if (!data || !queries || K < 1) return -1;
#pragma omp parallel for schedule(dynamic, 8)
    for (int64_t i = 0; i < nq; i++) {
        const float* q = queries + (size_t)i * dim;
        float* d2 = dist2_out + (size_t)i * K;
        int* idx = idx_out + (size_t)i * K;
        float* dtmp = (float*)malloc(sizeof(float) * K);
        float* d = dist_out ? dist_out + (size_t)i * K : dtmp;
        for (int k = 0; k < K; k++) { d2[k] = ORACLE_INFINITE; d[k] = ORACLE_INFINITE; idx[k] = -1; }
        for (int64_t r = 0; r < n; r++) {
            const float* p = data + (size_t)r * dim;
            float acc = 0.f;
            for (int j = 0; j < dim; j++) acc = acc_sq(acc, p[j] - q[j]);
            nn_update(d2, d, idx, K, acc, (int)r);
        }
        free(dtmp);
    }
    return 0;
}

/* The order-independent RESULT CONTRACT: the K smallest
 * (dist2, r, pos) triples, r = 0 for points of the query's approximate leaf else 1, pos = row in
 * sorted_points — computed by exhaustive scan, no pruning.  Independent second opinion on
 * scht_oracle_knn: wherever the reference's (un-margined) pruning is exact the two agree bit for bit. */
int scht_oracle_contract(const scht_tree_desc* t, const float* queries, int64_t nq, int K, int* idx_out,
                         float* dist2_out) {
    if (!t || !queries || K < 1) return -1;
    int dim = t->dim;
#pragma omp parallel for schedule(dynamic, 8)
    for (int64_t i = 0; i < nq; i++) {
        const float* q = queries + (size_t)i * dim;
        int cur = 0;
        while (t->node_leaf_index[cur] < 0) {
            float l = dist2_vec(q, t->node_left_center + (size_t)cur * dim, dim);
            float r = dist2_vec(q, t->node_right_center + (size_t)cur * dim, dim);
            cur = (l >= r) ? t->node_right[cur] : t->node_left[cur];
        }
        int al = t->node_leaf_index[cur];
        int64_t a0 = t->leaf_start[al], a1 = a0 + t->leaf_count[al];
        uint64_t* keys = (uint64_t*)malloc(sizeof(uint64_t) * K);
        int filled = 0;
        for (int64_t pos = 0; pos < t->n_points; pos++) {
            const float* p = t->sorted_points + (size_t)pos * dim;
            float acc = 0.f;
            for (int j = 0; j < dim; j++) acc = acc_sq(acc, p[j] - q[j]);
            if (!(acc < ORACLE_INFINITE)) continue; /* update() needs INFINITE > dist2 */
            uint32_t bits; memcpy(&bits, &acc, 4);
            uint64_t key = ((uint64_t)bits << 32) | ((pos >= a0 && pos < a1) ? 0u : 0x80000000u) | (uint32_t)pos;
            if (filled == K && key >= keys[K - 1]) continue;
            int j = filled < K ? filled++ : K - 1;
            while (j > 0 && keys[j - 1] > key) { keys[j] = keys[j - 1]; j--; }
            keys[j] = key;
        }
        for (int k = 0; k < K; k++) {
            if (k < filled) {
                uint32_t bits = (uint32_t)(keys[k] >> 32); float f; memcpy(&f, &bits, 4);
                dist2_out[(size_t)i * K + k] = f;
                idx_out[(size_t)i * K + k] = t->sorted_to_orig[keys[k] & 0x7fffffffu];
            } else { dist2_out[(size_t)i * K + k] = ORACLE_INFINITE; idx_out[(size_t)i * K + k] = -1; }
        }
        free(keys);
    }
    return 0;
}
