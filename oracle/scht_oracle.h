/* CPU ORACLE interface — TEST INFRASTRUCTURE, NOT PRODUCT CODE (see scht_oracle.c). */
#ifndef SCHT_ORACLE_H
#define SCHT_ORACLE_H
#include <stdint.h>

#include "../include/scht.h"
#ifdef __cplusplus
extern "C" {
#endif
/* alg: 2 = by-leaf (USE_CPU_LEAF_APPRXIMATE_QP), 0 = recursive (USE_CPU_RECURSIVE_APPRXIMATE_QP).
 * counters[4] = { dist computations incl. 2/level descent (dist_computation_times_in_host),
 *                 hyper-plane tests (quadprog_cal_num), dot products (scalar_product_num),
 *                 scanned points S = leaf-scan distance computations only }.
 * approx_node_out: NODE index of the approximate leaf (CPU_Convex_Tree.h:215). */
int scht_oracle_knn(const scht_tree_desc* t, const float* queries, int64_t nq, int K, int alg, int* idx_out,
                    float* dist2_out, float* dist_out, int* approx_node_out, int64_t* counters);
int scht_oracle_brute_force(const float* data, int64_t n, int dim, const float* queries, int64_t nq, int K,
                            int* idx_out, float* dist2_out, float* dist_out);
int scht_oracle_contract(const scht_tree_desc* t, const float* queries, int64_t nq, int K, int* idx_out,
                         float* dist2_out);
#ifdef __cplusplus
}
#endif
#endif
