// Exercises the C++ host mirror (semi-convex_hull_tree_b200/host/B200_Convex_Tree.h) the way Code/main.cc drives a tree:
// create, set_batch_size, kNN, get_kNN_*; plus the reference's error behaviour.  Needs a B200 (exit 77 = no GPU).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>

#include "../../semi-convex_hull_tree_b200/host/B200_Convex_Tree.h"

#define CHECK(c)                                                   \
    do {                                                           \
        if (!(c)) {                                                \
            fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #c); \
            return 1;                                              \
        }                                                          \
    } while (0)

int main() {
    if (scht_device_count() == 0) {
        // no CPU fallback: constructing a tree must throw
        scht::Matrix<float> d(100, 4);
        for (unsigned i = 0; i < 100; i++) for (unsigned j = 0; j < 4; j++) d[i][j] = (float)((i * 7 + j * 13) % 101);
        bool threw = false;
        try { scht::B200_Convex_Tree<float> t(d, 0.1f); } catch (const std::runtime_error&) { threw = true; }
        CHECK(threw);
        printf("no GPU: constructor throws as designed\n");
        return 77;
    }
    const unsigned n = 30000, dim = 12, nq = 1000;
    const int K = 10;
    std::mt19937 rng(123);
    std::uniform_real_distribution<float> U(0.f, 1.f);
    scht::Matrix<float> data(n, dim), q(nq, dim);
    for (unsigned i = 0; i < n; i++) for (unsigned j = 0; j < dim; j++) data[i][j] = U(rng);
    for (unsigned i = 0; i < nq; i++) for (unsigned j = 0; j < dim; j++) q[i][j] = (i < 100) ? data[i][j] : U(rng);

    // bad percent: std::runtime_error (Code/Convex_Tree.h:522-523)
    for (float pct : {0.f, 0.5f, 0.9f}) {
        bool threw = false;
        try { scht::B200_Convex_Tree<float> t(data, pct); } catch (const std::runtime_error&) { threw = true; }
        CHECK(threw);
    }
    scht::B200_Convex_Tree<float> tree(data, 0.005f);
    tree.print_tree_info();
    tree.set_batch_size(300);  // 4 batches
    tree.kNN(q, K);
    CHECK(tree.stats().batches == 4 && tree.stats().n_queries == nq);
    std::vector<int> idx(tree.get_kNN_indexes(0), tree.get_kNN_indexes(0) + (size_t)nq * K);
    std::vector<float> d2(tree.get_kNN_dists_squre(0), tree.get_kNN_dists_squre(0) + (size_t)nq * K);
    for (unsigned i = 0; i < nq; i++) {
        for (int k = 0; k < K; k++) {
            int id = idx[(size_t)i * K + k];
            CHECK(id >= 0 && id < (int)n);
            float acc = 0.f;  // the reference's arithmetic (Code/basic_functions.h:167-180)
            for (unsigned j = 0; j < dim; j++) { float d = data[id][j] - q[i][j]; acc = (float)((double)acc + (double)d * (double)d); }
            CHECK(acc == d2[(size_t)i * K + k]);
            CHECK(tree.get_kNN_dists(i)[k] == sqrtf(acc));
            if (k) CHECK(d2[(size_t)i * K + k - 1] <= acc);
        }
        if (i < 100) CHECK(idx[(size_t)i * K] == (int)i && d2[(size_t)i * K] == 0.f);
    }
    // exhaustive entry agrees on tie-free data
    tree.set_batch_size(1000000);
    tree.brute_force_kNN(q, K);
    for (size_t i = 0; i < (size_t)nq * K; i++) CHECK(tree.get_kNN_indexes(0)[i] == idx[i] && tree.get_kNN_dists_squre(0)[i] == d2[i]);
    // dimension mismatch: message on stdout, silent return (Code/Convex_Tree.h:458-461)
    scht::Matrix<float> bad(5, dim + 1);
    tree.kNN(bad, K);
    printf("\n");
    tree.print_kNN_running_time_info();
    printf("host mirror ok\n");
    return 0;
}
