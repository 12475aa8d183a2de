// Host tree builder of the B200 Semi-convex Hull Tree back-end (pure C++, no CUDA).
//
// Restates — with the reference's exact float/double arithmetic and its libc rand() stream — the
// constructor path of Convex_Tree<T> (reference: Code/Convex_Tree.h:518-548):
//     build_tree (:708-857) -> build_simplified_tree (:685-705) -> sort_raw_data (:593-633)
// and emits the flat SoA arrays of scht_tree_desc (include/scht.h).  Given the same data and seed the
// arrays are bit-identical to what the reference's own tree holds (tests/test_oracle_golden.py checks
// this against oracle/_ref when the reference is present, and against tests/golden/ always).
//
// This file must be compiled WITHOUT floating-point contraction (-ffp-contract=off) and without
// -ffast-math: the reference is plain x86-64 SSE2 code (mul and add rounded separately).
#include "../../include/scht.h"

#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "scht_error.h"
#include "scht_host_tree.h"

namespace {

// glibc's rand()/srand() (TYPE_3 additive feedback generator r[i] = r[i-3] + r[i-31]), re-implemented
// so the tree does not depend on the process-global libc state.  The reference draws its split seeds
// with rand() (Code/Convex_Tree.h:789).
class GlibcRand {
public:
    explicit GlibcRand(uint32_t seed) {
        if (seed == 0) seed = 1;
        int32_t r[34];
        r[0] = (int32_t)seed;
        for (int i = 1; i < 31; i++) {
            int64_t hi = r[i - 1] / 127773, lo = r[i - 1] % 127773;
            int64_t word = 16807 * lo - 2836 * hi;
            if (word < 0) word += 2147483647;
            r[i] = (int32_t)word;
        }
        for (int i = 31; i < 34; i++) r[i] = r[i - 31];
        for (int i = 0; i < 34; i++) ring_[i] = (uint32_t)r[i];
        pos_ = 34;
        for (int i = 34; i < 344; i++) step();
    }
    int next() { return (int)(step() >> 1); }

private:
    uint32_t step() {
        uint32_t v = ring_[(pos_ - 31) % 34] + ring_[(pos_ - 3) % 34];
        ring_[pos_ % 34] = v;
        pos_++;
        return v;
    }
    uint32_t ring_[34];
    uint64_t pos_;
};

// acc += pow(d, 2) with acc float and pow(float,int) evaluated in double
// (reference: Code/basic_functions.h:64,73,138,177; GCC folds pow(x,2) to x*x in double).
static inline float acc_sq(float acc, float d) { return (float)((double)acc + (double)d * (double)d); }

// scalar_product (Code/basic_functions.h:20-29): float products summed in a double, returned as float.
static inline float scalar_product(const float* x, const float* y, int n) {
    double sum = 0.0;
    for (int i = 0; i < n; i++) {
        float p = x[i] * y[i];
        sum += p;
    }
    return (float)sum;
}

struct Builder {
    const float* data;
    int64_t n;
    int dim;
    int threshold;
    GlibcRand rng;
    int depth_stat = 1;
    bool degenerate = false;

    // per node (creation order == DFS pre-order)
    std::vector<int32_t> left, right, parent, leaf_index, node_depth, beta_off;
    std::vector<float> left_center, right_center, normal, beta;
    std::vector<std::vector<int32_t>> leaf_points;  // by node id (empty for internal nodes)
    std::vector<int32_t> path;                      // current root->node chain, root excluded

    Builder(const float* d, int64_t n_, int dim_, int thr, uint32_t seed) : data(d), n(n_), dim(dim_), threshold(thr), rng(seed) {}

    const float* row(int64_t i) const { return data + i * dim; }

    // pdist2(Vector b, Matrix mat, Vector<int> indexes) — Code/basic_functions.h:129-143
    void dists_from(const float* b, const std::vector<int32_t>& idx, std::vector<float>& out) const {
        size_t m = idx.size();
        out.resize(m);
        for (size_t i = 0; i < m; i++) {
            const float* p = row(idx[i]);
            float acc = 0.f;
            for (int j = 0; j < dim; j++) acc = acc_sq(acc, p[j] - b[j]);
            out[i] = std::sqrt(acc);
        }
    }

    // getBeta_1 — Code/Convex_Tree.h:563-589: shift the hyper-plane to the node's extreme point.
    float tighten(const float* alpha, float beta0, const std::vector<int32_t>& idx) const {
        size_t m = idx.size();
        size_t best = 0;
        float best_v = 0.f;
        for (size_t i = 0; i < m; i++) {
            const float* p = row(idx[i]);
            float tmp = 0.f;
            for (int j = 0; j < dim; j++) {
                float prod = p[j] * alpha[j];
                tmp = tmp + prod;
            }
            float v = tmp + beta0;
            if (i == 0 || v < best_v) { best = i; best_v = v; }  // index_min: first minimum (Array.hh:907-916)
        }
        return -scalar_product(alpha, row(idx[best]), dim);
    }

    static size_t first_max(const std::vector<float>& v) {  // index_max (Array.hh:885-893)
        size_t mx = 0;
        for (size_t i = 1; i < v.size(); i++)
            if (v[i] > v[mx]) mx = i;
        return mx;
    }

    int new_node(int par) {
        int id = (int)left.size();
        left.push_back(-1); right.push_back(-1); parent.push_back(par); leaf_index.push_back(-1);
        node_depth.push_back(par < 0 ? 0 : node_depth[par] + 1);
        left_center.resize(left_center.size() + dim, 0.f);
        right_center.resize(right_center.size() + dim, 0.f);
        normal.resize(normal.size() + dim, 0.f);
        leaf_points.emplace_back();
        return id;
    }

    // build_tree — Code/Convex_Tree.h:708-857.  `alpha`/`beta_in` are the node's own (last) constraint.
    int build(int cur_depth, const float* alpha, float beta_in, int par, std::vector<int32_t>& idx) {
        int id = new_node(par);
        if (par >= 0) {
            // inherited constraints re-tightened to this node's points (:755-766), own constraint last (:767-771)
            int pc = node_depth[par];
            std::vector<float> my_beta(pc + 1);
            int pb = beta_off[par];
            for (int i = 0; i < pc; i++)
                my_beta[i] = tighten(&normal[(size_t)path[i] * dim], beta[pb + i], idx);
            my_beta[pc] = beta_in;
            memcpy(&normal[(size_t)id * dim], alpha, sizeof(float) * dim);
            beta_off.push_back((int32_t)beta.size());
            beta.insert(beta.end(), my_beta.begin(), my_beta.end());
            path.push_back(id);
        } else {
            beta_off.push_back(0);
        }

        int rows = (int)idx.size();
        bool make_leaf = rows <= threshold;  // :776
        std::vector<int32_t> l_idx, r_idx;
        std::vector<float> lc(dim), rc(dim);
        if (!make_leaf) {
            int pick = rng.next() % rows;  // :789
            std::vector<float> d0, dl, dr;
            dists_from(row(idx[pick]), idx, d0);
            size_t li = first_max(d0);  // :793-797
            memcpy(lc.data(), row(idx[li]), sizeof(float) * dim);
            dists_from(lc.data(), idx, dl);
            size_t ri = first_max(dl);  // :802-804
            memcpy(rc.data(), row(idx[ri]), sizeof(float) * dim);
            dists_from(rc.data(), idx, dr);
            for (int i = 0; i < rows; i++) {  // :809-821
                float diff = dl[i] - dr[i];
                if (diff <= 0.f) l_idx.push_back(idx[i]);
                if (diff > 0.f) r_idx.push_back(idx[i]);
            }
            // The reference recurses forever when one side is empty (more than `threshold` identical rows,
            // SURVEY §8d); we close the node as an oversized leaf instead.
            if (l_idx.empty() || r_idx.empty()) { make_leaf = true; degenerate = true; }
        }
        if (make_leaf) {
            leaf_points[id] = std::move(idx);
            if (par >= 0) path.pop_back();
            return id;
        }
        { std::vector<int32_t>().swap(idx); }

        std::vector<float> mean(dim), la(dim), ra(dim);
        for (int j = 0; j < dim; j++) {  // :826-828
            float s = lc[j];
            s += rc[j];
            s /= 2.f;
            mean[j] = s;
        }
        float nacc = 0.f;  // norm_v (Code/basic_functions.h:60-66)
        for (int j = 0; j < dim; j++) { la[j] = lc[j] - rc[j]; nacc = acc_sq(nacc, la[j]); }
        float nrm = std::sqrt(nacc);
        for (int j = 0; j < dim; j++) la[j] /= nrm;                 // :832
        for (int j = 0; j < dim; j++) ra[j] = (float)(-1) * la[j];  // :848, Array.hh:387-390

        // left (:835-843)
        float lbeta = tighten(la.data(), -scalar_product(la.data(), mean.data(), dim), l_idx);
        cur_depth++;
        if (cur_depth > depth_stat) depth_stat++;
        int lchild = build(cur_depth, la.data(), lbeta, id, l_idx);
        left[id] = lchild;
        memcpy(&left_center[(size_t)id * dim], lc.data(), sizeof(float) * dim);
        // right (:847-854)
        float rbeta = tighten(ra.data(), -scalar_product(ra.data(), mean.data(), dim), r_idx);
        int rchild = build(cur_depth, ra.data(), rbeta, id, r_idx);
        right[id] = rchild;
        memcpy(&right_center[(size_t)id * dim], rc.data(), sizeof(float) * dim);

        if (par >= 0) path.pop_back();
        return id;
    }
};

}  // namespace


extern "C" int scht_build_tree(const float* data, int64_t n_points, int32_t dim, float leaf_pts_percent,
                               uint32_t rand_seed, scht_host_tree** out) {
    if (!out) return scht::fail(SCHT_ERR_INVALID_ARG, "scht_build_tree: out is NULL");
    *out = nullptr;
    if (!data || n_points < 1 || n_points > 0x7fffffffLL || dim < 1)
        return scht::fail(SCHT_ERR_INVALID_ARG, "scht_build_tree: bad data/n_points/dim");
    // Code/Convex_Tree.h:522-523 throws std::runtime_error("illegal percent ...")
    if (!(leaf_pts_percent > 0.f) || leaf_pts_percent >= 0.5f)
        return scht::fail(SCHT_ERR_INVALID_ARG, "illegal percent, it should be in (0,0.5)");
    try {
        // (int)(data.nrows() * leaf_pts_percent) with unsigned*float -> float (Code/Convex_Tree.h:525)
        int thr = (int)((float)(unsigned)n_points * leaf_pts_percent);
        auto t0 = std::chrono::steady_clock::now();
        Builder b(data, n_points, dim, thr, rand_seed);
        std::vector<int32_t> all((size_t)n_points);
        for (int64_t i = 0; i < n_points; i++) all[(size_t)i] = (int32_t)i;
        b.build(1, nullptr, 0.f, -1, all);
        // build_simplified_tree (:685-705): leaf ids in node order
        int nn = (int)b.left.size();
        auto* t = new scht_host_tree();
        int L = 0;
        for (int i = 0; i < nn; i++)
            if (b.left[i] < 0 && b.right[i] < 0) b.leaf_index[i] = L++;
        auto t1 = std::chrono::steady_clock::now();
        // sort_raw_data (:593-633)
        t->sorted_points.resize((size_t)n_points * dim);
        t->sorted_to_orig.resize((size_t)n_points);
        t->leaf_start.resize(L); t->leaf_count.resize(L); t->leaf_node_id.resize(L);
        int64_t off = 0;
        for (int i = 0; i < nn; i++) {
            if (b.leaf_index[i] < 0) continue;
            int li = b.leaf_index[i];
            const std::vector<int32_t>& pts = b.leaf_points[i];
            t->leaf_start[li] = (int32_t)off; t->leaf_count[li] = (int32_t)pts.size(); t->leaf_node_id[li] = i;
            for (size_t k = 0; k < pts.size(); k++) {
                t->sorted_to_orig[(size_t)off + k] = pts[k];
                memcpy(&t->sorted_points[((size_t)off + k) * dim], data + (size_t)pts[k] * dim, sizeof(float) * dim);
            }
            off += (int64_t)pts.size();
        }
        if (off != n_points) { delete t; return scht::fail(SCHT_ERR_INTERNAL, "tree build lost points (NaN coordinates?)"); }
        b.beta_off.push_back((int32_t)b.beta.size());
        t->left.swap(b.left); t->right.swap(b.right); t->parent.swap(b.parent); t->leaf_index.swap(b.leaf_index);
        t->left_center.swap(b.left_center); t->right_center.swap(b.right_center); t->normal.swap(b.normal);
        t->beta.swap(b.beta); t->beta_off.swap(b.beta_off);
        if (t->beta.empty()) t->beta.push_back(0.f);  // keep the pointer non-NULL for single-leaf trees
        t->build_seconds = std::chrono::duration<double>(t1 - t0).count();
        scht_tree_desc& d = t->desc;
        d.dim = dim; d.n_points = n_points; d.n_nodes = nn; d.n_leaves = L; d.depth = b.depth_stat; d.leaf_threshold = thr;
        d.sorted_points = t->sorted_points.data(); d.sorted_to_orig = t->sorted_to_orig.data();
        d.leaf_start = t->leaf_start.data(); d.leaf_count = t->leaf_count.data(); d.leaf_node_id = t->leaf_node_id.data();
        d.node_left = t->left.data(); d.node_right = t->right.data(); d.node_parent = t->parent.data();
        d.node_leaf_index = t->leaf_index.data();
        d.node_left_center = t->left_center.data(); d.node_right_center = t->right_center.data();
        d.node_normal = t->normal.data(); d.node_beta_off = t->beta_off.data(); d.node_beta = t->beta.data();
        *out = t;
        return SCHT_OK;
    } catch (const std::bad_alloc&) {
        return scht::fail(SCHT_ERR_OOM, "scht_build_tree: out of host memory");
    } catch (const std::exception& e) {
        return scht::fail(SCHT_ERR_INTERNAL, std::string("scht_build_tree: ") + e.what());
    }
}

extern "C" const scht_tree_desc* scht_host_tree_desc(const scht_host_tree* tree) { return tree ? &tree->desc : nullptr; }
extern "C" double scht_host_tree_build_seconds(const scht_host_tree* tree) { return tree ? tree->build_seconds : 0.0; }
extern "C" void scht_host_tree_free(scht_host_tree* tree) { delete tree; }
