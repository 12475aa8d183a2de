// The host-side tree object behind scht_host_tree (include/scht.h): owner of the flat arrays a scht_tree_desc points
// into.  Filled by the host builder (tree_builder.cpp) or by the device builder (scht_build.cu).
#pragma once
#include <cstdint>
#include <vector>

#include "../../include/scht.h"

struct scht_host_tree {
    scht_tree_desc desc;
    double build_seconds;
    std::vector<float> sorted_points, left_center, right_center, normal, beta;
    std::vector<int32_t> sorted_to_orig, leaf_start, leaf_count, leaf_node_id, left, right, parent, leaf_index, beta_off;
    // point the descriptor at the vectors (call once they have their final sizes)
    void bind(int32_t dim, int64_t n_points, int32_t depth, int32_t leaf_threshold) {
        scht_tree_desc& d = desc;
        d.dim = dim; d.n_points = n_points; d.n_nodes = (int32_t)left.size(); d.n_leaves = (int32_t)leaf_start.size();
        d.depth = depth; d.leaf_threshold = leaf_threshold;
        d.sorted_points = sorted_points.data(); d.sorted_to_orig = sorted_to_orig.data();
        d.leaf_start = leaf_start.data(); d.leaf_count = leaf_count.data(); d.leaf_node_id = leaf_node_id.data();
        d.node_left = left.data(); d.node_right = right.data(); d.node_parent = parent.data();
        d.node_leaf_index = leaf_index.data();
        d.node_left_center = left_center.data(); d.node_right_center = right_center.data();
        d.node_normal = normal.data(); d.node_beta_off = beta_off.data(); d.node_beta = beta.data();
    }
};
