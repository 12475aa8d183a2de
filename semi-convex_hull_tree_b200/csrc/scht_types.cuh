// Device-side data layout of the Semi-convex Hull Tree query path: shared by the create path
// (scht_create.cu: upload, scan-order image, page balls) and the query kernels (scht_kernels.cuh, scht_tc.cuh,
// scht_select.cuh).  Replaces the flat OpenCL buffers of GPU_Convex_Tree<T>::init_openCL_sorted_data_set
// (reference: Code/GPU_Convex_Tree.h:522-748) and CONVEX_TREE (Code/ConvexNode.h:243-250).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace scht {

constexpr int kPage = 256;         // sorted points per page
constexpr int kRowsPerStage = 16;  // dims per shared-memory stage (16 rows x 256 floats = 16 KB)
constexpr int kStages = 4;
constexpr int kQPerWarp = 8;       // queries owned by one consumer warp
constexpr int kMaxLevel = 32;      // constraints beyond this depth are ignored by the pruning bound (still valid)
#ifndef SCHT_SEED_POINTS
#define SCHT_SEED_POINTS 256
#endif
constexpr int kSeedPoints = SCHT_SEED_POINTS;  // size of the seed neighbourhood (positions) scanned before the traversal bounds are taken
constexpr int kMaxCuts = 64;       // the traversal of one tile is split into <= 64 subtree jobs (one warp each)
constexpr int kJobRanges = 8;      // page ranges per (tile, subtree job); overflow coalesces gaps (scans more, stays exact)
constexpr int kRangeCap = kMaxCuts * kJobRanges;  // range slots per tile
constexpr uint64_t kEmptyKey = 0x7f7fffffffffffffull;  // FLT_MAX bits (INFINITE, Code/cyw_types.h:18) | all ones
constexpr float kPadCoord = 3.0e38f;                    // coordinate of padded positions: d2 overflows to +inf
constexpr int kMaxShards = 16;     // devices of one sharded-dataset handle / lists of one merge

struct __align__(32) NodeRec {
    int size;      // nodes in the subtree (DFS pre-order: subtree = [id, id+size))
    int depth;     // root = 0; the node has `depth` constraints
    int beta_off;  // into node_beta
    int pos0, pos1;  // sorted-position range covered by the subtree
    int leaf;      // leaf index, -1 for internal nodes
    int right;     // right child id (left child = id+1), -1 for leaves
    int pad;
};

// The tree-order ("sorted position") image.  A handle may hold only a SHARD of the points: positions
// [pos_lo, pos_hi) = a contiguous range of leaves (sharded-dataset mode).  `pages`, `rows` are then allocated for
// the pages [page_lo, page_hi) only and the pointers are offset so that GLOBAL page / position indices address
// them; the tree top (nodes, normals, offsets, centres, leaf tables) and orig_idx are always complete.
struct DevTree {
    int dim, dpad, n_points, n_pages, n_nodes, n_leaves;
    int pos_lo, pos_hi;       // positions resident on this device (whole data set: 0, n_points)
    int page_lo, page_hi;     // pages resident on this device: [pos_lo / 256, ceil(pos_hi / 256))
    const float* pages;       // [n_pages][dpad][256]
    const float* rows;        // [n_pages*256][dpad] the same values row-major when dpad > 16 (else null): an exact re-evaluation reads ONE point
    const int* orig_idx;      // [n_pages*256], -1 on padded positions
    const NodeRec* nodes;     // [n_nodes]
    const float* node_beta;   // CSR by NodeRec::beta_off, root-side constraint first
    const float* node_normal; // [n_nodes][dpad] the node's own (last) constraint normal, zero padded
    const float* node_lc;     // [n_nodes][dpad] left centre
    const float* node_rc;     // [n_nodes][dpad] right centre
    const int* leaf_start;    // [n_leaves]
    const int* leaf_count;    // [n_leaves]
    const int* leaf_rank;     // [n_leaves] bucket of a leaf: leaves ordered by (scan group holding most of its points, leaf id)
    const int* seed_start;    // [n_leaves] position range scanned first for a query whose approximate leaf this is: the
    const int* seed_end;      // [n_leaves]   smallest enclosing subtree with >= kSeedPoints points (its own leaf at least)
    float xnorm_max;          // max ||x||_2 over the resident points (pruning margin)
    int n_cuts, cut_depth;    // subtree roots the traversal is split at (pre-order), all at depth <= cut_depth
    const int* cut_path;      // [n_cuts][cut_depth+1] root-side path of each cut node (node ids, -1 padded), cut node last
};

// The SCAN-ORDER image used by the tensor-core prefilter scan: the resident points re-ordered so that 256 consecutive
// slots (one scan page) are spatially tight.  Built in scht_create.cu: fine k-means groups (~512 points, two Lloyd
// iterations over strided seeds), the groups ordered by a recursive bisection of their centroids (neighbours in the order
// are neighbours in space), and a page break (padding slots) wherever two consecutive groups are separated by empty
// space, so that no page straddles two clusters.  Runs of <= kSuperPages pages between breaks form SUPER-GROUPS.  Every
// super-group and every page carries a bounding ball (centre, radius): a query skips a ball that lies beyond its current
// K-th distance (k_select_pages).  slot_pos maps a slot back to the sorted position the result contract is written in
// (key = d2 | r | pos); padding slots hold -1.
constexpr int kSuperPages = 64;   // (the first-pass page marks of k_select_pages are two 32-bit words per super-group)
struct ScanImage {
    const unsigned char* tcpages;  // per scan page: FP16 image [kp/8][32][8][8] (kp*512 B): s p'_j, then norm hi/mid/lo
    const int* slot_pos;           // [n_spages*256] sorted position of a slot, -1 on padding
    const float* srows;            // [n_spages*256][dpad] the exact coordinates row-major in SCAN order (zeros on padding): a rescorer reads
                                   //   one survivor's row straight from its slot, without first looking up the position
    const int* pcnt;               // [n_spages] real points of the page
    const float* pcen;             // [n_spages][dpad] centre of the page's points
    const float* prad;             // [n_spages] radius (upper bound) around pcen
    const float* scen;             // [n_supers][dpad]
    const float* srad;             // [n_supers]
    const int* scnt;               // [n_supers] real points of the super-group
    const int* spage;              // [n_supers+1] first page of each super-group
    int n_slots, n_spages, n_supers, n_groups;  // n_slots = n_spages*256 (padding included); n_groups: k-means groups the order was built from
    const float* mean;             // [dpad] centre the image is taken around
    const float* pmax;             // [dpad] max |p_j - mean_j|
    float pnorm_max;               // max |p - mean|_2
    float pabs_sum;                // sum_j pmax_j
    float scale;                   // s (power of two)
    float norm_unit;               // power of two u: the norm slots hold s^2|p'|^2 / u, the query operand holds u
    int kp;                        // padded K (multiple of 16) = ceil16(dim + 3); 0: no image
};

__host__ __device__ inline size_t tc_page_bytes(int kp) { return (size_t)kp * 512; }

// relative slack that covers the rounding of an fp32 squared-distance / norm evaluation over `dim` terms (any order):
// |computed - exact| <= (dim + 2) 2^-24 |exact|; doubled and padded
__host__ __device__ inline float ball_eps(int dim) { return (float)(dim + 16) * 1.1920929e-7f; }

}  // namespace scht
