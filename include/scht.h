/*
 * scht.h — C ABI of the B200-native Semi-convex Hull Tree k-NN query path.
 *
 * This is the drop-in boundary: exactly what a binding of the reference's GPU back-end
 * (Code/GPU_Convex_Tree.h) would bind instead of OpenCL.  extern "C", plain pointers and sizes,
 * no C++/torch types.  Every entry point cites the reference interface it replaces
 * (paths relative to the reference repo, XiaoLHu/Semi-convex_Hull_Tree).
 *
 * Conventions
 *   - every function returns SCHT_OK (0) or a negative scht_status; no exception crosses the ABI;
 *     scht_last_error() returns a thread-local, human-readable message for the last failure.
 *   - there is NO CPU fallback: every query entry point fails with SCHT_ERR_NO_DEVICE / SCHT_ERR_CUDA
 *     when no sm_100 device (or no CUDA driver) is present.
 *   - output buffers are caller-allocated; a handle is not thread-safe (like the reference).
 *   - FLOAT_TYPE is float (Code/cyw_types.h:5-16, USE_DOUBLE 0); indices are int32 (the reference uses int).
 */
#ifndef SCHT_H_INCLUDED
#define SCHT_H_INCLUDED

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SCHT_ABI_VERSION 2
/* K limit of the device path. The reference's OpenCL path hard-codes MAX_NN_NUM=100
 * (Code/GPU_Convex_Tree.h:466, kernels/do_kNN_gpu.cl:360-362) without checking; we check. */
#define SCHT_MAX_K 128

typedef enum scht_status {
    SCHT_OK = 0,
    SCHT_ERR_INVALID_ARG = -1,  /* NULL pointer, K<1, K>SCHT_MAX_K, negative sizes, bad percent ... */
    SCHT_ERR_DIM_MISMATCH = -2, /* query dim != tree dim (Code/Convex_Tree.h:458-461 prints and returns) */
    SCHT_ERR_NO_DEVICE = -3,    /* no CUDA device / not sm_100: the product has no CPU fallback */
    SCHT_ERR_CUDA = -4,         /* a CUDA runtime call or kernel failed; see scht_last_error() */
    SCHT_ERR_OOM = -5,          /* host or device allocation failed */
    SCHT_ERR_DEGENERATE = -6,   /* reserved (the builders close an unsplittable node -- more than leaf-threshold identical rows -- as an
                                   oversized leaf and return SCHT_OK, where the reference would recurse forever, Code/Convex_Tree.h:809-856) */
    SCHT_ERR_INTERNAL = -7
} scht_status;

/* ------------------------------------------------------------------------------------------------
 * Flat host-side description of a built tree — the SoA restatement of what
 * GPU_Convex_Tree<T>::init_openCL_sorted_data_set flattens out of Convex_Tree<T>::nodes
 * (Code/GPU_Convex_Tree.h:522-748) and of CONVEX_TREE (Code/ConvexNode.h:243-250).
 *
 * Node ids are the reference's creation order = DFS pre-order (Code/Convex_Tree.h:725,739);
 * leaf ids are node-index order of leaves (Code/Convex_Tree.h:685-705); sorted_points holds the
 * dataset re-ordered so that each leaf is contiguous (sort_raw_data, Code/Convex_Tree.h:593-633).
 *
 * Constraints: node i at depth k (root = depth 0) has k half-spaces  a_j . x + b_j >= 0.
 * a_j is the unit normal of the j-th edge on the root->i path, stored ONCE per node as
 * node_normal[path_node] (the last ALPHA row of that node, Code/Convex_Tree.h:767; the right
 * child's is the exact negation of the left's, :848).  b_j is re-tightened per node
 * (Code/Convex_Tree.h:760-766,771): node_beta[node_beta_off[i] + j], j = 0..k-1, root-side first.
 * ---------------------------------------------------------------------------------------------- */
typedef struct scht_tree_desc {
    int32_t dim;      /* D */
    int64_t n_points; /* N (< 2^31) */
    int32_t n_nodes;
    int32_t n_leaves;
    int32_t depth;          /* reference's `depth` statistic (Code/Convex_Tree.h:285,837-839); informational */
    int32_t leaf_threshold; /* (int)(N*pct), Code/Convex_Tree.h:525; informational */

    const float* sorted_points;    /* [N*D]   m_sorted_data (Code/Convex_Tree.h:322) */
    const int32_t* sorted_to_orig; /* [N]     m_sorted_data_ori_indexes (:323) */

    const int32_t* leaf_start;   /* [L] leaf_nodes_start_pos_in_data_set (:325) */
    const int32_t* leaf_count;   /* [L] ConvexNode::pts_number (Code/ConvexNode.h:45) */
    const int32_t* leaf_node_id; /* [L] leaf_nodes_ori_indexes (Code/Convex_Tree.h:348) */

    const int32_t* node_left;       /* [n_nodes] ConvexNode::left  (-1 for leaves) */
    const int32_t* node_right;      /* [n_nodes] ConvexNode::right (-1 for leaves) */
    const int32_t* node_parent;     /* [n_nodes] ConvexNode::parent_index (-1 for root); informational: scht_create derives the parents
                                       from the validated child links and never reads this array (may be NULL) */
    const int32_t* node_leaf_index; /* [n_nodes] leaf id, -1 for internal nodes */

    const float* node_left_center;  /* [n_nodes*D] ConvexNode::left_center  (zeros for leaves) */
    const float* node_right_center; /* [n_nodes*D] ConvexNode::right_center (zeros for leaves) */
    const float* node_normal;       /* [n_nodes*D] last ALPHA row of the node (zeros for root) */
    const int32_t* node_beta_off;   /* [n_nodes+1] CSR offsets into node_beta; off[i+1]-off[i] = depth(i) */
    const float* node_beta;         /* [sum depth] ConvexNode::BETA */
} scht_tree_desc;

/* ------------------------------------------------------------------------------------------------
 * Host tree builder.  Restates Convex_Tree<T>'s constructor path
 * (Code/Convex_Tree.h:518-548 -> build_tree :708-857 -> build_simplified_tree :685-705 ->
 * sort_raw_data :593-633) with the reference's float/double arithmetic and its libc rand() stream
 * (glibc TYPE_3 generator re-implemented, seeded like srand(seed); seed 1 == "srand never called"),
 * so the flat arrays are bit-identical to the reference's tree for the same input and seed.
 * Pure host code; needs no GPU.
 *
 * leaf_pts_percent must be in (0, 0.5) — the reference throws std::runtime_error otherwise
 * (Code/Convex_Tree.h:522-523); here: SCHT_ERR_INVALID_ARG.
 * ---------------------------------------------------------------------------------------------- */
typedef struct scht_host_tree scht_host_tree;

int scht_build_tree(const float* data, int64_t n_points, int32_t dim, float leaf_pts_percent,
                    uint32_t rand_seed, scht_host_tree** out);
/* Borrowed view; valid until scht_host_tree_free. */
const scht_tree_desc* scht_host_tree_desc(const scht_host_tree* tree);
/* Seconds spent in build_tree+build_simplified_tree (timer_build_tree, Code/Convex_Tree.h:532-539). */
double scht_host_tree_build_seconds(const scht_host_tree* tree);
void scht_host_tree_free(scht_host_tree* tree);

/* Same construction executed on one B200: a persistent cooperative kernel walks the
 * reference's DFS (build_tree, Code/Convex_Tree.h:708-857) node by node — the rand() stream forces that
 * order — and parallelises the work inside each node (pivot distances, getBeta_1 re-tightening
 * :563-589, stable partition :809-821).  The arithmetic is the reference's, so the resulting
 * scht_host_tree is BIT-IDENTICAL to scht_build_tree's for the same data and seed.  Needs an sm_100
 * device (SCHT_ERR_NO_DEVICE otherwise; there is no silent fallback to the host builder) and
 * 2*N*D*4 + 13*N bytes of device memory; trees deeper than 64 levels -> SCHT_ERR_INVALID_ARG. */
int scht_build_tree_device(const float* data, int64_t n_points, int32_t dim, float leaf_pts_percent,
                           uint32_t rand_seed, int32_t device, scht_host_tree** out);

/* ------------------------------------------------------------------------------------------------
 * Device tree + query.  Replaces GPU_Convex_Tree<T>'s constructor tail
 * (init_openCL_device / init_openCL_sorted_data_set, Code/GPU_Convex_Tree.h:150-163,431-479,522-748)
 * and its do_kNN / do_brute_force_kNN (Code/GPU_Convex_Tree.h:208-323,328-427) together with the
 * two OpenCL kernels (Code/kernels/do_kNN_gpu.cl:298-446,462-536).
 * ---------------------------------------------------------------------------------------------- */
typedef struct scht_handle scht_handle;

/* Counters of one query call; mirror the reference's statistics surface
 * (Code/GPU_Convex_Tree.h:57-65,950-971; Code/Convex_Tree.h:288-296). */
typedef struct scht_stats {
    int64_t n_queries;
    int64_t dist_computations;   /* (query,point) distances the leaf scan evaluated, incl. the approximate leaves (S_gpu) */
    int64_t exact_reevaluations; /* pairs re-evaluated with the reference's float/double arithmetic (near the K-th distance) */
    int64_t list_insertions;     /* successful top-K insertions (NNResult::update hits, Code/Convex_Tree.h:121-151) */
    int64_t hyperplane_tests;    /* (query,node) lower-bound tests ("quadprog calls") */
    int64_t descent_dists;       /* 2 per level of the approximate-leaf descent (dist_computation_times_in_host) */
    int64_t batches;             /* "batch iteration times" (Code/GPU_Convex_Tree.h:968) */
    double ms_total;             /* wall time of the call (timer_whole_alg) */
    double ms_device;            /* CUDA-event time of the device work of the call */
    double ms_scan;              /* CUDA-event time of the leaf-scan kernel(s) only */
    int64_t scan_launches;       /* number of leaf-scan kernel launches */
    int64_t kernel_launches;     /* number of kernels launched by the call */
    int64_t pages_listed;        /* (query,page) pairs the traversal kept for the main scan, x256 ~ candidate points */
    int64_t prefilter_batches;   /* batches whose main scan used the tensor-core prefilter */
    double ms_seed_scan;         /* CUDA-event time of the seed scan (approximate leaves) */
    double ms_traverse;          /* CUDA-event time of the eligibility + page-selection / traversal kernels */
    double ms_main_scan;         /* CUDA-event time of the main scan (k_scan or k_scan_tc) */
} scht_stats;

/* Uploads the tree to `device` (CUDA ordinal) in the device layout described in DESIGN.md.
 * device == -1: every visible device, queries sharded (== scht_create_multi(tree, NULL, 0, SCHT_SHARD_QUERIES, out)). */
int scht_create(const scht_tree_desc* tree, int device, scht_handle** out);
void scht_destroy(scht_handle* h);

/* ---- several GPUs behind one handle (the reference drives a single device, Code/GPU_Convex_Tree.h:443-446) ----
 * One process, one host thread per GPU inside scht_query / scht_brute_force; the caller sees the same entry points and
 * the same bits as with one GPU.
 *   SCHT_SHARD_QUERIES  tree replicated on every device, each answers a contiguous slice of the queries of a call; the
 *                       devices never communicate.
 *   SCHT_SHARD_DATASET  each device holds a contiguous range of leaves (1/n of the points; the tree top is replicated) and
 *                       scans it for all queries; seed bounds are read from the peers and every device's scan kernel stores
 *                       its per-query K-lists directly into the memory of the device that owns the query (peer access over
 *                       NVLink), which merges them.  Needs peer access between all listed devices.
 * devices == NULL or n_devices <= 0: every visible device.  At most 16 devices. */
#define SCHT_SHARD_QUERIES 0
#define SCHT_SHARD_DATASET 1
int scht_create_multi(const scht_tree_desc* tree, const int32_t* devices, int32_t n_devices, int32_t mode, scht_handle** out);

/* A handle that holds only the leaves [leaf_begin, leaf_end) of the tree on `device` (the building block of the
 * sharded-dataset mode when every GPU belongs to a different process, see scht_partial_seed_device).  scht_query on
 * such a handle answers with the nearest neighbours among the resident points only. */
int scht_create_shard(const scht_tree_desc* tree, int device, int32_t leaf_begin, int32_t leaf_end, scht_handle** out);

/* Schedule knobs of a handle (none changes a result bit).  The SCHT_* environment variables only provide
 * the defaults read when the handle is created.  Names: "tc" (0 off / 1 always / 2 auto), "tc_split", "tc_issuers",
 * "tc_n256", "scan_variant", "scan_stages", "traverse_sampling", "sample_reuse", "seed_points", "page_select",
 * "brute_tc", "pipeline", "pipeline_split", "first_pass", "rescorer_sleep_ns", "tc_ring" (0 or a power of two in
 * 32..1024), "tc_recheck", "ring_wait_ns", "lazy_seed"; "groups" can only be read (fixed at creation).  Unknown name or value out
 * of range -> SCHT_ERR_INVALID_ARG. */
int scht_set_option(scht_handle* h, const char* name, int64_t value);
int scht_get_option(const scht_handle* h, const char* name, int64_t* value);

/* BATCH_SIZE_OF_QUERY_POINTS_PUSCH_TO_DEVICE (Code/Convex_Tree.h:352, set_batch_size
 * Code/GPU_Convex_Tree.h:197-204): queries are streamed through the device in batches of this
 * many.  Default 1,000,000 (Code/main.cc:670).  batch_size < 1 -> SCHT_ERR_INVALID_ARG. */
int scht_set_batch_size(scht_handle* h, int64_t batch_size);

/* Exact k-NN of n_queries host-resident queries ([n_queries*dim] row-major).  Host->device copy of
 * the queries and device->host copy of the results are part of the call (== kNN() + get_kNN_*,
 * Code/Convex_Tree.h:456-469,231-238).
 *   idx_out   [n_queries*K] original point indices (get_kNN_indexes), -1 where fewer than K points exist
 *   dist2_out [n_queries*K] ascending squared distances (get_kNN_dists_squre), FLT_MAX for empty slots
 *   dist_out  [n_queries*K] optional (may be NULL): sqrtf(dist2) (get_kNN_dists)
 * Result contract: bit-identical to CPU_Convex_Tree (Code/CPU_Convex_Tree.h:380-430): the K smallest
 * (dist2, approx-leaf-first, sorted position) triples; see DESIGN.md "Result contract". */
int scht_query(scht_handle* h, const float* queries, int64_t n_queries, int32_t dim, int32_t K,
               int32_t* idx_out, float* dist2_out, float* dist_out, scht_stats* stats /* may be NULL */);

/* Same computation with queries and outputs already resident in device memory of the handle's GPU
 * (device pointers; row-major as above).  Runs on `cuda_stream` (a cudaStream_t cast to void*; NULL =
 * the handle's own stream) and returns after enqueueing + synchronising that stream. */
int scht_query_device(scht_handle* h, const float* d_queries, int64_t n_queries, int32_t dim, int32_t K,
                      int32_t* d_idx_out, float* d_dist2_out, float* d_dist_out, void* cuda_stream,
                      scht_stats* stats);

/* Exhaustive search over all points, same result contract with "approximate leaf" = none, i.e. ties
 * go to the lower ORIGINAL index (brute_force_kNN / do_brute_force_kNN,
 * Code/Convex_Tree.h:473-486, Code/GPU_Convex_Tree.h:328-427, kernels/do_kNN_gpu.cl:462-536). */
int scht_brute_force(scht_handle* h, const float* queries, int64_t n_queries, int32_t dim, int32_t K,
                     int32_t* idx_out, float* dist2_out, float* dist_out, scht_stats* stats);

/* scht_brute_force with device-resident queries and outputs (single-GPU handles), like scht_query_device. */
int scht_brute_force_device(scht_handle* h, const float* d_queries, int64_t n_queries, int32_t dim, int32_t K,
                            int32_t* d_idx_out, float* d_dist2_out, float* d_dist_out, void* cuda_stream,
                            scht_stats* stats);

/* Approximate-leaf descent only (do_find_approximate_nodes, Code/GPU_Convex_Tree.h:168-193):
 * leaf_out[i] = leaf index reached by query i. Host buffers. */
int scht_find_approximate_leaves(scht_handle* h, const float* queries, int64_t n_queries, int32_t dim,
                                 int32_t* leaf_out);

/* ---- sharded-dataset mode, one process per GPU ---------------------------------------------------------
 * Each rank creates a shard handle (scht_create_shard) for its leaf range and, per batch of queries (all ranks pass the
 * same queries; n_queries * max(K, dim) < 2^31):
 *   1. scht_partial_seed_device   seed scan where the query's approximate leaf is resident; d_bound_out[i] = K-th squared
 *                                 distance of query i's seed list (FLT_MAX when the list is not full / nothing resident)
 *   2. the caller takes the element-wise MINIMUM of d_bound over the ranks (one all-reduce of 4 B per query)
 *   3. scht_partial_scan_device   main scan of the shard with that bound; per query K sorted 64-bit keys
 *                                     key = (float_bits(dist2) << 32) | (is_not_approx_leaf << 31) | sorted_position
 *                                 (FLT_MAX-keyed for empty slots) into d_keys_out [n_queries*K]
 *   4. the ranks exchange the key lists (NCCL all-to-all over NVLink; scht_comm_* below does 2.-5. inside the library)
 *   5. scht_merge_partial_device  merges n_lists lists laid out [n_lists][n_queries][K], maps positions back to original
 *                                 indices and writes the final idx / dist2 / dist arrays.
 * scht_query_partial_device is the one-call form for a handle that holds MORE than the requested leaf range (e.g. the whole
 * tree): the seed scan uses everything resident, its K-th distance bounds the scan, and only positions of
 * [leaf_begin, leaf_end) enter the key lists (leaf_begin = leaf_end = -1: the resident range). */
int scht_query_partial_device(scht_handle* h, const float* d_queries, int64_t n_queries, int32_t dim, int32_t K,
                              int32_t leaf_begin, int32_t leaf_end, uint64_t* d_keys_out, void* cuda_stream,
                              scht_stats* stats);
int scht_partial_seed_device(scht_handle* h, const float* d_queries, int64_t n_queries, int32_t dim, int32_t K,
                             float* d_bound_out, void* cuda_stream);
int scht_partial_scan_device(scht_handle* h, const float* d_queries, int64_t n_queries, int32_t dim, int32_t K,
                             const float* d_bound, uint64_t* d_keys_out, void* cuda_stream, scht_stats* stats);
int scht_merge_partial_device(scht_handle* h, const uint64_t* d_keys, int32_t n_lists, const float* d_queries,
                              int64_t n_queries, int32_t dim, int32_t K, int32_t* d_idx_out, float* d_dist2_out,
                              float* d_dist_out, void* cuda_stream);

/* The same five steps inside the library, one process per GPU, NCCL over NVLink for steps 2 and 4 (libnccl.so.2 is loaded
 * at run time, only by these entry points):
 *   rank 0: scht_comm_unique_id(id) and hands the SCHT_COMM_ID_BYTES bytes to every rank (MPI, torch.distributed, a file ...)
 *   every rank: scht_comm_create(id, rank, world, device, &comm); scht_create_shard(tree, device, its leaf range, &h)
 *   per batch:  scht_query_sharded_device(h, comm, d_queries (ALL queries, same on every rank), ...) fills d_idx_out /
 *               d_dist2_out / d_dist_out for the rank's contiguous query slice [begin, end) (scht_comm_query_slice;
 *               the arrays hold (end - begin) * K elements).  Collective: every rank must call it with the same queries.
 * Results are bit-identical to scht_query on a single GPU holding the whole tree. */
#define SCHT_COMM_ID_BYTES 128
typedef struct scht_comm scht_comm;
int scht_comm_unique_id(void* id_out /* SCHT_COMM_ID_BYTES */);
int scht_comm_create(const void* id, int32_t rank, int32_t world, int32_t device, scht_comm** out);
void scht_comm_destroy(scht_comm* c);
int scht_comm_query_slice(const scht_comm* c, int64_t n_queries, int64_t* begin, int64_t* end);
int scht_query_sharded_device(scht_handle* h, scht_comm* c, const float* d_queries, int64_t n_queries, int32_t dim,
                              int32_t K, int32_t* d_idx_out, float* d_dist2_out, float* d_dist_out, void* cuda_stream,
                              scht_stats* stats);

/* Tree facts of a handle (for logs: print_tree_info, Code/Convex_Tree.h:946-958). */
int scht_handle_info(const scht_handle* h, int32_t* dim, int64_t* n_points, int32_t* n_nodes, int32_t* n_leaves,
                     int32_t* device, int64_t* device_bytes);
/* Device layout facts: devices behind the handle, sharding mode, k-means groups / super-groups / scan pages (padding
 * included) of the prefilter image of the first device (0 when it was not built) and the leaves resident there.  Any
 * pointer may be NULL. */
int scht_handle_layout(const scht_handle* h, int32_t* n_devices, int32_t* mode, int32_t* n_scan_groups,
                       int32_t* n_super_groups, int64_t* n_scan_pages, int32_t* leaf_begin, int32_t* leaf_end);

/* ------------------------------------------------------------------------------------------------
 * Data-set files.  Replaces read_data / read_data_dim_size / get_dim / get_data
 * (Code/data_processor.h:5-99: freopen + gets + strtok + atof, one point per line).
 *   scht_read_text   same file contract: dim = number of `delims`-separated tokens of the first line, one row
 *                    per line (empty lines included), value = (float)atof(token); mmap'd and parsed on all host
 *                    cores.  Defined where the reference is not: short lines are zero-filled, extra tokens
 *                    ignored, a trailing '\r' is a delimiter.  delims NULL/"" = " " (Code/main.cc:146).
 *   scht_write_text  writes that format ("%.9g", space separated): floats round-trip exactly.
 *   scht_write_bin / scht_map_bin / scht_unmap_bin
 *                    flat binary: 32-byte header {"SCHTBIN1", int32 dim, int32 0, int64 n_points, int64 0}
 *                    followed by n_points*dim little-endian floats, row-major; scht_map_bin returns a
 *                    read-only mmap view (rows 32-byte aligned), valid until scht_unmap_bin(mapping).
 *   scht_read_fvecs  the .fvecs format of the SIFT/GIST sets: per row int32 dim, then dim floats.
 * Matrices returned through float** are malloc'd; release with scht_free.  N*D must stay below 2^32 like in
 * the reference (Code/Array.hh:1175-1183). */
int scht_read_text(const char* path, const char* delims, float** data_out, int64_t* n_points, int32_t* dim);
int scht_write_text(const char* path, const float* data, int64_t n_points, int32_t dim);
int scht_write_bin(const char* path, const float* data, int64_t n_points, int32_t dim);
int scht_map_bin(const char* path, const float** data_out, int64_t* n_points, int32_t* dim, void** mapping_out);
void scht_unmap_bin(void* mapping);
int scht_read_fvecs(const char* path, float** data_out, int64_t* n_points, int32_t* dim);
void scht_free(void* p);

/* ------------------------------------------------------------------------------------------------
 * Experiment-log surface: the text blocks the reference's drivers append to their log
 * files, line for line -- Convex_Tree<T>::save_tree_info (Code/Convex_Tree.h:964-991) and
 * GPU_Convex_Tree<T>::save_KNN_running_time_info (Code/GPU_Convex_Tree.h:973-1005; print_ variant :950-971),
 * numbers formatted like my_strcat / my_strcat_longlong / my_strcat_double (Code/basic_functions.h:313-334).
 * Counter mapping: (2) quadprog calls = scht_stats.hyperplane_tests, (3) = dist_computations,
 * (4) = descent_dists, (6) = batch size / scht_stats.batches.
 * Writes a NUL-terminated string into buf[cap]; *needed (optional) = bytes required incl. the NUL;
 * buf == NULL with needed != NULL only measures. */
int scht_format_tree_info(const scht_host_tree* tree, float leaf_pts_percent, char* buf, size_t cap, size_t* needed);
int scht_format_knn_log(const scht_stats* stats, int64_t n_queries, int32_t K, int64_t batch_size, char* buf, size_t cap,
                        size_t* needed);

const char* scht_last_error(void);
int scht_abi_version(void);
/* Number of visible CUDA devices (0 if none / no driver); never fails. */
int scht_device_count(void);

#ifdef __cplusplus
}
#endif
#endif /* SCHT_H_INCLUDED */
