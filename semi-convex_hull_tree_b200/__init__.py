"""B200-native exact k-NN query path of the Semi-convex Hull Tree (drop-in for the reference's GPU_Convex_Tree path).

Product: libscht_b200.so (C ABI, include/scht.h) built from csrc/.  This Python package only binds that ABI for
the tests and the benchmark harness; see host/B200_Convex_Tree.h for the C++ mirror of the reference interface.
"""
from ._capi import (LIB_PATH, SCHT_MAX_K, SHARD_DATASET, SHARD_QUERIES, SYMBOLS, Comm, HostTree, Index, SchtError, Stats, TreeDesc, build_library,  # noqa: F401
                    desc_from_arrays, format_knn_log, format_tree_info, lib, read_bin, read_fvecs, read_text, write_bin,
                    write_text)
