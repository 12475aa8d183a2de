"""ctypes binding of the C ABI in include/scht.h (libscht_b200.so).

This is harness plumbing for tests/ and bench.py: the product is the shared library; the C++ host mirror of the
reference's Convex_Tree<T> interface lives in host/B200_Convex_Tree.h.  Nothing here computes a neighbour: every
query goes through the CUDA library, and a missing library or GPU raises (no CPU fallback).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SCHT_LIB") or os.path.join(_HERE, "libscht_b200.so")  # SCHT_LIB: A/B experiments only
SCHT_MAX_K = 128

STATUS = {0: "SCHT_OK", -1: "SCHT_ERR_INVALID_ARG", -2: "SCHT_ERR_DIM_MISMATCH", -3: "SCHT_ERR_NO_DEVICE",
          -4: "SCHT_ERR_CUDA", -5: "SCHT_ERR_OOM", -6: "SCHT_ERR_DEGENERATE", -7: "SCHT_ERR_INTERNAL"}

f32p = C.POINTER(C.c_float)
i32p = C.POINTER(C.c_int32)


class TreeDesc(C.Structure):
    """scht_tree_desc"""
    _fields_ = [("dim", C.c_int32), ("n_points", C.c_int64), ("n_nodes", C.c_int32), ("n_leaves", C.c_int32),
                ("depth", C.c_int32), ("leaf_threshold", C.c_int32),
                ("sorted_points", f32p), ("sorted_to_orig", i32p),
                ("leaf_start", i32p), ("leaf_count", i32p), ("leaf_node_id", i32p),
                ("node_left", i32p), ("node_right", i32p), ("node_parent", i32p), ("node_leaf_index", i32p),
                ("node_left_center", f32p), ("node_right_center", f32p), ("node_normal", f32p),
                ("node_beta_off", i32p), ("node_beta", f32p)]


class Stats(C.Structure):
    """scht_stats"""
    _fields_ = [("n_queries", C.c_int64), ("dist_computations", C.c_int64), ("exact_reevaluations", C.c_int64),
                ("list_insertions", C.c_int64), ("hyperplane_tests", C.c_int64), ("descent_dists", C.c_int64),
                ("batches", C.c_int64), ("ms_total", C.c_double), ("ms_device", C.c_double), ("ms_scan", C.c_double),
                ("scan_launches", C.c_int64), ("kernel_launches", C.c_int64), ("pages_listed", C.c_int64),
                ("prefilter_batches", C.c_int64), ("ms_seed_scan", C.c_double), ("ms_traverse", C.c_double), ("ms_main_scan", C.c_double)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class SchtError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("%s (%d): %s" % (STATUS.get(code, "?"), code, msg))
        self.code = code


# every symbol include/scht.h declares: (restype, argtypes)
SYMBOLS = {
    "scht_build_tree": (C.c_int, [f32p, C.c_int64, C.c_int32, C.c_float, C.c_uint32, C.POINTER(C.c_void_p)]),
    "scht_build_tree_device": (C.c_int, [f32p, C.c_int64, C.c_int32, C.c_float, C.c_uint32, C.c_int32, C.POINTER(C.c_void_p)]),
    "scht_host_tree_desc": (C.POINTER(TreeDesc), [C.c_void_p]),
    "scht_host_tree_build_seconds": (C.c_double, [C.c_void_p]),
    "scht_host_tree_free": (None, [C.c_void_p]),
    "scht_create": (C.c_int, [C.POINTER(TreeDesc), C.c_int, C.POINTER(C.c_void_p)]),
    "scht_create_multi": (C.c_int, [C.POINTER(TreeDesc), i32p, C.c_int32, C.c_int32, C.POINTER(C.c_void_p)]),
    "scht_create_shard": (C.c_int, [C.POINTER(TreeDesc), C.c_int, C.c_int32, C.c_int32, C.POINTER(C.c_void_p)]),
    "scht_destroy": (None, [C.c_void_p]),
    "scht_set_option": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int64]),
    "scht_get_option": (C.c_int, [C.c_void_p, C.c_char_p, C.POINTER(C.c_int64)]),
    "scht_set_batch_size": (C.c_int, [C.c_void_p, C.c_int64]),
    "scht_query": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                             C.POINTER(Stats)]),
    "scht_query_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.POINTER(Stats)]),
    "scht_brute_force": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.POINTER(Stats)]),
    "scht_brute_force_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                          C.c_void_p, C.POINTER(Stats)]),
    "scht_find_approximate_leaves": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p]),
    "scht_query_partial_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                            C.c_void_p, C.c_void_p, C.POINTER(Stats)]),
    "scht_partial_seed_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "scht_partial_scan_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                           C.POINTER(Stats)]),
    "scht_merge_partial_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int64, C.c_int32, C.c_int32,
                                            C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "scht_handle_info": (C.c_int, [C.c_void_p, i32p, C.POINTER(C.c_int64), i32p, i32p, i32p, C.POINTER(C.c_int64)]),
    "scht_comm_unique_id": (C.c_int, [C.c_void_p]),
    "scht_comm_create": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_void_p)]),
    "scht_comm_destroy": (None, [C.c_void_p]),
    "scht_comm_query_slice": (C.c_int, [C.c_void_p, C.c_int64, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "scht_query_sharded_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                            C.c_void_p, C.c_void_p, C.POINTER(Stats)]),
    "scht_handle_layout": (C.c_int, [C.c_void_p, i32p, i32p, i32p, i32p, C.POINTER(C.c_int64), i32p, i32p]),
    "scht_read_text": (C.c_int, [C.c_char_p, C.c_char_p, C.POINTER(f32p), C.POINTER(C.c_int64), C.POINTER(C.c_int32)]),
    "scht_write_text": (C.c_int, [C.c_char_p, C.c_void_p, C.c_int64, C.c_int32]),
    "scht_write_bin": (C.c_int, [C.c_char_p, C.c_void_p, C.c_int64, C.c_int32]),
    "scht_map_bin": (C.c_int, [C.c_char_p, C.POINTER(f32p), C.POINTER(C.c_int64), C.POINTER(C.c_int32), C.POINTER(C.c_void_p)]),
    "scht_unmap_bin": (None, [C.c_void_p]),
    "scht_read_fvecs": (C.c_int, [C.c_char_p, C.POINTER(f32p), C.POINTER(C.c_int64), C.POINTER(C.c_int32)]),
    "scht_free": (None, [C.c_void_p]),
    "scht_format_tree_info": (C.c_int, [C.c_void_p, C.c_float, C.c_char_p, C.c_size_t, C.POINTER(C.c_size_t)]),
    "scht_format_knn_log": (C.c_int, [C.POINTER(Stats), C.c_int64, C.c_int32, C.c_int64, C.c_char_p, C.c_size_t, C.POINTER(C.c_size_t)]),
    "scht_last_error": (C.c_char_p, []),
    "scht_abi_version": (C.c_int, []),
    "scht_device_count": (C.c_int, []),
}

_lib = None


def build_library(verbose=False):
    """Compile csrc/ into libscht_b200.so for sm_100a (nvcc cross-compiles without a GPU)."""
    r = subprocess.run(["make", "-C", os.path.join(_HERE, "csrc")], capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout[-4000:])
        print(r.stderr[-4000:])
    if r.returncode != 0:
        raise RuntimeError("building libscht_b200.so failed")
    return LIB_PATH


def lib():
    """The loaded CUDA library.  Fails loudly when it has not been built: there is no fallback implementation."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("libscht_b200.so is missing: run __graft_entry__.build() (make -C semi-convex_hull_tree_b200/csrc)")
        _lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(_lib, name)
            fn.restype = res
            fn.argtypes = args
    return _lib


def _check(rc):
    if rc != 0:
        raise SchtError(rc, lib().scht_last_error().decode(errors="replace"))


def _np(ptr, shape, dtype):
    n = int(np.prod(shape))
    if n == 0:
        return np.zeros(shape, dtype=dtype)
    return np.ctypeslib.as_array(ptr, shape=(n,)).view(dtype).reshape(shape)


class HostTree:
    """Built tree in host memory: the reference's constructor path, Code/Convex_Tree.h:518-548.  `device=None` runs the
    host builder (scht_build_tree); `device=i` runs the same construction on GPU i (scht_build_tree_device) — the
    arrays are bit-identical either way."""

    def __init__(self, data, leaf_pts_percent, seed=1, device=None):
        data = np.ascontiguousarray(data, dtype=np.float32)
        if data.ndim != 2:
            raise ValueError("data must be [n, dim]")
        self._h = C.c_void_p()
        self.data = data
        if device is None:
            _check(lib().scht_build_tree(data.ctypes.data_as(f32p), data.shape[0], data.shape[1], float(leaf_pts_percent),
                                         int(seed), C.byref(self._h)))
        else:
            _check(lib().scht_build_tree_device(data.ctypes.data_as(f32p), data.shape[0], data.shape[1],
                                                float(leaf_pts_percent), int(seed), int(device), C.byref(self._h)))
        self.desc = lib().scht_host_tree_desc(self._h)
        self.build_seconds = lib().scht_host_tree_build_seconds(self._h)

    def arrays(self):
        """numpy views (borrowed) of the flat arrays."""
        d = self.desc.contents
        n, dim, nn, L = d.n_points, d.dim, d.n_nodes, d.n_leaves
        off = _np(d.node_beta_off, (nn + 1,), np.int32)
        return dict(dim=dim, n_points=n, n_nodes=nn, n_leaves=L, depth=d.depth, leaf_threshold=d.leaf_threshold,
                    sorted_points=_np(d.sorted_points, (n, dim), np.float32), sorted_to_orig=_np(d.sorted_to_orig, (n,), np.int32),
                    leaf_start=_np(d.leaf_start, (L,), np.int32), leaf_count=_np(d.leaf_count, (L,), np.int32),
                    leaf_node_id=_np(d.leaf_node_id, (L,), np.int32), node_left=_np(d.node_left, (nn,), np.int32),
                    node_right=_np(d.node_right, (nn,), np.int32), node_parent=_np(d.node_parent, (nn,), np.int32),
                    node_leaf_index=_np(d.node_leaf_index, (nn,), np.int32),
                    node_left_center=_np(d.node_left_center, (nn, dim), np.float32),
                    node_right_center=_np(d.node_right_center, (nn, dim), np.float32),
                    node_normal=_np(d.node_normal, (nn, dim), np.float32), node_beta_off=off,
                    node_beta=_np(d.node_beta, (max(int(off[-1]), 1),), np.float32)[: int(off[-1])])

    def close(self):
        if self._h:
            lib().scht_host_tree_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def desc_from_arrays(t):
    """Build a TreeDesc from a dict of numpy arrays (e.g. a tree dumped by the reference itself).  Returns
    (desc, keepalive)."""
    keep = {}

    def f(name, dtype):
        a = np.ascontiguousarray(t[name], dtype=dtype)
        if a.size == 0:
            a = np.zeros(1, dtype=dtype)
        keep[name] = a
        return a.ctypes.data_as(f32p if dtype == np.float32 else i32p)

    d = TreeDesc()
    d.dim, d.n_points, d.n_nodes, d.n_leaves = int(t["dim"]), int(t["n_points"]), int(t["n_nodes"]), int(t["n_leaves"])
    d.depth, d.leaf_threshold = int(t.get("depth", 0)), int(t.get("leaf_threshold", 0))
    for name in ("sorted_points", "node_left_center", "node_right_center", "node_normal", "node_beta"):
        setattr(d, name, f(name, np.float32))
    for name in ("sorted_to_orig", "leaf_start", "leaf_count", "leaf_node_id", "node_left", "node_right", "node_parent",
                 "node_leaf_index", "node_beta_off"):
        setattr(d, name, f(name, np.int32))
    return d, keep


SHARD_QUERIES, SHARD_DATASET = 0, 1


class Index:
    """Device-resident tree + query entry points (scht_create / scht_query ...).

    device=i            one GPU (scht_create)
    devices=[...]       several GPUs behind one handle (scht_create_multi; devices=[] = every visible device),
                        mode = SHARD_QUERIES or SHARD_DATASET
    leaf_range=(b, e)   only the leaves [b, e) on `device` (scht_create_shard)"""

    def __init__(self, tree, device=0, devices=None, mode=SHARD_QUERIES, leaf_range=None, options=None):
        self._h = C.c_void_p()
        if isinstance(tree, HostTree):
            desc = tree.desc
            self._keep = tree
        else:
            d, self._keep = desc_from_arrays(tree)
            desc = C.pointer(d)
        self.dim = desc.contents.dim
        self.n_points = desc.contents.n_points
        if devices is not None:
            arr = (C.c_int32 * max(1, len(devices)))(*devices)
            _check(lib().scht_create_multi(desc, arr if len(devices) else None, len(devices), int(mode), C.byref(self._h)))
        elif leaf_range is not None:
            _check(lib().scht_create_shard(desc, int(device), int(leaf_range[0]), int(leaf_range[1]), C.byref(self._h)))
        else:
            _check(lib().scht_create(desc, int(device), C.byref(self._h)))
        self.device = device
        for k, v in (options or {}).items():
            self.set_option(k, v)

    def set_option(self, name, value):
        """scht_set_option: schedule knobs ("tc", "tc_split", "page_select", ...); results never depend on them"""
        _check(lib().scht_set_option(self._h, name.encode(), int(value)))

    def get_option(self, name):
        v = C.c_int64()
        _check(lib().scht_get_option(self._h, name.encode(), C.byref(v)))
        return v.value

    def layout(self):
        nd, mode, ng, ns, lb, le = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
        npg = C.c_int64()
        _check(lib().scht_handle_layout(self._h, C.byref(nd), C.byref(mode), C.byref(ng), C.byref(ns), C.byref(npg), C.byref(lb), C.byref(le)))
        return dict(n_devices=nd.value, mode=mode.value, scan_groups=ng.value, super_groups=ns.value, scan_pages=npg.value,
                    leaf_begin=lb.value, leaf_end=le.value)

    def set_batch_size(self, n):
        _check(lib().scht_set_batch_size(self._h, int(n)))

    def _host_call(self, fn, q, K, want_dist, want_stats):
        q = np.ascontiguousarray(q, dtype=np.float32)
        if q.ndim != 2:
            raise ValueError("queries must be [n, dim]")
        nq = q.shape[0]
        idx = np.empty((nq, K), dtype=np.int32)
        d2 = np.empty((nq, K), dtype=np.float32)
        dist = np.empty((nq, K), dtype=np.float32) if want_dist else None
        st = Stats()
        _check(fn(self._h, q.ctypes.data, nq, q.shape[1], int(K), idx.ctypes.data, d2.ctypes.data,
                  dist.ctypes.data if want_dist else None, C.byref(st) if want_stats else None))
        out = [idx, d2]
        if want_dist:
            out.append(dist)
        if want_stats:
            out.append(st.as_dict())
        return tuple(out)

    def query(self, q, K, want_dist=False, want_stats=False):
        """kNN(Q, K) + get_kNN_indexes / get_kNN_dists_squre (/ get_kNN_dists)  — host buffers in and out."""
        return self._host_call(lib().scht_query, q, K, want_dist, want_stats)

    def brute_force(self, q, K, want_dist=False, want_stats=False):
        return self._host_call(lib().scht_brute_force, q, K, want_dist, want_stats)

    def query_raw(self, q_ptr, nq, dim, K, idx_ptr, d2_ptr, dist_ptr=None, stats=None):
        """scht_query with caller-owned host pointers (e.g. pinned torch tensors)."""
        _check(lib().scht_query(self._h, q_ptr, nq, dim, K, idx_ptr, d2_ptr, dist_ptr, C.byref(stats) if stats is not None else None))

    def query_device(self, q_ptr, nq, dim, K, idx_ptr, d2_ptr, dist_ptr=None, stream=None, stats=None):
        """scht_query_device: all pointers are device pointers on this handle's GPU."""
        _check(lib().scht_query_device(self._h, q_ptr, nq, dim, K, idx_ptr, d2_ptr, dist_ptr, stream,
                                       C.byref(stats) if stats is not None else None))

    def brute_force_device(self, q_ptr, nq, dim, K, idx_ptr, d2_ptr, dist_ptr=None, stream=None, stats=None):
        _check(lib().scht_brute_force_device(self._h, q_ptr, nq, dim, K, idx_ptr, d2_ptr, dist_ptr, stream,
                                             C.byref(stats) if stats is not None else None))

    def partial_seed_device(self, q_ptr, nq, dim, K, bound_ptr, stream=None):
        """phase 1 of the sharded-dataset mode: seed scan on this shard, bound[i] = K-th squared seed distance"""
        _check(lib().scht_partial_seed_device(self._h, q_ptr, nq, dim, K, bound_ptr, stream))

    def partial_scan_device(self, q_ptr, nq, dim, K, bound_ptr, keys_ptr, stream=None, stats=None):
        """phase 2: main scan of this shard with the (all-reduced) bounds -> sorted key lists"""
        _check(lib().scht_partial_scan_device(self._h, q_ptr, nq, dim, K, bound_ptr, keys_ptr, stream,
                                              C.byref(stats) if stats is not None else None))

    def query_partial_device(self, q_ptr, nq, dim, K, leaf_begin, leaf_end, keys_ptr, stream=None, stats=None):
        _check(lib().scht_query_partial_device(self._h, q_ptr, nq, dim, K, leaf_begin, leaf_end, keys_ptr, stream,
                                               C.byref(stats) if stats is not None else None))

    def merge_partial_device(self, keys_ptr, n_lists, nq, dim, K, idx_ptr, d2_ptr, dist_ptr=None, stream=None):
        _check(lib().scht_merge_partial_device(self._h, keys_ptr, n_lists, None, nq, dim, K, idx_ptr, d2_ptr, dist_ptr, stream))

    def find_approximate_leaves(self, q):
        q = np.ascontiguousarray(q, dtype=np.float32)
        out = np.empty(q.shape[0], dtype=np.int32)
        _check(lib().scht_find_approximate_leaves(self._h, q.ctypes.data, q.shape[0], q.shape[1], out.ctypes.data))
        return out

    def info(self):
        dim, nn, L, dev = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
        n, nbytes = C.c_int64(), C.c_int64()
        _check(lib().scht_handle_info(self._h, C.byref(dim), C.byref(n), C.byref(nn), C.byref(L), C.byref(dev), C.byref(nbytes)))
        return dict(dim=dim.value, n_points=n.value, n_nodes=nn.value, n_leaves=L.value, device=dev.value, device_bytes=nbytes.value)

    def close(self):
        if self._h:
            lib().scht_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Comm:
    """scht_comm: NCCL communicator of the library for the sharded-dataset mode with one process per GPU.
    `id_bytes` = Comm.unique_id() of rank 0, distributed by the caller (e.g. torch.distributed.broadcast)."""
    ID_BYTES = 128

    @staticmethod
    def unique_id():
        buf = C.create_string_buffer(Comm.ID_BYTES)
        _check(lib().scht_comm_unique_id(buf))
        return buf.raw

    def __init__(self, id_bytes, rank, world, device):
        self._c = C.c_void_p()
        self.rank, self.world = rank, world
        _check(lib().scht_comm_create(C.c_char_p(bytes(id_bytes)), rank, world, int(device), C.byref(self._c)))

    def query_slice(self, nq):
        b, e = C.c_int64(), C.c_int64()
        _check(lib().scht_comm_query_slice(self._c, nq, C.byref(b), C.byref(e)))
        return b.value, e.value

    def query_sharded_device(self, index, q_ptr, nq, dim, K, idx_ptr, d2_ptr, dist_ptr=None, stream=None, stats=None):
        """scht_query_sharded_device: collective; fills the outputs for this rank's query slice"""
        _check(lib().scht_query_sharded_device(index._h, self._c, q_ptr, nq, dim, K, idx_ptr, d2_ptr, dist_ptr, stream,
                                               C.byref(stats) if stats is not None else None))

    def close(self):
        if self._c:
            lib().scht_comm_destroy(self._c)
            self._c = C.c_void_p()


# ---- data-set files and log text (include/scht.h "Data-set files", "Experiment-log surface") ----------------------
def _take_matrix(ptr, n, dim):
    """copy a malloc'd [n*dim] float matrix into numpy and release it"""
    out = np.ctypeslib.as_array(ptr, shape=(int(n) * int(dim),)).reshape(int(n), int(dim)).copy()
    lib().scht_free(ptr)
    return out


def read_text(path, delims=" "):
    """scht_read_text: the reference's text format (Code/data_processor.h:75-99) -> float32 [n, dim]"""
    ptr, n, d = f32p(), C.c_int64(), C.c_int32()
    _check(lib().scht_read_text(os.fsencode(path), delims.encode(), C.byref(ptr), C.byref(n), C.byref(d)))
    return _take_matrix(ptr, n.value, d.value)


def write_text(path, data):
    data = np.ascontiguousarray(data, dtype=np.float32)
    _check(lib().scht_write_text(os.fsencode(path), data.ctypes.data, data.shape[0], data.shape[1]))


def write_bin(path, data):
    data = np.ascontiguousarray(data, dtype=np.float32)
    _check(lib().scht_write_bin(os.fsencode(path), data.ctypes.data, data.shape[0], data.shape[1]))


def read_bin(path):
    """scht_map_bin + copy (the mapping itself is zero-copy; numpy here owns a copy so the file can be unmapped)"""
    ptr, n, d, m = f32p(), C.c_int64(), C.c_int32(), C.c_void_p()
    _check(lib().scht_map_bin(os.fsencode(path), C.byref(ptr), C.byref(n), C.byref(d), C.byref(m)))
    try:
        if n.value == 0:
            return np.zeros((0, d.value), dtype=np.float32)
        return np.ctypeslib.as_array(ptr, shape=(n.value * d.value,)).reshape(n.value, d.value).copy()
    finally:
        lib().scht_unmap_bin(m)


def read_fvecs(path):
    ptr, n, d = f32p(), C.c_int64(), C.c_int32()
    _check(lib().scht_read_fvecs(os.fsencode(path), C.byref(ptr), C.byref(n), C.byref(d)))
    return _take_matrix(ptr, n.value, d.value)


def _format(call):
    need = C.c_size_t()
    _check(call(None, 0, C.byref(need)))
    buf = C.create_string_buffer(need.value)
    _check(call(buf, need.value, None))
    return buf.value.decode()


def format_tree_info(host_tree, leaf_pts_percent):
    """save_tree_info's text block (Code/Convex_Tree.h:964-991) for a HostTree"""
    return _format(lambda b, c, n: lib().scht_format_tree_info(host_tree._h, C.c_float(leaf_pts_percent), b, c, n))


def format_knn_log(stats, n_queries, K, batch_size):
    """save_KNN_running_time_info's text block (Code/GPU_Convex_Tree.h:973-1005) from a stats dict or Stats"""
    st = stats
    if isinstance(stats, dict):
        st = Stats()
        for k, _ in Stats._fields_:
            setattr(st, k, stats.get(k, 0))
    return _format(lambda b, c, n: lib().scht_format_knn_log(C.byref(st), n_queries, K, batch_size, b, c, n))
