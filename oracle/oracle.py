"""TEST INFRASTRUCTURE — ctypes binding of the CPU oracle (oracle/scht_oracle.c) and a runner for the compiled
reference (oracle/_ref/ref_knn).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module; the product never does."""
import ctypes as C
import importlib
import os
import subprocess
import tempfile

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_LIB = os.path.join(_HERE, "libscht_oracle.so")
REF_BIN = os.path.join(_HERE, "_ref", "ref_knn")
_lib = None


def build():
    r = subprocess.run(["make", "-C", _HERE], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + r.stdout[-2000:] + r.stderr[-2000:])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(ORACLE_LIB):
            build()
        _lib = C.CDLL(ORACLE_LIB)
    return _lib


def _desc(tree):
    """tree: semi-convex_hull_tree_b200.HostTree or dict of arrays -> (ctypes pointer to scht_tree_desc, keepalive)"""
    pkg = importlib.import_module("semi-convex_hull_tree_b200")
    if isinstance(tree, pkg.HostTree):
        return tree.desc, tree
    d, keep = pkg.desc_from_arrays(tree)
    return C.pointer(d), (d, keep)


def knn(tree, queries, K, alg=2, threads=None):
    """CPU_Convex_Tree restatement: returns idx, dist2, dist, approx_node, counters[4]
    (counters: dist computations incl. descent, hyper-plane tests, dot products, scanned points S)."""
    desc, keep = _desc(tree)
    q = np.ascontiguousarray(queries, dtype=np.float32)
    nq = q.shape[0]
    idx = np.empty((nq, K), np.int32)
    d2 = np.empty((nq, K), np.float32)
    d = np.empty((nq, K), np.float32)
    appr = np.empty(nq, np.int32)
    cnt = np.zeros(4, np.int64)
    if threads is not None:
        try:
            C.CDLL("libgomp.so.1").omp_set_num_threads(int(threads))
        except OSError:
            pass
    rc = lib().scht_oracle_knn(desc, q.ctypes.data_as(C.c_void_p), C.c_int64(nq), C.c_int(K), C.c_int(alg), idx.ctypes.data_as(C.c_void_p),
                               d2.ctypes.data_as(C.c_void_p), d.ctypes.data_as(C.c_void_p), appr.ctypes.data_as(C.c_void_p),
                               cnt.ctypes.data_as(C.c_void_p))
    if rc != 0:
        raise RuntimeError("scht_oracle_knn failed")
    return idx, d2, d, appr, cnt


def brute_force(data, queries, K):
    data = np.ascontiguousarray(data, dtype=np.float32)
    q = np.ascontiguousarray(queries, dtype=np.float32)
    nq = q.shape[0]
    idx = np.empty((nq, K), np.int32)
    d2 = np.empty((nq, K), np.float32)
    d = np.empty((nq, K), np.float32)
    rc = lib().scht_oracle_brute_force(data.ctypes.data_as(C.c_void_p), C.c_int64(data.shape[0]), C.c_int(data.shape[1]),
                                       q.ctypes.data_as(C.c_void_p), C.c_int64(nq), C.c_int(K), idx.ctypes.data_as(C.c_void_p),
                                       d2.ctypes.data_as(C.c_void_p), d.ctypes.data_as(C.c_void_p))
    if rc != 0:
        raise RuntimeError("scht_oracle_brute_force failed")
    return idx, d2, d


def contract(tree, queries, K):
    """Order-independent result contract by exhaustive scan (no pruning)."""
    desc, keep = _desc(tree)
    q = np.ascontiguousarray(queries, dtype=np.float32)
    nq = q.shape[0]
    idx = np.empty((nq, K), np.int32)
    d2 = np.empty((nq, K), np.float32)
    rc = lib().scht_oracle_contract(desc, q.ctypes.data_as(C.c_void_p), C.c_int64(nq), C.c_int(K), idx.ctypes.data_as(C.c_void_p),
                                    d2.ctypes.data_as(C.c_void_p))
    if rc != 0:
        raise RuntimeError("scht_oracle_contract failed")
    return idx, d2


def have_reference():
    return os.path.exists(REF_BIN) and os.access(REF_BIN, os.X_OK)


def run_reference(data, queries, K, pct, seed=1, alg=2, want_tree=False, q_begin=0, q_end=None):
    """Run the compiled UNMODIFIED reference (CPU_Convex_Tree<float>) on the inputs.  One kNN() per process
    (SURVEY §9.1).  Returns dict(idx, dist2, dist, approx_leaf_node, counters, seconds[, tree])."""
    from io_formats import read_records, tree_from_records, write_fmat  # oracle/ is on sys.path for the tests
    if q_end is None:
        q_end = len(queries)
    with tempfile.TemporaryDirectory() as td:
        dp, qp, op, tp = [os.path.join(td, n) for n in ("d.fmat", "q.fmat", "o.knn", "o.tree")]
        write_fmat(dp, data)
        write_fmat(qp, queries)
        cmd = [REF_BIN, "knn", dp, repr(float(pct)), str(int(seed)), qp, str(K), str(alg), str(q_begin), str(q_end), op]
        if want_tree:
            cmd.append(tp)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("ref_knn failed: " + r.stderr[-2000:])
        out = read_records(op)
        nq = q_end - q_begin
        for k in ("idx", "dist2", "dist"):
            out[k] = out[k].reshape(nq, K)
        if want_tree:
            out["tree"] = tree_from_records(read_records(tp))
    return out
