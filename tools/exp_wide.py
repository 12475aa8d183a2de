"""one wide-point case per process (the caller wraps it in `timeout`): python tools/exp_wide.py <dim> <tc 0|1> [n nq]"""
import sys, importlib, os, time, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
dim, tc = int(sys.argv[1]), sys.argv[2]
n, nq = (int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else (12000, 300)
os.environ["SCHT_TC"] = tc
pkg = importlib.import_module("semi-convex_hull_tree_b200")
import oracle as orc
rng = np.random.default_rng(dim)
c = np.random.default_rng(4).random((12, dim))
data = (c[rng.integers(0, 12, n)] + 0.08 * rng.standard_normal((n, dim))).astype(np.float32)
q = (c[rng.integers(0, 12, nq)] + 0.08 * rng.standard_normal((nq, dim))).astype(np.float32)
ht = pkg.HostTree(data, 0.02, seed=1)
oi, od2, _, _, _ = orc.knn(ht, q, 10, alg=2)
ix = pkg.Index(ht)
print("dim", dim, "tc", tc, "created", ix.layout(), flush=True)
t0 = time.time()
idx, d2, st = ix.query(q, 10, want_stats=True)
print("dim", dim, "tc", tc, "ok", np.array_equal(idx, oi) and np.array_equal(d2.view(np.uint32), od2.view(np.uint32)), "prefilter", st["prefilter_batches"],
      "%.2fs" % (time.time() - t0), flush=True)
