"""Small end-to-end run for compute-sanitizer: exact and prefilter paths, brute force, partial mode."""
import sys, importlib, os, numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/oracle')
pkg = importlib.import_module("semi-convex_hull_tree_b200")
import oracle as orc
rng = np.random.default_rng(0)
K = 10
for n, dim, nq in ((20000, 16, 600), (12000, 72, 400)):  # narrow points; wide points (row-major copy, 2-stage ring, N=256 chains)
    data = rng.random((n, dim), dtype=np.float32); q = rng.random((nq, dim), dtype=np.float32)
    ht = pkg.HostTree(data, 0.01, seed=1)
    oi, od2, _, _, _ = orc.knn(ht, q, K, alg=2)
    for tc in ("0", "1"):
        os.environ["SCHT_TC"] = tc
        ix = pkg.Index(ht)
        idx, d2 = ix.query(q, K)
        bi, bd2 = ix.brute_force(q[:100], K)
        print("dim", dim, "tc", tc, "ok", np.array_equal(idx, oi) and np.array_equal(d2.view(np.uint32), od2.view(np.uint32)), np.array_equal(bi, oi[:100]), flush=True)
        ix.close()
