"""Pruning effectiveness: points scanned per query by the reference algorithm (oracle counter) vs by the GPU path."""
import sys, importlib, os, numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/oracle')
pkg = importlib.import_module("semi-convex_hull_tree_b200")
import oracle as orc
def clusters(rng, n, dim, k, sigma, scale=1.0):
    c = np.random.default_rng(7).random((k, dim))
    return ((c[rng.integers(0, k, n)] + sigma * rng.standard_normal((n, dim))) * scale).astype(np.float32)
rng = np.random.default_rng(1)
NQ = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
for (n, dim, k, sigma, pct) in [(200000, 32, 100, 0.05, 0.0025), (200000, 15, 64, 0.05, 0.0025), (200000, 16, 0, 0, 0.0025)]:
    if k: data, q = clusters(rng, n, dim, k, sigma), clusters(rng, NQ, dim, k, sigma)
    else: data, q = rng.random((n, dim), dtype=np.float32), rng.random((NQ, dim), dtype=np.float32)
    ht = pkg.HostTree(data, pct, seed=1)
    a = ht.arrays()
    oi, od2, od, _, cnt = orc.knn(ht, q[:256], 10, alg=2)
    os.environ["SCHT_TC"] = "0"
    ix = pkg.Index(ht)
    idx, d2, st = ix.query(q, 10, want_stats=True)
    ok = np.array_equal(idx[:256], oi)
    print("n=%d dim=%d clusters=%d leaves=%d depth=%d | ref scanned/query %.0f (%.1f%%) | gpu scanned/query %.0f (%.1f%%), listed pages/query %.1f of %d, tests/query %.0f | parity %s" % (
        n, dim, k, a["n_leaves"], a["depth"], cnt[3] / 256, 100 * cnt[3] / 256 / n, st["dist_computations"] / NQ, 100 * st["dist_computations"] / NQ / n,
        st["pages_listed"] / NQ, (n + 255) // 256, st["hyperplane_tests"] / NQ, ok), flush=True)
    ix.close()
