"""One cfg2 step (uniform 1Mx16, 100k queries, K=10) with kernel timings; optional self-check against brute force on a sample."""
import importlib, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("semi-convex_hull_tree_b200")
rng = np.random.default_rng(1234)
data = rng.random((1_000_000, 16), dtype=np.float32)
q = np.random.default_rng(2234).random((100_000, 16), dtype=np.float32)
ht = pkg.HostTree(data, 0.0005, 1, device=0)
ix = pkg.Index(ht)
best = None
for _ in range(4):
    idx, d2, st = ix.query(q, 10, want_stats=True)
    if best is None or st["ms_device"] < best["ms_device"]:
        best = st
print({k: round(best[k], 3) for k in ("ms_device", "ms_seed_scan", "ms_traverse", "ms_main_scan")}, "exact_reeval", best["exact_reevaluations"])
if "--check" in sys.argv:
    bi, bd2 = ix.brute_force(q[:2000], 10)
    print("check vs exhaustive scan:", np.array_equal(bi, idx[:2000]), np.array_equal(bd2.view(np.uint32), d2[:2000].view(np.uint32)))
