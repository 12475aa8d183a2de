"""Small batches on the cfg2 tree (uniform 1Mx16, K=10): device time per batch for several batch sizes (tail / occupancy effects)."""
import importlib, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("semi-convex_hull_tree_b200")
data = np.random.default_rng(1234).random((1_000_000, 16), dtype=np.float32)
ht = pkg.HostTree(data, 0.0005, 1, device=0)
ix = pkg.Index(ht)
for nq in (512, 4096, 16384, 50000):
    q = np.random.default_rng(7).random((nq, 16), dtype=np.float32)
    best = None
    for _ in range(3):
        idx, d2, st = ix.query(q, 10, want_stats=True)
        if best is None or st["ms_device"] < best["ms_device"]:
            best = st
    bi, bd2 = ix.brute_force(q[:256], 10)
    ok = np.array_equal(bi, idx[:256]) and np.array_equal(bd2.view(np.uint32), d2[:256].view(np.uint32))
    print("nq", nq, {k: round(best[k], 3) for k in ("ms_device", "ms_seed_scan", "ms_main_scan")}, "exact", ok)
