"""one-off: how much of a 100M x 16 uniform image do the ball tests keep for a 1M-query batch (tile unions)?"""
import sys, importlib, os, time, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("semi-convex_hull_tree_b200")
n, dim, nq, K = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000, 16, int(float(sys.argv[2])) if len(sys.argv) > 2 else 1_000_000, 10
data = np.random.default_rng(1234).random((n, dim), dtype=np.float32)
q = np.random.default_rng(2234).random((nq, dim), dtype=np.float32)
t0 = time.time(); ht = pkg.HostTree(data, 500.5 / n, 1, device=0); print("build %.1fs" % (time.time() - t0), flush=True)
t0 = time.time(); ix = pkg.Index(ht); print("create %.1fs" % (time.time() - t0), ix.layout(), flush=True)
for sampling in (1, 0):
    ix.set_option("traverse_sampling", sampling)
    t0 = time.time(); idx, d2, st = ix.query(q, K, want_stats=True); dt = time.time() - t0
    print("sampling %d: %.2fs, pages listed per query %.0f of %d (kept %.3f), scanned/query %.0f, select %.1f ms main %.1f ms seed %.1f ms" % (
        sampling, dt, st["pages_listed"] / nq, ix.layout()["scan_pages"], st["pages_listed"] / nq / ix.layout()["scan_pages"], st["dist_computations"] / nq,
        st["ms_traverse"], st["ms_main_scan"], st["ms_seed_scan"]), flush=True)
    if sampling == 1: ref = idx
    else: print("same result:", np.array_equal(ref, idx))
