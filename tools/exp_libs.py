"""A/B of library builds on one bench workload (GPU box only): each library runs in its own process on the same data,
prints the phase times of its best step and a hash of the results (all builds must agree bit for bit).
usage: python tools/exp_libs.py WORKLOAD lib1.so [lib2.so ...]      (libraries from tools/build_variant.sh, or the product .so)"""
import hashlib, importlib, os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "--child":
    import numpy as np
    sys.path.insert(0, ROOT)
    import bench
    pkg = importlib.import_module("semi-convex_hull_tree_b200")
    pkg._capi.LIB_PATH = os.environ["SCHT_AB_LIB"]
    wl = sys.argv[2]
    n, dim, nq, K, pct, kind = bench.WORKLOADS[wl]
    data, q = bench.make_data(kind, n, dim, nq, 1234)
    ix = pkg.Index(pkg.HostTree(data, pct, 1, device=0))
    best = None
    for _ in range(4):
        idx, d2, st = ix.query(q, K, want_stats=True)
        if best is None or st["ms_device"] < best["ms_device"]:
            best = st
    h = hashlib.sha1(idx.tobytes() + d2.tobytes()).hexdigest()[:12]
    print("%-22s %s: device %.2f ms = seed %.2f + select %.2f + main %.2f | evals/q %.0f ins/q %.0f pairs/q %.0f | %s" % (
        os.path.basename(os.environ["SCHT_AB_LIB"]), wl, best["ms_device"], best["ms_seed_scan"], best["ms_traverse"], best["ms_main_scan"],
        best["exact_reevaluations"] / nq, best["list_insertions"] / nq, best["dist_computations"] / nq, h), flush=True)
    sys.exit(0)
wl = sys.argv[1]
for lib in sys.argv[2:]:
    subprocess.run([sys.executable, os.path.abspath(__file__), "--child", wl], env=dict(os.environ, SCHT_AB_LIB=os.path.abspath(lib)), check=False)
