"""Schedule options on one bench workload (GPU box only): one tree, one handle, scht_set_option between runs; every setting must
return the bits of the first.  usage: python tools/exp_opts.py WORKLOAD name=value[,name=value...] [...]     ("-" = defaults)"""
import hashlib, importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
pkg = importlib.import_module("semi-convex_hull_tree_b200")
wl = sys.argv[1]
n, dim, nq, K, pct, kind = bench.WORKLOADS[wl]
data, q = bench.make_data(kind, n, dim, nq, 1234)
ix = pkg.Index(pkg.HostTree(data, pct, 1, device=0))
defaults, ref = {}, None
for spec in sys.argv[2:] or ["-"]:
    for k, v in defaults.items():
        ix.set_option(k, v)
    if spec != "-":
        for kv in spec.split(","):
            k, v = kv.split("=")
            defaults.setdefault(k, ix.get_option(k))
            ix.set_option(k, int(v))
    best = None
    for _ in range(4):
        idx, d2, st = ix.query(q, K, want_stats=True)
        if best is None or st["ms_device"] < best["ms_device"]:
            best = st
    h = hashlib.sha1(idx.tobytes() + d2.tobytes()).hexdigest()[:12]
    ref = ref or h
    print("%-28s %s: device %.2f ms = seed %.2f + select %.2f + main %.2f | evals/q %.0f ins/q %.0f | %s" % (
        spec, wl, best["ms_device"], best["ms_seed_scan"], best["ms_traverse"], best["ms_main_scan"], best["exact_reevaluations"] / nq,
        best["list_insertions"] / nq, "same bits" if h == ref else "DIFFERENT BITS " + h), flush=True)
