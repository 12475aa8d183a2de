"""Times scht_build_tree_device (one B200) against scht_build_tree (one host core) and checks bit-identity.
usage: python tools/build_bench.py N DIM PCT [kind=uniform|gauss] [--no-host]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import importlib

pkg = importlib.import_module("semi-convex_hull_tree_b200")

n, dim, pct = int(sys.argv[1]), int(sys.argv[2]), float(sys.argv[3])
kind = sys.argv[4] if len(sys.argv) > 4 and not sys.argv[4].startswith("--") else "uniform"
host = "--no-host" not in sys.argv
rng = np.random.default_rng(1)
if kind == "uniform":
    data = rng.random((n, dim), dtype=np.float32)
else:
    c = rng.random((1000, dim), dtype=np.float32)
    data = c[rng.integers(0, 1000, n)] + np.float32(0.05) * rng.standard_normal((n, dim), dtype=np.float32)
out = {"n": n, "dim": dim, "pct": pct, "kind": kind}
pkg.HostTree(data[:20000], 0.01, 1, device=0).close()  # context + module load
t0 = time.time()
dev = pkg.HostTree(data, pct, 1, device=0)
out["device_build_s"] = round(time.time() - t0, 3)
out["device_build_reported_s"] = round(dev.build_seconds, 3)
a = dev.arrays()
out.update(n_nodes=int(a["n_nodes"]), n_leaves=int(a["n_leaves"]), depth=int(a["depth"]))
if host:
    t0 = time.time()
    ht = pkg.HostTree(data, pct, 1)
    out["host_build_s"] = round(time.time() - t0, 3)
    b = ht.arrays()
    out["identical"] = all(np.array_equal(np.asarray(a[k]).view(np.uint32) if np.asarray(a[k]).dtype == np.float32 else a[k],
                                          np.asarray(b[k]).view(np.uint32) if np.asarray(b[k]).dtype == np.float32 else b[k])
                           for k in a if hasattr(a[k], "shape"))
print(json.dumps(out))
