"""A/B of two builds of the library on one cfg2 step (uniform 1Mx16, 100k queries, K=10), alternating, best of N.
usage: python tools/ab_cfg2.py libA.so libB.so   (GPU box only; each library runs in its own process)"""
import os, subprocess, sys
CODE = r'''
import importlib, os, sys
import numpy as np
sys.path.insert(0, os.getcwd())
pkg = importlib.import_module("semi-convex_hull_tree_b200")
pkg._capi.LIB_PATH = os.environ["SCHT_AB_LIB"]
data = np.random.default_rng(1234).random((1_000_000, 16), dtype=np.float32)
q = np.random.default_rng(2234).random((100_000, 16), dtype=np.float32)
ix = pkg.Index(pkg.HostTree(data, 0.0005, 1, device=0))
ms = []
for _ in range(6):
    idx, d2, st = ix.query(q, 10, want_stats=True)
    ms.append((st["ms_main_scan"], st["ms_seed_scan"], st["ms_device"]))
print(os.path.basename(os.environ["SCHT_AB_LIB"]), "main/seed/device ms, best of 6:", [round(min(m[i] for m in ms), 3) for i in range(3)])
'''
for rnd in range(2):
    for lib in sys.argv[1:]:
        subprocess.run([sys.executable, "-c", CODE], env=dict(os.environ, SCHT_AB_LIB=os.path.abspath(lib)), check=False)
