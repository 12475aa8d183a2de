#!/bin/bash
# Timing experiments on k_scan_tc: builds variants of the library with -DSCHT_EXP=<bits> into tools/bin/ and runs one cfg2
# step with each.  bits: 1 = survivors are never pushed, 2 = no min tree, 4 = rescorers poll every 20 us.  GPU box only.
set -e
cd "$(dirname "$0")/.."
C=semi-convex_hull_tree_b200/csrc
for e in ${SCHT_EXP_LIST:-0 1 3 5}; do
  mkdir -p tools/bin/exp$e
  NV="nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -DSCHT_EXP=$e $SCHT_EXP_FLAGS"
  $NV -c -o tools/bin/exp$e/scht_device.o $C/scht_device.cu
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o tools/bin/libscht_exp$e.so tools/bin/exp$e/scht_device.o $C/build/scht_build.o $C/build/tree_builder.o $C/build/scht_error.o $C/build/scht_io.o -cudart shared
done
for e in ${SCHT_EXP_LIST:-0 1 3 5}; do
SCHT_EXP_ID=$e timeout 90 python - 2>&1 <<'PY'
import importlib, sys, os
import numpy as np
sys.path.insert(0, os.getcwd())
pkg = importlib.import_module("semi-convex_hull_tree_b200")
pkg._capi.LIB_PATH = os.path.join(os.getcwd(), "tools/bin/libscht_exp%s.so" % os.environ["SCHT_EXP_ID"])
rng = np.random.default_rng(1234)
data = rng.random((1_000_000, 16), dtype=np.float32)
q = np.random.default_rng(2234).random((100_000, 16), dtype=np.float32)
ht = pkg.HostTree(data, 0.0005, 1, device=0)
ix = pkg.Index(ht)
best = 1e9
for _ in range(3):
    idx, d2, st = ix.query(q, 10, want_stats=True)
    best = min(best, st["ms_main_scan"])
print("SCHT_EXP", os.environ["SCHT_EXP_ID"], os.environ.get("SCHT_EXP_FLAGS", ""), "main_scan_ms", round(best, 3), "seed_ms", round(st["ms_seed_scan"], 3), "device_ms", round(st["ms_device"], 3), "exact", st["exact_reevaluations"])
PY
done
