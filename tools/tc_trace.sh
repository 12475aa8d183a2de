#!/bin/bash
# Builds a profiling variant of the library (-DSCHT_TC_PROF: clock64 stamps of the k_scan_tc pipeline roles for 40 pages of
# tile 0) into tools/bin/ and runs one cfg2 step with it; the stamps go to stderr.  GPU box only.
set -e
cd "$(dirname "$0")/.."
mkdir -p tools/bin/prof
NV="nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -DSCHT_TC_PROF"
C=semi-convex_hull_tree_b200/csrc
$NV -c -o tools/bin/prof/scht_device.o $C/scht_device.cu
$NV -c -o tools/bin/prof/scht_build.o $C/scht_build.cu
g++ -O2 -std=c++17 -fPIC -ffp-contract=off -c -o tools/bin/prof/tree_builder.o $C/tree_builder.cpp
g++ -O2 -std=c++17 -fPIC -c -o tools/bin/prof/scht_error.o $C/scht_error.cpp
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o tools/bin/libscht_prof.so tools/bin/prof/*.o -cudart shared
for dbg in ${SCHT_DBG_LIST:-0 1 2}; do SCHT_TC_DEBUG=$dbg timeout 90 python - 2>&1 <<'PY'
import importlib, sys, os
import numpy as np
sys.path.insert(0, os.getcwd())
pkg = importlib.import_module("semi-convex_hull_tree_b200")
pkg._capi.LIB_PATH = os.path.join(os.getcwd(), "tools/bin/libscht_prof.so")
rng = np.random.default_rng(1234)
data = rng.random((1_000_000, 16), dtype=np.float32)
q = np.random.default_rng(2234).random((100_000, 16), dtype=np.float32)
ht = pkg.HostTree(data, 0.0005, 1, device=0)
ix = pkg.Index(ht)
for _ in range(2):
    idx, d2, st = ix.query(q, 10, want_stats=True)
print("SCHT_TC_DEBUG", os.environ.get("SCHT_TC_DEBUG"), "main_scan_ms", round(st["ms_main_scan"], 3))
PY
done
