"""CPU-only: the C-ABI library loads, exports every symbol include/scht.h declares, validates arguments, and has no
CPU fallback (without a GPU every device entry point fails with SCHT_ERR_NO_DEVICE)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "scht.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(scht_[a-z_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported(pkg):
    lib = C.CDLL(pkg.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 17
    for n in names:
        assert hasattr(lib, n), "libscht_b200.so does not export " + n
    assert sorted(pkg.SYMBOLS) == names, "python binding table out of sync with include/scht.h"
    assert pkg.lib().scht_abi_version() == 2


def test_product_does_not_link_the_oracle(pkg):
    import subprocess
    out = subprocess.run(["ldd", pkg.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in out
    syms = subprocess.run(["nm", "-D", pkg.LIB_PATH], capture_output=True, text=True).stdout
    assert "scht_oracle" not in syms


def test_build_tree_argument_errors(pkg):
    data = np.random.default_rng(0).random((100, 4), dtype=np.float32)
    for pct in (0.0, 0.5, 0.7, -1.0):  # Code/Convex_Tree.h:522-523 throws "illegal percent"
        with pytest.raises(pkg.SchtError) as e:
            pkg.HostTree(data, pct)
        assert e.value.code == -1
    assert "percent" in str(e.value)


def test_no_cpu_fallback_without_gpu(pkg):
    if pkg.lib().scht_device_count() > 0:
        pytest.skip("a GPU is visible")
    data = np.random.default_rng(0).random((500, 4), dtype=np.float32)
    ht = pkg.HostTree(data, 0.1)
    with pytest.raises(pkg.SchtError) as e:
        pkg.Index(ht)
    assert e.value.code == -3  # SCHT_ERR_NO_DEVICE


def test_host_tree_invariants(pkg):
    """every point lies inside every half-space of its leaf (check_node_valid, Code/Convex_Tree.h:862-944)"""
    rng = np.random.default_rng(5)
    data = rng.random((3000, 6), dtype=np.float32)
    ht = pkg.HostTree(data, 0.01, seed=3)
    a = ht.arrays()
    assert sorted(a["sorted_to_orig"].tolist()) == list(range(3000))
    assert np.array_equal(a["sorted_points"], data[a["sorted_to_orig"]])
    assert int(a["leaf_count"].sum()) == 3000 and int(a["leaf_count"].max()) <= a["leaf_threshold"]
    for leaf in range(a["n_leaves"]):
        node = int(a["leaf_node_id"][leaf])
        pts = a["sorted_points"][a["leaf_start"][leaf]: a["leaf_start"][leaf] + a["leaf_count"][leaf]].astype(np.float64)
        b0, b1 = a["node_beta_off"][node], a["node_beta_off"][node + 1]
        cur, j = node, b1 - b0 - 1
        while j >= 0:
            v = pts @ a["node_normal"][cur].astype(np.float64) + float(a["node_beta"][b0 + j])
            assert v.min() > -1e-4
            cur = int(a["node_parent"][cur])
            j -= 1
