"""The C++ host side: the standalone mirror of the reference interface (host/B200_Convex_Tree.h) and the drop-in
subclass compiled against the reference's own headers (integration/B200_Convex_Tree.h via oracle/ref_b200_harness.cc)."""
import os
import subprocess
import tempfile

import numpy as np
import pytest

from conftest import ROOT, assert_bit_equal

MIRROR_BIN = os.path.join(ROOT, "tests", "cpp", "bin", "test_host_mirror")
DROPIN_BIN = os.path.join(ROOT, "oracle", "_ref", "ref_b200_knn")


def _build_mirror(pkg):
    if not os.path.exists(MIRROR_BIN):
        subprocess.run(["make", "-C", os.path.join(ROOT, "tests", "cpp")], check=True, capture_output=True)


def test_host_mirror_without_gpu_throws(pkg):
    """CPU: the mirror class compiles against include/scht.h and refuses to work without a GPU (no fallback)."""
    if pkg.lib().scht_device_count() > 0:
        pytest.skip("a GPU is visible")
    _build_mirror(pkg)
    r = subprocess.run([MIRROR_BIN], capture_output=True, text=True)
    assert r.returncode == 77, r.stdout + r.stderr


@pytest.mark.gpu
def test_host_mirror_on_gpu(pkg):
    _build_mirror(pkg)
    r = subprocess.run([MIRROR_BIN], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "host mirror ok" in r.stdout
    assert "the dim of query points is 13, while the dim of data set is 12" in r.stdout  # reference's message


@pytest.mark.gpu
@pytest.mark.parametrize("brute", [0, 1])
def test_dropin_subclass_with_reference_host_tree(pkg, orc, golden, brute):
    """Convex_Tree<float>* t = new B200_Convex_Tree<float>(data, pct): the REFERENCE's constructor builds the tree, our
    back-end answers kNN() through the reference's virtual hooks.  Same bits as the reference's CPU_Convex_Tree."""
    if not os.path.exists(DROPIN_BIN):
        pytest.skip("oracle/_ref/ref_b200_knn not built (needs the reference sources at build time)")
    from io_formats import read_records, write_fmat
    K = golden["K"]
    with tempfile.TemporaryDirectory() as td:
        dp, qp, op = [os.path.join(td, n) for n in ("d.fmat", "q.fmat", "o.knn")]
        write_fmat(dp, golden["data"])
        write_fmat(qp, golden["queries"])
        r = subprocess.run([DROPIN_BIN, dp, repr(golden["pct"]), str(golden["seed"]), qp, str(K), str(brute), op], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        out = read_records(op)
    nq = len(golden["queries"])
    if brute:
        want_idx, want_d2, want_d = orc.brute_force(golden["data"], golden["queries"], K)
    else:
        want_idx, want_d2, want_d = golden["alg2_idx"], golden["alg2_dist2"], golden["alg2_dist"]
    assert_bit_equal(out["idx"].reshape(nq, K), want_idx, "idx")
    assert_bit_equal(out["dist2"].reshape(nq, K), want_d2, "dist2")
    assert_bit_equal(out["dist"].reshape(nq, K), want_d, "dist")


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [0, 1])
def test_dropin_subclass_multi_gpu(pkg, orc, mode):
    """the drop-in with a device list (every visible GPU, at most 4): Convex_Tree<float>::kNN() drives them all from one C++
    process (mode 0 = queries sharded, 1 = data set sharded); a repeated call replaces the results; same bits as the oracle"""
    if not os.path.exists(DROPIN_BIN):
        pytest.skip("oracle/_ref/ref_b200_knn not built (needs the reference sources at build time)")
    from conftest import load_golden
    from io_formats import read_records, write_fmat
    golden = load_golden("uniform16")
    K, nq = golden["K"], len(golden["queries"])
    n_gpus = min(4, pkg.lib().scht_device_count())
    with tempfile.TemporaryDirectory() as td:
        dp, qp, op = [os.path.join(td, n) for n in ("d.fmat", "q.fmat", "o.knn")]
        write_fmat(dp, golden["data"])
        write_fmat(qp, golden["queries"])
        r = subprocess.run([DROPIN_BIN, dp, repr(golden["pct"]), str(golden["seed"]), qp, str(K), "0", op, str(n_gpus), str(mode), "2"],
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        assert "call_ms=" in r.stderr
        out = read_records(op)
    assert_bit_equal(out["idx"].reshape(nq, K), golden["alg2_idx"], "idx")
    assert_bit_equal(out["dist2"].reshape(nq, K), golden["alg2_dist2"], "dist2")
    assert_bit_equal(out["dist"].reshape(nq, K), golden["alg2_dist"], "dist")
