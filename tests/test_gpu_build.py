"""GPU tests (-m gpu) of the device tree builder (scht_build_tree_device): every flat array must be
bit-identical to the host builder's, which tests/test_oracle_golden.py pins against the reference's own tree."""
import numpy as np
import pytest

from conftest import assert_bit_equal

pytestmark = pytest.mark.gpu

ARRAYS = ("sorted_points", "sorted_to_orig", "leaf_start", "leaf_count", "leaf_node_id", "node_left", "node_right", "node_parent",
          "node_leaf_index", "node_left_center", "node_right_center", "node_normal", "node_beta_off", "node_beta")
SCALARS = ("dim", "n_points", "n_nodes", "n_leaves", "depth", "leaf_threshold")


def assert_same_tree(a, b):
    for k in SCALARS:
        assert a[k] == b[k], "%s: %r vs %r" % (k, a[k], b[k])
    for k in ARRAYS:
        assert_bit_equal(a[k], b[k], k)


def clusters(rng, n, dim, k, sigma, scale=1.0):
    c = rng.random((k, dim))
    return ((c[rng.integers(0, k, n)] + sigma * rng.standard_normal((n, dim))) * scale).astype(np.float32)


def test_golden_trees(pkg, golden):
    """device-built tree == the reference's own dumped tree (golden fixture)"""
    dev = pkg.HostTree(golden["data"], golden["pct"], golden["seed"], device=0)  # keep alive: arrays() are borrowed views
    dt = dev.arrays()
    ref = golden["tree"]
    for k in ARRAYS:
        if k in ref:
            assert_bit_equal(dt[k], ref[k], k)


@pytest.mark.parametrize("n,dim,pct,kind,seed", [
    (60000, 16, 0.0025, "uniform", 1),
    (200000, 16, 0.0005, "uniform", 7),
    (40000, 32, 0.005, "gauss", 1),
    (30000, 15, 0.005, "posture", 3),
    (20000, 128, 0.01, "sift", 1),
    (50000, 3, 0.002, "uniform", 1),
    (20000, 17, 0.01, "gauss", 12345),
    (5000, 5, 0.4, "uniform", 1),       # two-level tree
    (700, 4, 0.0005, "uniform", 1),     # threshold 0: every point its own leaf
])
def test_device_tree_equals_host_tree(pkg, n, dim, pct, kind, seed):
    rng = np.random.default_rng(n * 7 + dim)
    if kind == "uniform":
        data = rng.random((n, dim), dtype=np.float32)
    elif kind == "gauss":
        data = clusters(rng, n, dim, 50, 0.05)
    elif kind == "posture":
        data = clusters(rng, n, dim, 64, 0.05, 1e5)
    else:
        data = np.clip(np.rint(clusters(rng, n, dim, 20, 0.1, 100.0)), 0, 218).astype(np.float32)  # integer data: tied distances
    host = pkg.HostTree(data, pct, seed)
    dev = pkg.HostTree(data, pct, seed, device=0)
    assert_same_tree(dev.arrays(), host.arrays())


def test_duplicate_rows_close_as_oversized_leaf(pkg):
    """more than `threshold` identical rows: the reference recurses forever; both builders close the node as a leaf"""
    rng = np.random.default_rng(5)
    data = rng.random((4000, 8), dtype=np.float32)
    data[1000:1600] = data[1000]
    host = pkg.HostTree(data, 0.05, 1)
    dev = pkg.HostTree(data, 0.05, 1, device=0)
    assert_same_tree(dev.arrays(), host.arrays())


def test_query_on_device_built_tree(pkg, orc):
    rng = np.random.default_rng(11)
    data = rng.random((50000, 16), dtype=np.float32)
    q = rng.random((1500, 16), dtype=np.float32)
    dev = pkg.HostTree(data, 0.004, 1, device=0)
    ix = pkg.Index(dev)
    idx, d2 = ix.query(q, 10)
    oi, od2, _, _, _ = orc.knn(dev, q, 10, alg=2)
    assert_bit_equal(idx, oi, "idx")
    assert_bit_equal(d2, od2, "dist2")
    ix.close()


def test_bad_arguments(pkg):
    data = np.zeros((10, 4), dtype=np.float32)
    with pytest.raises(pkg.SchtError) as e:
        pkg.HostTree(data, 0.7, 1, device=0)
    assert e.value.code == -1
    with pytest.raises(pkg.SchtError) as e:
        pkg.HostTree(data, 0.1, 1, device=99)
    assert e.value.code == -1
