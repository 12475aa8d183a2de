"""CPU-only: pins the oracle (oracle/scht_oracle.c) and the host tree builder (csrc/tree_builder.cpp) against
fixtures produced by the UNMODIFIED reference (tests/golden/make_golden.py), and — when the compiled reference
is present (oracle/_ref/ref_knn) — against the reference run live on fresh seeded inputs."""
import numpy as np
import pytest

from conftest import assert_bit_equal

TREE_KEYS = ["sorted_points", "sorted_to_orig", "leaf_start", "leaf_count", "leaf_node_id", "node_left", "node_right", "node_parent",
             "node_leaf_index", "node_left_center", "node_right_center", "node_normal", "node_beta_off", "node_beta"]


def test_builder_matches_reference_tree(pkg, golden):
    """scht_build_tree == Convex_Tree<T> ctor (Code/Convex_Tree.h:518-548,708-857,593-633), bit for bit."""
    ht = pkg.HostTree(golden["data"], golden["pct"], golden["seed"])
    a, r = ht.arrays(), golden["tree"]
    for k in ("dim", "n_points", "n_nodes", "n_leaves", "depth", "leaf_threshold"):
        assert int(a[k]) == int(r[k]), k
    for k in TREE_KEYS:
        assert_bit_equal(a[k], r[k], k)


@pytest.mark.parametrize("alg", [2, 0])
def test_oracle_matches_reference_outputs(pkg, orc, golden, alg):
    """scht_oracle_knn == CPU_Convex_Tree::do_kNN (Code/CPU_Convex_Tree.h:131-152, :380-430 / :294-376)."""
    idx, d2, d, appr, cnt = orc.knn(golden["tree"], golden["queries"], golden["K"], alg=alg)
    assert_bit_equal(idx, golden["alg%d_idx" % alg], "idx")
    assert_bit_equal(d2, golden["alg%d_dist2" % alg], "dist2")
    assert_bit_equal(d, golden["alg%d_dist" % alg], "dist")
    assert_bit_equal(appr, golden["alg%d_approx_leaf_node" % alg], "approx leaf")
    assert cnt[:3].tolist() == golden["alg%d_counters" % alg].tolist()  # dist computations, hyperplane tests, dot products


def test_both_reference_algorithms_agree(golden):
    assert_bit_equal(golden["alg0_idx"], golden["alg2_idx"])
    assert_bit_equal(golden["alg0_dist2"], golden["alg2_dist2"])


def test_contract_matches_reference(pkg, orc, golden):
    """The order-independent contract (K smallest (d2, approx-leaf-first, sorted position)) reproduces the reference."""
    idx, d2 = orc.contract(golden["tree"], golden["queries"], golden["K"])
    assert_bit_equal(idx, golden["alg2_idx"], "idx")
    assert_bit_equal(d2, golden["alg2_dist2"], "dist2")


def test_oracle_on_host_tree_equals_oracle_on_reference_tree(pkg, orc, golden):
    ht = pkg.HostTree(golden["data"], golden["pct"], golden["seed"])
    idx, d2, _, _, _ = orc.knn(ht, golden["queries"], golden["K"], alg=2)
    assert_bit_equal(idx, golden["alg2_idx"])
    assert_bit_equal(d2, golden["alg2_dist2"])


def test_brute_force_oracle_agrees_on_tie_free_data(orc):
    g = __import__("conftest").load_golden("uniform16")
    idx, d2, _ = orc.brute_force(g["data"], g["queries"], g["K"])
    assert_bit_equal(idx, g["alg2_idx"])
    assert_bit_equal(d2, g["alg2_dist2"])


@pytest.mark.parametrize("n,dim,nq,K,pct,seed", [(6000, 10, 80, 10, 0.005, 1), (5000, 16, 60, 30, 0.01, 7), (2500, 4, 50, 3, 0.02, 3)])
def test_live_reference(pkg, orc, n, dim, nq, K, pct, seed):
    """Fresh inputs through the compiled reference (only where oracle/_ref/ref_knn exists)."""
    if not orc.have_reference():
        pytest.skip("oracle/_ref/ref_knn not built (no reference sources on this box)")
    rng = np.random.default_rng(n + dim)
    data = (rng.random((n, dim)) * 1000).astype(np.float32)  # test_on_random: rand()/1000-style magnitudes (Code/main.cc:110-139)
    q = (rng.random((nq, dim)) * 1000).astype(np.float32)
    ref = orc.run_reference(data, q, K, pct, seed=seed, alg=2, want_tree=True)
    ht = pkg.HostTree(data, pct, seed)
    a = ht.arrays()
    for k in TREE_KEYS:
        assert_bit_equal(a[k], ref["tree"][k], k)
    idx, d2, d, appr, cnt = orc.knn(ht, q, K, alg=2)
    assert_bit_equal(idx, ref["idx"])
    assert_bit_equal(d2, ref["dist2"])
    assert_bit_equal(d, ref["dist"])
    assert cnt[:3].tolist() == ref["counters"].tolist()
