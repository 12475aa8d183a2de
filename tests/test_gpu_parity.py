"""GPU parity tests (-m gpu): the CUDA path, called through the C ABI, against the reference's golden outputs and the
CPU oracle.  Bar: neighbour indices, squared distances and distances bit-exact."""
import numpy as np
import pytest

from conftest import assert_bit_equal, load_golden

pytestmark = pytest.mark.gpu


def clusters(rng, n, dim, k, sigma, scale=1.0):
    c = rng.random((k, dim))
    return ((c[rng.integers(0, k, n)] + sigma * rng.standard_normal((n, dim))) * scale).astype(np.float32)


def test_library_is_loaded_and_gpu_visible(pkg):
    assert pkg.lib().scht_device_count() >= 1


def test_golden_host_tree(pkg, golden):
    """tree from the product's builder -> scht_query == reference CPU_Convex_Tree outputs"""
    ht = pkg.HostTree(golden["data"], golden["pct"], golden["seed"])
    ix = pkg.Index(ht)
    idx, d2, d, st = ix.query(golden["queries"], golden["K"], want_dist=True, want_stats=True)
    assert_bit_equal(idx, golden["alg2_idx"], "idx")
    assert_bit_equal(d2, golden["alg2_dist2"], "dist2")
    assert_bit_equal(d, golden["alg2_dist"], "dist")
    assert st["n_queries"] == len(golden["queries"]) and st["dist_computations"] > 0 and st["kernel_launches"] >= 5
    ix.close()


def test_golden_reference_tree(pkg, golden):
    """tree flattened from the reference's own Convex_Tree (dumped by oracle/ref_harness.cc) -> same outputs"""
    ix = pkg.Index(golden["tree"])
    idx, d2 = ix.query(golden["queries"], golden["K"])
    assert_bit_equal(idx, golden["alg2_idx"], "idx")
    assert_bit_equal(d2, golden["alg2_dist2"], "dist2")
    ix.close()


def test_golden_approximate_leaves(pkg, golden):
    """device descent == approximate_kNN's descent (Code/CPU_Convex_Tree.h:181-197)"""
    ix = pkg.Index(golden["tree"])
    leaves = ix.find_approximate_leaves(golden["queries"])
    want = golden["tree"]["node_leaf_index"][golden["alg2_approx_leaf_node"]]
    assert_bit_equal(leaves, want.astype(np.int32))
    ix.close()


def test_golden_brute_force(pkg, orc, golden):
    """scht_brute_force == every row in original order through NNResult::update (Code/main.cc:59-68)"""
    ix = pkg.Index(golden["tree"])
    idx, d2, d = ix.brute_force(golden["queries"], golden["K"], want_dist=True)
    oi, od2, od = orc.brute_force(golden["data"], golden["queries"], golden["K"])
    assert_bit_equal(idx, oi, "idx")
    assert_bit_equal(d2, od2, "dist2")
    assert_bit_equal(d, od, "dist")
    ix.close()


@pytest.mark.parametrize("n,dim,nq,K,pct,kind", [
    (60000, 16, 3000, 10, 0.0025, "uniform"),
    (60000, 16, 1000, 1, 0.0025, "uniform"),
    (40000, 32, 1500, 100, 0.005, "gauss"),
    (30000, 15, 2000, 10, 0.005, "posture"),
    (20000, 128, 500, 10, 0.01, "sift"),
    (50000, 3, 2000, 10, 0.002, "uniform"),
    (20000, 17, 700, 33, 0.01, "gauss"),
])
def test_random_vs_oracle(pkg, orc, n, dim, nq, K, pct, kind):
    rng = np.random.default_rng(n * 31 + dim)
    if kind == "uniform":
        data, q = rng.random((n, dim), dtype=np.float32), rng.random((nq, dim), dtype=np.float32)
    elif kind == "gauss":
        data, q = clusters(rng, n, dim, 50, 0.05), clusters(rng, nq, dim, 50, 0.05)
    elif kind == "posture":
        data = clusters(rng, n, dim, 64, 0.05, 1e5)
        q = data[rng.integers(0, n, nq)].copy()  # self-queries: d2 == 0 at rank 0
    else:
        data = np.clip(np.rint(clusters(rng, n, dim, 20, 0.1, 100.0)), 0, 218).astype(np.float32)  # integers: real ties
        q = np.clip(np.rint(clusters(rng, nq, dim, 20, 0.1, 100.0)), 0, 218).astype(np.float32)
    ht = pkg.HostTree(data, pct, seed=1)
    oi, od2, od, _, cnt = orc.knn(ht, q, K, alg=2)
    ix = pkg.Index(ht)
    idx, d2, d, st = ix.query(q, K, want_dist=True, want_stats=True)
    assert_bit_equal(idx, oi, "idx")
    assert_bit_equal(d2, od2, "dist2")
    assert_bit_equal(d, od, "dist")
    # device-resident entry point, small batches
    ix.set_batch_size(777)
    idx2, d22 = ix.query(q, K)
    assert_bit_equal(idx2, oi, "idx (batched)")
    assert_bit_equal(d22, od2, "dist2 (batched)")
    ix.close()


def test_edge_cases(pkg, orc):
    rng = np.random.default_rng(11)
    data = rng.random((700, 5), dtype=np.float32)
    ht = pkg.HostTree(data, 0.1, seed=1)
    ix = pkg.Index(ht)
    # empty batch
    idx, d2 = ix.query(np.zeros((0, 5), np.float32), 10)
    assert idx.shape == (0, 10)
    # single query, ragged tile
    for nq in (1, 7, 63, 65):
        q = rng.random((nq, 5), dtype=np.float32)
        oi, od2, _, _, _ = orc.knn(ht, q, 10, alg=2)
        idx, d2 = ix.query(q, 10)
        assert_bit_equal(idx, oi)
        assert_bit_equal(d2, od2)
    # K larger than a leaf and K = SCHT_MAX_K
    q = rng.random((40, 5), dtype=np.float32)
    for K in (64, 128):
        oi, od2, _, _, _ = orc.knn(ht, q, K, alg=2)
        idx, d2 = ix.query(q, K)
        assert_bit_equal(idx, oi)
        assert_bit_equal(d2, od2)
    # dimension mismatch (Code/Convex_Tree.h:458-461) and bad K
    with pytest.raises(pkg.SchtError) as e:
        ix.query(np.zeros((3, 4), np.float32), 10)
    assert e.value.code == -2
    for K in (0, 129):
        with pytest.raises(pkg.SchtError) as e:
            ix.query(q, K)
        assert e.value.code == -1
    ix.close()
    # fewer points than K: unfilled slots are (-1, INFINITE)
    small = rng.random((12, 3), dtype=np.float32)
    hs = pkg.HostTree(small, 0.4, seed=1)
    ixs = pkg.Index(hs)
    idx, d2 = ixs.query(rng.random((5, 3), dtype=np.float32), 20)
    assert (idx[:, 12:] == -1).all() and (d2[:, 12:] == np.float32(3.402823466e+38)).all()
    assert (np.sort(idx[:, :12], axis=1) == np.arange(12)).all()
    ixs.close()


def test_duplicate_points_and_exact_hits(pkg, orc):
    """duplicated rows (below the leaf threshold) and queries equal to data points: ties at d2 == 0"""
    rng = np.random.default_rng(3)
    base = rng.random((3000, 8), dtype=np.float32)
    data = np.concatenate([base, base[:500], base[:200]])
    ht = pkg.HostTree(data, 0.01, seed=1)
    q = base[:300].copy()
    oi, od2, _, _, _ = orc.knn(ht, q, 10, alg=2)
    ix = pkg.Index(ht)
    idx, d2 = ix.query(q, 10)
    assert_bit_equal(idx, oi)
    assert_bit_equal(d2, od2)
    assert (d2[:, 0] == 0).all()
    ix.close()


def test_full_size_properties(pkg):
    """BASELINE config 2 size (1M x 16, K=10): size-independent properties instead of a full oracle run —
    tree query == exhaustive device scan on tie-free data, self-queries find themselves at distance 0,
    lists are sorted, and a 200-query subsample matches the CPU oracle bit for bit."""
    import oracle as orc
    rng = np.random.default_rng(2)
    n, dim, nq, K = 1_000_000, 16, 20_000, 10
    data = rng.random((n, dim), dtype=np.float32)
    q = rng.random((nq, dim), dtype=np.float32)
    q[:1000] = data[:1000]
    ht = pkg.HostTree(data, 0.0005, seed=1)
    ix = pkg.Index(ht)
    idx, d2, st = ix.query(q, K, want_stats=True)
    bi, bd2 = ix.brute_force(q, K)
    assert_bit_equal(idx, bi, "tree vs exhaustive idx")
    assert_bit_equal(d2, bd2, "tree vs exhaustive dist2")
    assert (np.diff(d2, axis=1) >= 0).all()
    assert (idx[:1000, 0] == np.arange(1000)).all() and (d2[:1000, 0] == 0).all()
    oi, od2, _, _, _ = orc.knn(ht, q[5000:5200], K, alg=2)
    assert_bit_equal(idx[5000:5200], oi)
    assert_bit_equal(d2[5000:5200], od2)
    assert st["dist_computations"] <= n * nq * 1.05  # the prefilter pass re-reads the seed pages
    ix.close()


@pytest.mark.parametrize("tc", ["0", "1"])
@pytest.mark.parametrize("n,dim,nq,K,pct,kind", [
    (60000, 16, 3000, 10, 0.0025, "uniform"),
    (40000, 32, 1000, 100, 0.005, "gauss"),
    (30000, 15, 1000, 10, 0.005, "posture"),
    (20000, 128, 300, 10, 0.01, "sift"),
    (30000, 3, 1000, 1, 0.002, "uniform"),
    (3000, 8, 200, 128, 0.05, "uniform"),
])
def test_prefilter_on_and_off(pkg, orc, monkeypatch, tc, n, dim, nq, K, pct, kind):
    """SCHT_TC=1 forces the tensor-core prefilter (even where its margin is useless, e.g. 1e5-scaled data: everything
    becomes a candidate and must still come out exact); SCHT_TC=0 forces the FP32 scan.  Same bits either way."""
    monkeypatch.setenv("SCHT_TC", tc)
    rng = np.random.default_rng(n + dim + K)
    if kind == "uniform":
        data, q = rng.random((n, dim), dtype=np.float32), rng.random((nq, dim), dtype=np.float32)
    elif kind == "gauss":
        data, q = clusters(rng, n, dim, 50, 0.05), clusters(rng, nq, dim, 50, 0.05)
    elif kind == "posture":
        data = clusters(rng, n, dim, 64, 0.05, 1e5)
        q = data[rng.integers(0, n, nq)].copy()
    else:
        data = np.clip(np.rint(clusters(rng, n, dim, 20, 0.1, 100.0)), 0, 218).astype(np.float32)
        q = np.clip(np.rint(clusters(rng, nq, dim, 20, 0.1, 100.0)), 0, 218).astype(np.float32)
    ht = pkg.HostTree(data, pct, seed=1)
    oi, od2, od, _, _ = orc.knn(ht, q, K, alg=2)
    ix = pkg.Index(ht)
    idx, d2, d = ix.query(q, K, want_dist=True)
    assert_bit_equal(idx, oi, "idx")
    assert_bit_equal(d2, od2, "dist2")
    assert_bit_equal(d, od, "dist")
    ix.close()


@pytest.mark.parametrize("env", [{"SCHT_SCAN_VARIANT": "0"}, {"SCHT_SCAN_VARIANT": "1"}, {"SCHT_SCAN_VARIANT": "2"}, {"SCHT_SCAN_VARIANT": "3"}, {"SCHT_SCAN_STAGES": "2"}, {"SCHT_SCAN_STAGES": "4"},
                                 {"SCHT_TRAVERSE_SAMPLING": "0"}, {"SCHT_TRAVERSE_SAMPLING": "0", "SCHT_TC": "1"},
                                 {"SCHT_TC": "1", "SCHT_TC_SPLIT": "0"}, {"SCHT_TC": "1", "SCHT_TC_SPLIT": "3"},
                                 {"SCHT_TC": "1", "SCHT_TC_SPLIT": "8"}, {"SCHT_SEED_POINTS": "256"}, {"SCHT_SEED_POINTS": "8192"},
                                 {"SCHT_TC": "1", "SCHT_TC_N256": "1"}, {"SCHT_TC": "1", "SCHT_TC_N256": "0"},
                                 {"SCHT_TC": "1", "SCHT_TC_ISSUERS": "4"}, {"SCHT_TC": "1", "SCHT_TC_ISSUERS": "2", "SCHT_TC_N256": "0"}])
def test_kernel_variants_and_knobs(pkg, orc, monkeypatch, env):
    """every leaf-scan kernel variant (8/7/4 consumer warps, 2- and 4-stage rings), the traversal with sampling switched off, prefilter tiles
    split into page segments (partial lists merged), other seed-neighbourhood sizes and the MMA issue shapes (two or four
    issuers, half-page or whole-page chains; the test data has 16-D one-slice and 20-D two-slice pages) give the
    reference's bits (the knobs only change the schedule)"""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    rng = np.random.default_rng(77)
    n, nq, K = 50000, 9000, 10  # 9000 queries: enough tiles for the sampled traversal to engage when it is on
    for dim, data, q in ((16, rng.random((n, 16), dtype=np.float32), rng.random((nq, 16), dtype=np.float32)),
                         (20, clusters(rng, n, 20, 40, 0.05), clusters(rng, nq, 20, 40, 0.05))):
        ht = pkg.HostTree(data, 0.004, seed=1)
        oi, od2, _, _, _ = orc.knn(ht, q, K, alg=2)
        ix = pkg.Index(ht)
        idx, d2, st = ix.query(q, K, want_stats=True)
        assert_bit_equal(idx, oi, "idx dim=%d" % dim)
        assert_bit_equal(d2, od2, "dist2 dim=%d" % dim)
        if env.get("SCHT_TC") == "1":
            assert st["prefilter_batches"] == 1
        ix.close()


def test_stats_surface(pkg):
    """the statistics the reference prints (Code/GPU_Convex_Tree.h:950-971) are filled in"""
    rng = np.random.default_rng(5)
    data, q = rng.random((30000, 8), dtype=np.float32), rng.random((3000, 8), dtype=np.float32)
    ix = pkg.Index(pkg.HostTree(data, 0.005, seed=1))
    ix.set_batch_size(1000)
    _, _, st = ix.query(q, 10, want_stats=True)
    assert st["n_queries"] == 3000 and st["batches"] == 3 and st["kernel_launches"] >= 21
    assert st["dist_computations"] > 0 and st["hyperplane_tests"] > 0 and st["descent_dists"] > 0 and st["list_insertions"] >= 30000
    assert st["ms_total"] >= st["ms_device"] > 0 and abs(st["ms_scan"] - st["ms_seed_scan"] - st["ms_main_scan"]) < 1e-3
    ix.close()


@pytest.mark.parametrize("dim", [16, 40])
def test_sample_reuse_and_wide_rows(pkg, monkeypatch, dim):
    """repeated batches of one shape skip the traversal sample after two "list every page" verdicts (SCHT_SAMPLE_REUSE);
    wide points (dim > 16) are re-evaluated from the row-major copy: every call must equal the exhaustive device scan"""
    rng = np.random.default_rng(77 + dim)
    data, q = rng.random((60000, dim), dtype=np.float32), rng.random((9000, dim), dtype=np.float32)
    monkeypatch.setenv("SCHT_TC", "1")
    ix = pkg.Index(pkg.HostTree(data, 0.004, seed=1))
    bi, bd2 = ix.brute_force(q, 10)
    tests = []
    for rep in range(4):
        idx, d2, st = ix.query(q, 10, want_stats=True)
        assert_bit_equal(idx, bi, "idx rep=%d" % rep)
        assert_bit_equal(d2, bd2, "dist2 rep=%d" % rep)
        assert st["prefilter_batches"] == 1
        tests.append(st["hyperplane_tests"])
    assert tests[0] == tests[1] and tests[2] <= tests[0] and tests[3] <= tests[0]
    ix.close()
    monkeypatch.setenv("SCHT_SAMPLE_REUSE", "0")
    ix = pkg.Index(pkg.HostTree(data, 0.004, seed=1))
    for rep in range(3):
        idx, d2, st = ix.query(q, 10, want_stats=True)
        assert_bit_equal(idx, bi, "idx (no reuse) rep=%d" % rep)
        assert st["hyperplane_tests"] == tests[0]
    ix.close()
