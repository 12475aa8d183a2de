"""GPU tests (-m gpu) of the round-2 additions behind the C ABI: several GPUs behind one handle (scht_create_multi, both
modes; with one visible GPU the same code runs with a single device), shard handles + the two-phase sharded-dataset
calls, the brute-force entry through the tensor-core prefilter, page selection by bounding balls on clustered data,
the option interface and the pipelined multi-batch scht_query.  Bar everywhere: bit-exact against the CPU oracle."""
import numpy as np
import pytest

from conftest import assert_bit_equal

pytestmark = pytest.mark.gpu


def clusters(rng, n, dim, k, sigma, scale=1.0, seed_centres=None):
    c = (np.random.default_rng(seed_centres) if seed_centres is not None else rng).random((k, dim))
    return ((c[rng.integers(0, k, n)] + sigma * rng.standard_normal((n, dim))) * scale).astype(np.float32)


def n_devices(pkg):
    return pkg.lib().scht_device_count()


@pytest.mark.parametrize("mode", ["queries", "dataset"])
@pytest.mark.parametrize("kind,n,dim,nq,K", [("uniform", 60000, 16, 5000, 10), ("gauss", 80000, 32, 4000, 10), ("gauss", 30000, 8, 1500, 100)])
def test_multi_handle_matches_oracle(pkg, orc, mode, kind, n, dim, nq, K):
    """scht_create_multi over every visible GPU (at most 4): scht_query and scht_brute_force give the oracle's bits in both modes"""
    rng = np.random.default_rng(n + dim + K)
    if kind == "uniform":
        data, q = rng.random((n, dim), dtype=np.float32), rng.random((nq, dim), dtype=np.float32)
    else:
        data, q = clusters(rng, n, dim, 40, 0.04, seed_centres=5), clusters(rng, nq, dim, 40, 0.04, seed_centres=5)
    ht = pkg.HostTree(data, 0.004, seed=1)
    oi, od2, od, _, _ = orc.knn(ht, q, K, alg=2)
    devs = list(range(min(4, n_devices(pkg))))
    ix = pkg.Index(ht, devices=devs, mode=pkg.SHARD_QUERIES if mode == "queries" else pkg.SHARD_DATASET)
    lay = ix.layout()
    assert lay["n_devices"] == len(devs) and lay["mode"] == (0 if mode == "queries" else 1)
    idx, d2, d, st = ix.query(q, K, want_dist=True, want_stats=True)
    assert_bit_equal(idx, oi, "idx")
    assert_bit_equal(d2, od2, "dist2")
    assert_bit_equal(d, od, "dist")
    assert st["n_queries"] == nq and st["kernel_launches"] >= 5
    ix.set_batch_size(1777)  # several batches per call
    idx, d2 = ix.query(q, K)
    assert_bit_equal(idx, oi, "idx (batched)")
    assert_bit_equal(d2, od2, "dist2 (batched)")
    bi, bd2 = ix.brute_force(q[:700], K)
    obi, obd2, _ = orc.brute_force(data, q[:700], K)
    assert_bit_equal(bi, obi, "brute idx")
    assert_bit_equal(bd2, obd2, "brute dist2")
    ix.close()


@pytest.mark.parametrize("shards", [2, 3])
def test_shard_handles_two_phase(pkg, orc, shards):
    """the one-process-per-GPU form of the sharded-dataset mode, emulated with `shards` shard handles on one GPU:
    seed bounds per shard -> element-wise minimum -> shard scans -> merge == oracle"""
    import torch
    sh = __import__("importlib").import_module("semi-convex_hull_tree_b200.sharding")
    rng = np.random.default_rng(21)
    n, dim, nq, K = 70000, 16, 3000, 10
    data, q = clusters(rng, n, dim, 30, 0.05, seed_centres=3), clusters(rng, nq, dim, 30, 0.05, seed_centres=3)
    ht = pkg.HostTree(data, 0.004, seed=1)
    oi, od2, od, _, _ = orc.knn(ht, q, K, alg=2)
    parts = sh.leaf_shards(ht.arrays()["leaf_count"], shards)
    hs = [pkg.Index(ht, device=0, leaf_range=p) for p in parts]
    assert [h.layout()["leaf_begin"] for h in hs] == [p[0] for p in parts]
    dq = torch.from_numpy(q).cuda()
    bounds = torch.empty((shards, nq), dtype=torch.float32, device="cuda")
    keys = torch.empty((shards, nq, K), dtype=torch.int64, device="cuda")
    # the phases of one shard must stay together (state lives in the handle): seed all shards first, then scan all
    for s, h in enumerate(hs):
        h.partial_seed_device(dq.data_ptr(), nq, dim, K, bounds[s].data_ptr())
    bound = bounds.min(dim=0).values.contiguous()
    assert (bound < 3e38).all()  # every query was seeded by the shard holding its approximate leaf
    for s, h in enumerate(hs):
        h.partial_scan_device(dq.data_ptr(), nq, dim, K, bound.data_ptr(), keys[s].data_ptr())
    didx = torch.empty((nq, K), dtype=torch.int32, device="cuda")
    dd2 = torch.empty((nq, K), dtype=torch.float32, device="cuda")
    dd = torch.empty((nq, K), dtype=torch.float32, device="cuda")
    hs[0].merge_partial_device(keys.data_ptr(), shards, nq, dim, K, didx.data_ptr(), dd2.data_ptr(), dd.data_ptr())
    assert_bit_equal(didx.cpu().numpy(), oi, "idx")
    assert_bit_equal(dd2.cpu().numpy(), od2, "dist2")
    assert_bit_equal(dd.cpu().numpy(), od, "dist")
    # a scan without its seed phase is refused
    with pytest.raises(pkg.SchtError):
        hs[0].partial_scan_device(dq.data_ptr(), nq, dim, K, bound.data_ptr(), keys[0].data_ptr())
    for h in hs:
        h.close()


@pytest.mark.parametrize("n,dim,nq,K,kind", [(300000, 16, 4000, 10, "uniform"), (200000, 32, 3000, 10, "gauss"), (100000, 128, 1000, 10, "sift"),
                                             (50000, 8, 2000, 100, "uniform")])
def test_brute_force_through_prefilter(pkg, orc, n, dim, nq, K, kind):
    """scht_brute_force: tensor-core prefilter path (default) == exact FP32 path (brute_tc = 0) == oracle; ties by original index"""
    rng = np.random.default_rng(n + dim)
    if kind == "uniform":
        data, q = rng.random((n, dim), dtype=np.float32), rng.random((nq, dim), dtype=np.float32)
    elif kind == "gauss":
        data, q = clusters(rng, n, dim, 100, 0.05, seed_centres=9), clusters(rng, nq, dim, 100, 0.05, seed_centres=9)
    else:
        data = np.clip(np.rint(clusters(rng, n, dim, 20, 0.1, 100.0, seed_centres=2)), 0, 218).astype(np.float32)  # integers: real ties
        q = np.clip(np.rint(clusters(rng, nq, dim, 20, 0.1, 100.0, seed_centres=2)), 0, 218).astype(np.float32)
    ht = pkg.HostTree(data, 0.002, seed=1, device=0)
    ix = pkg.Index(ht)
    oi, od2, od = orc.brute_force(data, q[:400], K)
    idx, d2, d, st = ix.brute_force(q, K, want_dist=True, want_stats=True)
    assert st["prefilter_batches"] == 1, "the brute-force entry did not take the prefilter path"
    assert_bit_equal(idx[:400], oi, "idx vs oracle")
    assert_bit_equal(d2[:400], od2, "dist2 vs oracle")
    assert_bit_equal(d[:400], od, "dist vs oracle")
    ix.set_option("brute_tc", 0)
    eidx, ed2, st2 = ix.brute_force(q, K, want_stats=True)
    assert st2["prefilter_batches"] == 0
    assert_bit_equal(idx, eidx, "prefilter vs exact idx")
    assert_bit_equal(d2, ed2, "prefilter vs exact dist2")
    ix.close()


@pytest.mark.parametrize("dim,K", [(32, 10), (16, 1), (24, 100)])
def test_page_selection_prunes_clustered_data(pkg, orc, dim, K):
    """well separated clusters: the ball test must leave a small fraction of the scan pages (pages_listed), and the result
    must not depend on it (page_select 0/1, one scan group or many, sampling on/off)"""
    rng = np.random.default_rng(4 + dim)
    n, nq = 400000, 20000
    data, q = clusters(rng, n, dim, 200, 0.03, seed_centres=8), clusters(rng, nq, dim, 200, 0.03, seed_centres=8)
    ht = pkg.HostTree(data, 0.001, seed=1, device=0)
    oi, od2, _, _, _ = orc.knn(ht, q[:1500], K, alg=2)
    ix = pkg.Index(ht, options={"tc": 1})
    lay = ix.layout()
    assert lay["scan_groups"] > 1 and lay["super_groups"] >= 200 and (n + 255) // 256 <= lay["scan_pages"] <= 1.1 * n / 256
    idx, d2, st = ix.query(q, K, want_stats=True)
    assert_bit_equal(idx[:1500], oi, "idx")
    assert_bit_equal(d2[:1500], od2, "dist2")
    assert st["prefilter_batches"] == 1
    frac = st["pages_listed"] / (nq * lay["scan_pages"])
    assert frac < 0.06, "ball test kept %.3f of the scan pages on well separated clusters" % frac
    for opts in ({"page_select": 0}, {"traverse_sampling": 0}, {"tc_split": 4}, {"tc": 0}):
        for k, v in opts.items():
            ix.set_option(k, v)
        i2, d22 = ix.query(q, K)
        assert_bit_equal(i2, idx, "idx with %r" % opts)
        assert_bit_equal(d22, d2, "dist2 with %r" % opts)
        for k in opts:
            ix.set_option(k, {"page_select": 1, "traverse_sampling": 1, "tc_split": 1, "tc": 1}[k])
    ix.close()


@pytest.mark.parametrize("groups", ["1", "7", "300"])
def test_scan_group_counts(pkg, orc, monkeypatch, groups):
    """SCHT_GROUPS: tree-order image (1 group) or any number of scan groups -- same bits"""
    monkeypatch.setenv("SCHT_GROUPS", groups)
    monkeypatch.setenv("SCHT_TC", "1")
    rng = np.random.default_rng(12)
    n, dim, nq, K = 90000, 20, 5000, 10
    data, q = clusters(rng, n, dim, 60, 0.04, seed_centres=6), clusters(rng, nq, dim, 60, 0.04, seed_centres=6)
    ht = pkg.HostTree(data, 0.003, seed=1)
    oi, od2, _, _, _ = orc.knn(ht, q, K, alg=2)
    ix = pkg.Index(ht)
    assert ix.layout()["scan_groups"] == int(groups) and ix.get_option("groups") == int(groups)
    idx, d2, st = ix.query(q, K, want_stats=True)
    assert_bit_equal(idx, oi, "idx")
    assert_bit_equal(d2, od2, "dist2")
    assert st["prefilter_batches"] == 1
    bi, bd2 = ix.brute_force(q[:500], K)
    obi, obd2, _ = orc.brute_force(data, q[:500], K)
    assert_bit_equal(bi, obi, "brute idx")
    assert_bit_equal(bd2, obd2, "brute dist2")
    ix.close()


def test_options_interface(pkg):
    rng = np.random.default_rng(1)
    ix = pkg.Index(pkg.HostTree(rng.random((5000, 6), dtype=np.float32), 0.01, seed=1))
    assert ix.get_option("tc") == 2 and ix.get_option("seed_points") == 256
    ix.set_option("tc", 0)
    ix.set_option("seed_points", 2048)
    assert ix.get_option("tc") == 0 and ix.get_option("seed_points") == 2048
    for name, value in (("no_such_option", 1), ("tc", 7), ("groups", 4), ("tc_ring", 48), ("tc_ring", 16)):
        with pytest.raises(pkg.SchtError) as e:
            ix.set_option(name, value)
        assert e.value.code == -1
    ix.close()


def test_survivor_ring_depth(pkg, orc):
    """k_scan_tc's survivor rings (options tc_ring, tc_recheck): depth and re-check change how many stale survivors reach
    the exact evaluation, never a result bit"""
    rng = np.random.default_rng(21)
    n, dim, nq, K = 60000, 24, 3000, 10
    c = rng.random((40, dim))
    data = (c[rng.integers(0, 40, n)] + 0.03 * rng.standard_normal((n, dim))).astype(np.float32)
    q = (c[rng.integers(0, 40, nq)] + 0.03 * rng.standard_normal((nq, dim))).astype(np.float32)
    ht = pkg.HostTree(data, 0.004, seed=1)
    oi, od2, _, _, _ = orc.knn(ht, q, K, alg=2)
    ix = pkg.Index(ht)
    ix.set_option("tc", 1)
    evals = {}
    for ring, recheck in ((0, 1), (32, 1), (128, 0), (1024, 1), (1024, 0)):
        ix.set_option("tc_ring", ring)
        ix.set_option("tc_recheck", recheck)  # entries carry the prefilter value; stale ones are dropped before the exact evaluation
        idx, d2, st = ix.query(q, K, want_stats=True)
        assert st["prefilter_batches"] == 1
        assert_bit_equal(idx, oi, "idx ring=%d recheck=%d" % (ring, recheck))
        assert_bit_equal(d2, od2, "dist2 ring=%d recheck=%d" % (ring, recheck))
        evals[(ring, recheck)] = st["exact_reevaluations"]
    assert evals[(1024, 1)] <= evals[(1024, 0)]
    ix.close()


@pytest.mark.parametrize("dim,K", [(24, 10), (12, 10), (24, 100), (130, 5)])
def test_lazy_seed_scan(pkg, orc, dim, K):
    """option lazy_seed: the seed scan keeps its lists by fast-path distances and evaluates only what is left with the
    reference arithmetic; same bits as the exact seed scan, for the tree query (prefilter and exact main scan) and the
    brute-force entry; duplicates of the data set among the queries give exact ties"""
    rng = np.random.default_rng(31 + dim)
    n, nq = 50000, 2500
    c = rng.random((30, dim))
    data = (c[rng.integers(0, 30, n)] + 0.03 * rng.standard_normal((n, dim))).astype(np.float32)
    data[1000:1200] = data[2000:2200]  # duplicate points: equal distances, ties by position / original index
    q = (c[rng.integers(0, 30, nq)] + 0.03 * rng.standard_normal((nq, dim))).astype(np.float32)
    q[:100] = data[2000:2100]
    ht = pkg.HostTree(data, 0.004, seed=1)
    oi, od2, _, _, _ = orc.knn(ht, q, K, alg=2)
    obi, obd2, _ = orc.brute_force(data, q[:300], K)
    ix = pkg.Index(ht)
    for tc in (1, 0):  # prefilter forced (the seed scan is lazy from the first batch on) / exact main scan (never lazy for tree batches)
        ix.set_option("tc", tc)
        for lazy in (1, 0):
            ix.set_option("lazy_seed", lazy)
            for rep in range(2):
                idx, d2 = ix.query(q, K)
                assert_bit_equal(idx, oi, "idx tc=%d lazy=%d" % (tc, lazy))
                assert_bit_equal(d2, od2, "dist2 tc=%d lazy=%d" % (tc, lazy))
            bi, bd2 = ix.brute_force(q[:300], K)
            assert_bit_equal(bi, obi, "brute idx tc=%d lazy=%d" % (tc, lazy))
            assert_bit_equal(bd2, obd2, "brute dist2 tc=%d lazy=%d" % (tc, lazy))
    ix.close()


@pytest.mark.parametrize("pinned", [False, True])
def test_pipelined_batches(pkg, orc, pinned):
    """a scht_query call of many batches overlaps its copies with the kernels (two copy streams, pinned staging for
    pageable callers): same bits as one batch, with and without the pipeline"""
    import torch
    rng = np.random.default_rng(8)
    n, dim, nq, K = 50000, 12, 10000, 10
    data, q = rng.random((n, dim), dtype=np.float32), rng.random((nq, dim), dtype=np.float32)
    ht = pkg.HostTree(data, 0.004, seed=1)
    oi, od2, od, _, _ = orc.knn(ht, q, K, alg=2)
    ix = pkg.Index(ht)
    tq = torch.from_numpy(q)
    ti, td2, td = torch.empty((nq, K), dtype=torch.int32), torch.empty((nq, K), dtype=torch.float32), torch.empty((nq, K), dtype=torch.float32)
    if pinned:
        tq, ti, td2, td = tq.pin_memory(), ti.pin_memory(), td2.pin_memory(), td.pin_memory()
    for bs, pipe, split in ((1300, 1, 0), (1300, 0, 0), (4096, 1, 0), (1000000, 1, 3)):
        ix.set_batch_size(bs)
        ix.set_option("pipeline", pipe)
        ix.set_option("pipeline_split", split)
        ti.fill_(-7)
        st = pkg.Stats()
        ix.query_raw(tq.data_ptr(), nq, dim, K, ti.data_ptr(), td2.data_ptr(), td.data_ptr(), st)
        assert_bit_equal(ti.numpy(), oi, "idx bs=%d pipe=%d" % (bs, pipe))
        assert_bit_equal(td2.numpy(), od2, "dist2 bs=%d pipe=%d" % (bs, pipe))
        assert_bit_equal(td.numpy(), od, "dist bs=%d pipe=%d" % (bs, pipe))
        assert st.batches == (split if split > 1 else (nq + bs - 1) // bs)
    ix.close()


@pytest.mark.parametrize("dim", [256, 512])
def test_prefilter_margin_wide_points(pkg, orc, monkeypatch, dim):
    """the accumulation term of the prefilter margin grows with the number of K16 steps: forced prefilter
    on D = 256 / 512 against the oracle (the path falls back to the exact scan by itself when the operand does not fit
    in shared memory; either way the bits are the oracle's)"""
    monkeypatch.setenv("SCHT_TC", "1")
    rng = np.random.default_rng(dim)
    n, nq, K = 12000, 300, 10
    data, q = clusters(rng, n, dim, 12, 0.08, seed_centres=4), clusters(rng, nq, dim, 12, 0.08, seed_centres=4)
    ht = pkg.HostTree(data, 0.02, seed=1)
    oi, od2, _, _, _ = orc.knn(ht, q, K, alg=2)
    ix = pkg.Index(ht)
    idx, d2, st = ix.query(q, K, want_stats=True)
    assert_bit_equal(idx, oi, "idx")
    assert_bit_equal(d2, od2, "dist2")
    ix.close()


def test_tiny_magnitude_data(pkg, orc):
    """coordinates around 1e-18: s^2 of the prefilter would overflow, so the image must not be built and the
    exact scan answers; 1e-12 still takes the prefilter"""
    rng = np.random.default_rng(2)
    for mag, want_tc in ((1e-18, False), (1e-12, True)):
        data = (rng.random((20000, 8)) * mag).astype(np.float32)
        q = (rng.random((500, 8)) * mag).astype(np.float32)
        ht = pkg.HostTree(data, 0.01, seed=1)
        oi, od2, _, _, _ = orc.knn(ht, q, 10, alg=2)
        ix = pkg.Index(ht)
        idx, d2, st = ix.query(q, 10, want_stats=True)
        assert_bit_equal(idx, oi, "idx mag=%g" % mag)
        assert_bit_equal(d2, od2, "dist2 mag=%g" % mag)
        assert (ix.layout()["scan_pages"] > 0) == want_tc
        ix.close()


def test_bad_parent_array_is_harmless(pkg):
    """scht_create derives parents from the validated child links: a cyclic node_parent array must not hang"""
    rng = np.random.default_rng(3)
    data, q = rng.random((9000, 5), dtype=np.float32), rng.random((100, 5), dtype=np.float32)
    ht = pkg.HostTree(data, 0.01, seed=1)
    a = dict(ht.arrays())
    good = pkg.Index(a)
    i0, d0 = good.query(q, 5)
    good.close()
    bad = dict(a)
    bad["node_parent"] = np.roll(np.arange(a["n_nodes"], dtype=np.int32), 1)  # one big cycle
    ix = pkg.Index(bad)
    i1, d1 = ix.query(q, 5)
    assert_bit_equal(i1, i0)
    assert_bit_equal(d1, d0)
    ix.close()
