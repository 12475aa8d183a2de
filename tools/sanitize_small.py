"""Small end-to-end run for compute-sanitizer: exact and prefilter paths (scan groups, page selection), brute force through
both paths, shard handles (two-phase calls) and the multi handle in both modes (one device)."""
import sys, importlib, os, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
pkg = importlib.import_module("semi-convex_hull_tree_b200")
import oracle as orc
import torch
rng = np.random.default_rng(0)
K = 10
def clusters(r, n, dim, k, sigma):
    c = np.random.default_rng(3).random((k, dim))
    return (c[r.integers(0, k, n)] + sigma * r.standard_normal((n, dim))).astype(np.float32)
for n, dim, nq in ((20000, 16, 600), (12000, 72, 400)):  # narrow points; wide points (row-major copy, 2-stage ring, N=256 chains)
    data = clusters(rng, n, dim, 12, 0.03); q = clusters(rng, nq, dim, 12, 0.03)
    ht = pkg.HostTree(data, 0.01, seed=1)
    oi, od2, _, _, _ = orc.knn(ht, q, K, alg=2)
    obi, obd2, _ = orc.brute_force(data, q[:100], K)
    os.environ["SCHT_GROUPS"] = "9"
    for tc in ("0", "1"):
        os.environ["SCHT_TC"] = tc
        ix = pkg.Index(ht)
        idx, d2, st = ix.query(q, K, want_stats=True)
        bi, bd2 = ix.brute_force(q[:100], K)
        print("dim", dim, "tc", tc, "ok", np.array_equal(idx, oi) and np.array_equal(d2.view(np.uint32), od2.view(np.uint32)), np.array_equal(bi, obi),
              "pages listed/all", st["pages_listed"], nq * max(1, ix.layout()["scan_pages"]), flush=True)
        ix.close()
    os.environ["SCHT_TC"] = "1"
    for mode in (pkg.SHARD_QUERIES, pkg.SHARD_DATASET):
        mx = pkg.Index(ht, devices=[0], mode=mode)
        idx, d2 = mx.query(q, K)
        bi, bd2 = mx.brute_force(q[:100], K)
        print("dim", dim, "multi mode", mode, "ok", np.array_equal(idx, oi) and np.array_equal(d2.view(np.uint32), od2.view(np.uint32)), np.array_equal(bi, obi), flush=True)
        mx.close()
    L = ht.arrays()["n_leaves"]
    parts = [(0, L // 2), (L // 2, L)]
    hs = [pkg.Index(ht, device=0, leaf_range=p) for p in parts]
    dq = torch.from_numpy(q).cuda()
    bounds = torch.empty((2, nq), dtype=torch.float32, device="cuda"); keys = torch.empty((2, nq, K), dtype=torch.int64, device="cuda")
    for s, h in enumerate(hs):
        h.partial_seed_device(dq.data_ptr(), nq, dim, K, bounds[s].data_ptr())
    bound = bounds.min(dim=0).values.contiguous()
    for s, h in enumerate(hs):
        h.partial_scan_device(dq.data_ptr(), nq, dim, K, bound.data_ptr(), keys[s].data_ptr())
    di = torch.empty((nq, K), dtype=torch.int32, device="cuda"); dd = torch.empty((nq, K), dtype=torch.float32, device="cuda")
    hs[0].merge_partial_device(keys.data_ptr(), 2, nq, dim, K, di.data_ptr(), dd.data_ptr(), None)
    print("dim", dim, "shards ok", np.array_equal(di.cpu().numpy(), oi) and np.array_equal(dd.cpu().numpy().view(np.uint32), od2.view(np.uint32)), flush=True)
    for h in hs:
        h.close()
