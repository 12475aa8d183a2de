"""Multi-GPU logic.  CPU: world_size-2 gloo run of the sharded-dataset exchange + merge, and the query partition.
GPU: the same mode emulated on one GPU (three leaf shards scanned one after the other, merged on the device)."""
import importlib
import os
import sys

import numpy as np
import pytest

from conftest import ROOT, assert_bit_equal


def sharding():
    return importlib.import_module("semi-convex_hull_tree_b200.sharding")


def ref_chain_d2(points, q):
    """the reference's distance arithmetic, vectorised over points (Code/basic_functions.h:167-180)"""
    acc = np.zeros(len(points), np.float32)
    for j in range(points.shape[1]):
        d = (points[:, j] - q[j]).astype(np.float32)
        acc = (acc.astype(np.float64) + d.astype(np.float64) * d.astype(np.float64)).astype(np.float32)
    return acc


def partial_keys_cpu(tree, queries, approx_leaf, K, leaf_lo, leaf_hi):
    """numpy emulation of scht_query_partial_device: K smallest keys among the leaves [leaf_lo, leaf_hi)"""
    sh = sharding()
    ls, lc = tree["leaf_start"], tree["leaf_count"]
    p0 = int(ls[leaf_lo]) if leaf_lo < len(ls) else int(tree["n_points"])
    p1 = int(ls[leaf_hi - 1] + lc[leaf_hi - 1]) if leaf_hi > leaf_lo else p0
    pts = tree["sorted_points"][p0:p1]
    out = np.full((len(queries), K), sh.EMPTY_KEY, np.uint64)
    for i, q in enumerate(queries):
        if p1 == p0:
            continue
        d2 = ref_chain_d2(pts, q)
        pos = np.arange(p0, p1, dtype=np.uint64)
        a0 = int(ls[approx_leaf[i]])
        a1 = a0 + int(lc[approx_leaf[i]])
        r = np.where((pos >= a0) & (pos < a1), 0, 1).astype(np.uint64)
        keys = (d2.view(np.uint32).astype(np.uint64) << np.uint64(32)) | (r << np.uint64(31)) | pos
        keys = np.sort(keys[d2 < np.float32(3.402823466e+38)])[:K]
        out[i, :len(keys)] = keys
    return out


def test_query_shard_partition():
    sh = sharding()
    for nq in (0, 1, 7, 100, 1001):
        for world in (1, 2, 3, 8):
            parts = [sh.query_shard(nq, world, r) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == nq
            assert all(parts[r][1] == parts[r + 1][0] for r in range(world - 1))
            sizes = [e - b for b, e in parts]
            assert max(sizes) - min(sizes) <= 1


def test_leaf_shards_balance(pkg):
    sh = sharding()
    rng = np.random.default_rng(0)
    cnt = rng.integers(1, 500, 1000)
    for world in (1, 2, 4, 8):
        parts = sh.leaf_shards(cnt, world)
        assert parts[0][0] == 0 and parts[-1][1] == 1000
        assert all(parts[r][1] == parts[r + 1][0] for r in range(world - 1))
        pts = [cnt[b:e].sum() for b, e in parts]
        assert max(pts) - min(pts) <= 2 * 500


def _gloo_worker(rank, world, port, tmpdir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    from conftest import load_golden
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sh = sharding()
    g = load_golden("uniform16")
    tree, q, K = g["tree"], g["queries"], g["K"]
    approx_leaf = tree["node_leaf_index"][g["alg2_approx_leaf_node"]]
    lb, le = sh.leaf_shards(tree["leaf_count"], world)[rank]
    mine = partial_keys_cpu(tree, q, approx_leaf, K, lb, le)
    t = torch.from_numpy(mine.view(np.int64))
    recv, (qb, qe) = sh.exchange_partial_keys(t, world, rank, sh.torch_all_to_all())
    merged = sh.merge_keys_host([r.numpy().view(np.uint64) for r in recv], K)
    idx, d2 = sh.keys_to_results(merged, tree["sorted_to_orig"])
    ok = np.array_equal(idx, g["alg2_idx"][qb:qe]) and np.array_equal(d2.view(np.uint32), g["alg2_dist2"][qb:qe].view(np.uint32))
    # query sharding needs no collective: only check that the slices tile the batch
    parts = [None] * world
    dist.all_gather_object(parts, (qb, qe))
    ok = ok and parts[0][0] == 0 and parts[-1][1] == len(q)
    open(os.path.join(tmpdir, "ok%d" % rank), "w").write("1" if ok else "0")
    dist.destroy_process_group()


def test_sharded_dataset_exchange_gloo_world2(pkg, tmp_path):
    """two CPU processes over gloo: each emulates one leaf shard, exchanges its partial K-lists and merges its query
    slice; the merged result must be the reference's (golden) result."""
    import torch.multiprocessing as mp
    port = 29600 + os.getpid() % 300
    mp.spawn(_gloo_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert [open(os.path.join(tmp_path, "ok%d" % r)).read() for r in range(2)] == ["1", "1"]


@pytest.mark.gpu
@pytest.mark.parametrize("shards", [1, 3])
def test_sharded_dataset_on_one_gpu(pkg, orc, shards):
    """scht_query_partial_device over `shards` leaf ranges + scht_merge_partial_device == scht_query == oracle"""
    import torch
    sh = sharding()
    rng = np.random.default_rng(9)
    n, dim, nq, K = 40000, 16, 1500, 10
    data, q = rng.random((n, dim), dtype=np.float32), rng.random((nq, dim), dtype=np.float32)
    ht = pkg.HostTree(data, 0.005, seed=1)
    oi, od2, od, _, _ = orc.knn(ht, q, K, alg=2)
    ix = pkg.Index(ht)
    dq = torch.from_numpy(q).cuda()
    keys = torch.empty((shards, nq, K), dtype=torch.int64, device="cuda")
    for s, (lb, le) in enumerate(sh.leaf_shards(ht.arrays()["leaf_count"], shards)):
        ix.query_partial_device(dq.data_ptr(), nq, dim, K, lb, le, keys[s].data_ptr())
    # the device lists against the numpy emulation of the contract: a shard's list holds keys of that shard only, in order,
    # and EVERY key of the shard that can belong to the final answer (d2 <= the query's final K-th distance); keys beyond
    # the seed bound may be missing (the scan never evaluates them)
    a = ht.arrays()
    approx = ix.find_approximate_leaves(q)
    lb, le = sh.leaf_shards(a["leaf_count"], shards)[0]
    want0 = partial_keys_cpu(a, q[:64], approx[:64], K, lb, le)
    got0 = keys[0, :64].cpu().numpy().view(np.uint64)
    for i in range(64):
        g = got0[i][got0[i] != sh.EMPTY_KEY]
        assert np.array_equal(g, want0[i][:len(g)]), "shard list of query %d is not a prefix of the shard's sorted keys" % i
        final_kth = np.uint64(od2[i, K - 1].view(np.uint32)) << np.uint64(32) | np.uint64(0xFFFFFFFF)
        assert len(g) >= int((want0[i] <= final_kth).sum())
    didx = torch.empty((nq, K), dtype=torch.int32, device="cuda")
    dd2 = torch.empty((nq, K), dtype=torch.float32, device="cuda")
    dd = torch.empty((nq, K), dtype=torch.float32, device="cuda")
    ix.merge_partial_device(keys.data_ptr(), shards, nq, dim, K, didx.data_ptr(), dd2.data_ptr(), dd.data_ptr())
    assert_bit_equal(didx.cpu().numpy(), oi, "idx")
    assert_bit_equal(dd2.cpu().numpy(), od2, "dist2")
    assert_bit_equal(dd.cpu().numpy(), od, "dist")
    ix.close()
