"""Sharded-dataset mode on real GPUs (optional mode): torchrun --nproc-per-node N tools/sharded_dataset_demo.py

Every rank holds the tree but scans only its contiguous LEAF shard for ALL queries (scht_query_partial_device), the
per-shard sorted key lists are exchanged with ONE NCCL all-to-all so that rank g owns the lists of its query slice, and
scht_merge_partial_device merges them exactly.  Rank 0 also answers its slice with the ordinary replicated-tree query and
checks that both give the same bits.  Prints one JSON line with the timings (CUDA events, max over ranks)."""
import importlib, json, os, sys, time
import numpy as np
import torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("semi-convex_hull_tree_b200")
sh = importlib.import_module("semi-convex_hull_tree_b200.sharding")

def main():
    rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    dist.init_process_group("nccl", device_id=dev)
    n, dim, nq, K = 1_000_000, 16, 65_536, 10
    rng = np.random.default_rng(5)
    data, q = rng.random((n, dim), dtype=np.float32), np.random.default_rng(6).random((nq, dim), dtype=np.float32)
    ht = pkg.HostTree(data, 0.0005, 1)
    ix = pkg.Index(ht, device=lr)
    lb, le = sh.leaf_shards(ht.arrays()["leaf_count"], world)[rank]
    dq = torch.from_numpy(q).to(dev)
    keys = torch.empty((nq, K), dtype=torch.int64, device=dev)
    qb, qe = sh.query_shard(nq, world, rank)
    idx = torch.empty((qe - qb, K), dtype=torch.int32, device=dev)
    d2 = torch.empty((qe - qb, K), dtype=torch.float32, device=dev)
    stream = torch.cuda.current_stream().cuda_stream

    def step():
        ix.query_partial_device(dq.data_ptr(), nq, dim, K, lb, le, keys.data_ptr(), stream)
        recv, _ = sh.exchange_partial_keys(keys, world, rank, sh.torch_all_to_all())
        stacked = torch.stack(recv).contiguous()  # [world][nq_mine][K]
        ix.merge_partial_device(stacked.data_ptr(), world, qe - qb, dim, K, idx.data_ptr(), d2.data_ptr(), None, stream)

    step()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        step()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / 3], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    # reference: the replicated-tree query of this rank's slice
    ridx = torch.empty_like(idx); rd2 = torch.empty_like(d2)
    ix.query_device(dq[qb:qe].contiguous().data_ptr(), qe - qb, dim, K, ridx.data_ptr(), rd2.data_ptr(), None, stream)
    torch.cuda.synchronize()
    ok = torch.tensor([int(torch.equal(ridx, idx) and torch.equal(rd2.view(torch.int32), d2.view(torch.int32)))], device=dev)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(json.dumps({"mode": "sharded-dataset", "n_gpus": world, "n_points": n, "dim": dim, "queries": nq, "K": K,
                          "ms_per_pass": float(t[0]), "queries_per_s": nq / (float(t[0]) / 1e3),
                          "all_to_all_bytes_per_rank": nq * K * 8 * (world - 1) // world,
                          "bit_identical_to_replicated_query": bool(ok.item())}))
    dist.destroy_process_group()

if __name__ == "__main__":
    main()
