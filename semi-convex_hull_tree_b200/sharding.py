"""Multi-GPU plumbing of the query path.  One process per GPU; torch.distributed is only the transport.

Default mode — QUERY SHARDING: queries are independent (the reference kernel is thread-per-query with no cross-query
state, Code/kernels/do_kNN_gpu.cl:322-326), so every rank holds the whole tree and answers a contiguous slice of the
queries.  No data-path collective.

Optional mode — SHARDED DATASET: every rank scans only a contiguous range of LEAVES (a contiguous row range of the
leaf-sorted point matrix) for all queries and emits, per query, K sorted 64-bit keys
(float_bits(d2) << 32 | not_approx_leaf << 31 | sorted_position).  One all-to-all moves rank r's partial lists for the
query slice of rank g to rank g, which merges the world_size lists exactly (keys are global, so the merge of the
per-shard K-smallest sets is the global K-smallest set).
"""
import numpy as np

EMPTY_KEY = np.uint64(0x7F7FFFFFFFFFFFFF)


def query_shard(n_queries, world, rank):
    """contiguous slice [begin, end) of the queries answered by `rank`"""
    base, rem = divmod(int(n_queries), int(world))
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def leaf_shards(leaf_count, world):
    """split leaves [0, L) into `world` contiguous ranges holding about the same number of points"""
    leaf_count = np.asarray(leaf_count, dtype=np.int64)
    L = len(leaf_count)
    csum = np.concatenate([[0], np.cumsum(leaf_count)])
    total = csum[-1]
    cuts = [0]
    for r in range(1, world):
        target = total * r / world
        c = int(np.searchsorted(csum, target, side="left"))
        cuts.append(min(max(c, cuts[-1]), L))
    cuts.append(L)
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def merge_keys_host(key_lists, K):
    """exact merge of per-shard key lists [n_lists][nq][K] (numpy, for the CPU tests of the exchange)"""
    allk = np.concatenate([np.asarray(k, dtype=np.uint64) for k in key_lists], axis=1)
    allk.sort(axis=1)
    return allk[:, :K]


def keys_to_results(keys, sorted_to_orig):
    """keys [nq][K] -> (idx int32, dist2 float32) with the library's conventions for empty slots"""
    keys = np.asarray(keys, dtype=np.uint64)
    d2 = (keys >> np.uint64(32)).astype(np.uint32).view(np.float32)
    pos = (keys & np.uint64(0x7FFFFFFF)).astype(np.int64)
    empty = keys == EMPTY_KEY
    idx = np.where(empty, -1, np.asarray(sorted_to_orig)[np.where(empty, 0, pos)]).astype(np.int32)
    return idx, d2


def exchange_partial_keys(my_keys, world, rank, all_to_all):
    """my_keys: [nq][K] partial lists of THIS rank's leaf shard for ALL queries.
    Returns ([world][nq_mine][K] lists for this rank's query slice, (begin, end)).
    `all_to_all(send_chunks) -> recv_chunks` is the transport (torch.distributed.all_to_all on NCCL or gloo)."""
    nq = my_keys.shape[0]
    send = []
    for g in range(world):
        b, e = query_shard(nq, world, g)
        send.append(my_keys[b:e])
    recv = all_to_all(send)
    return recv, query_shard(nq, world, rank)


def torch_all_to_all(group=None):
    """transport for exchange_partial_keys over torch.distributed (int64 views of the uint64 keys)"""
    import torch
    import torch.distributed as dist

    def fn(send):
        world = dist.get_world_size(group)
        rank = dist.get_rank(group)
        send = [s.contiguous() for s in send]
        recv = [torch.empty_like(send[rank]) for _ in range(world)]
        if dist.get_backend(group) == "gloo":
            # gloo has no all_to_all: pairwise isend/irecv
            reqs = []
            for g in range(world):
                if g == rank:
                    recv[g].copy_(send[g])
                else:
                    reqs.append(dist.isend(send[g], g, group=group))
                    reqs.append(dist.irecv(recv[g], g, group=group))
            for r in reqs:
                r.wait()
        else:
            dist.all_to_all(recv, send, group=group)
        return recv
    return fn
