import sys, importlib, time, os, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/oracle')
pkg = importlib.import_module("semi-convex_hull_tree_b200")
import oracle as orc
rng = np.random.default_rng(0)
n,dim,nq,K=200000,16,4000,10
data = rng.random((n,dim), dtype=np.float32); q = rng.random((nq,dim), dtype=np.float32)
ht = pkg.HostTree(data, 0.0025, seed=1)
oi,od2,od,_,cnt = orc.knn(ht, q, K, alg=2)
for tc in ("0","1","2"):
    os.environ["SCHT_TC"]=tc
    ix = pkg.Index(ht)
    idx,d2,st = ix.query(q,K,want_stats=True)
    print("tc",tc,"idx eq",np.array_equal(idx,oi),"d2 eq",np.array_equal(d2.view(np.uint32),od2.view(np.uint32)), {k:st[k] for k in ("dist_computations","exact_reevaluations","list_insertions","ms_device","ms_scan","kernel_launches","pages_listed","prefilter_batches","hyperplane_tests")}, flush=True)
    ix.close()
