"""Generates tests/golden/*.npz by running the UNMODIFIED reference (oracle/_ref/ref_knn, compiled from
/root/reference/Code by oracle/Makefile) on small seeded inputs.  Run in the build container only:

    make -C oracle && python tests/golden/make_golden.py

Each fixture holds the inputs (data, queries, K, pct, srand seed, algorithm), the reference's outputs
(idx, dist2, dist, approximate-leaf node, work counters) and the reference's own flattened tree, so the
-m "not gpu" tests pin the oracle and the host tree builder, and the -m gpu tests pin the CUDA path, against
the reference itself without /root/reference being present.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle as orc  # noqa: E402


def clusters(rng, n, dim, k, sigma, scale=1.0):
    c = rng.random((k, dim))
    x = c[rng.integers(0, k, n)] + sigma * rng.standard_normal((n, dim))
    return (x * scale).astype(np.float32)


def cases():
    rng = np.random.default_rng(20181117)
    out = []
    d = rng.random((4000, 16), dtype=np.float32)
    out.append(("uniform16", d, rng.random((64, 16), dtype=np.float32), 10, 0.01))
    d = clusters(rng, 3000, 15, 16, 0.05, 1e5)
    out.append(("posture_like15", d, d[:50].copy(), 10, 0.005))
    d = rng.integers(0, 4, (2000, 8)).astype(np.float32)
    out.append(("integer_ties8", d, rng.integers(0, 4, (40, 8)).astype(np.float32), 5, 0.02))
    d = clusters(rng, 3000, 32, 30, 0.05)
    q = clusters(rng, 40, 32, 30, 0.05)
    out.append(("gauss32_k1", d, q, 1, 0.01))
    out.append(("gauss32_k100", d, q, 100, 0.01))
    d = np.clip(np.rint(clusters(rng, 1500, 128, 10, 0.1, 100.0)), 0, 218).astype(np.float32)
    out.append(("siftlike128", d, d[100:130] + rng.integers(0, 2, (30, 128)).astype(np.float32), 10, 0.02))
    d = rng.random((300, 3), dtype=np.float32)
    out.append(("tiny3", d, rng.random((20, 3), dtype=np.float32), 10, 0.1))
    return out


def main():
    assert orc.have_reference(), "build oracle/_ref first (make -C oracle)"
    for name, data, q, K, pct in cases():
        rec = {"data": data, "queries": q, "K": K, "pct": np.float32(pct), "seed": 1}
        for alg in (2, 0):
            r = orc.run_reference(data, q, K, pct, seed=1, alg=alg, want_tree=(alg == 2))
            for k in ("idx", "dist2", "dist", "approx_leaf_node", "counters"):
                rec["alg%d_%s" % (alg, k)] = r[k]
            if alg == 2:
                for k, v in r["tree"].items():
                    rec["tree_" + k] = np.asarray(v)
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **rec)
        print(name, data.shape, q.shape, "K", K, "leaves", int(rec["tree_n_leaves"]), "depth", int(rec["tree_depth"]),
              "%.0f KB" % (os.path.getsize(path) / 1024))


if __name__ == "__main__":
    main()
