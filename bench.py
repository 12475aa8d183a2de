#!/usr/bin/env python
"""Benchmark of the batched exact k-NN query path (BASELINE.json metric: exact k-NN queries/s, K=10).

Default workload = BASELINE.json configs[1]: synthetic uniform 1M points x 16-D, 100k queries, K=10, leaf <= 500
(pct 0.0005).  A "step" is one pass of the hot path over one batch of 100k queries.

    python bench.py [--gpus N] [--steps K] [--warmup W]            this repo's CUDA path (one process per GPU under torchrun)
    python bench.py --impl reference [...]                         the reference's CPU_Convex_Tree on the host cores

value  : queries/s with the queries already resident in HBM (scht_query_device), CUDA events, max over ranks
e2e    : queries/s through scht_query with pinned HOST buffers (H2D of the queries and D2H of idx+dist2 inside the timed region)
roofline: leaf-scan kernel: algorithmic bytes (points the REFERENCE algorithm scans x D x 4 B) / measured scan-kernel time
cpu_baseline: the reference (oracle/_ref/ref_knn, kind "reference") or the C oracle (kind "port") on a bounded query sample
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

WORKLOADS = {
    # name: (n, dim, nq, K, pct, kind)
    "uniform_1Mx16_q100k_K10": (1_000_000, 16, 100_000, 10, 0.0005, "uniform"),
    "gauss_10Mx32_q1M_K10": (10_000_000, 32, 1_000_000, 10, 0.00005, "gauss1000"),
    "gauss_2Mx32_q200k_K10": (2_000_000, 32, 200_000, 10, 0.00025, "gauss1000"),
    "gauss_2Mx32_q200k_K100": (2_000_000, 32, 200_000, 100, 0.00025, "gauss1000"),
    "gauss_2Mx32_q200k_K1": (2_000_000, 32, 200_000, 1, 0.00025, "gauss1000"),
    "siftlike_1Mx128_q100k_K10": (1_000_000, 128, 100_000, 10, 0.0005, "sift"),
    "posture_like_78095x15_self_K10": (78_095, 15, 78_095, 10, 0.005, "posture"),
    "smoke_50kx16_q4k_K10": (50_000, 16, 4_096, 10, 0.005, "uniform"),
    "uniform_100Mx16_q100k_K10": (100_000_000, 16, 100_000, 10, 0.000005, "uniform"),  # cfg5's data set on one GPU (13 GB device image)
}


def make_data(kind, n, dim, nq, seed):
    rng = np.random.default_rng(seed)
    if kind == "uniform":
        return rng.random((n, dim), dtype=np.float32), np.random.default_rng(seed + 1000).random((nq, dim), dtype=np.float32)

    def mix(r, m, k, sigma, scale=1.0):
        c = np.random.default_rng(seed + 7).random((k, dim))
        return ((c[r.integers(0, k, m)] + sigma * r.standard_normal((m, dim))) * scale).astype(np.float32)
    if kind == "gauss1000":
        return mix(rng, n, 1000, 0.05), mix(np.random.default_rng(seed + 1000), nq, 1000, 0.05)
    if kind == "sift":
        d = np.clip(np.rint(mix(rng, n, 200, 0.12, 100.0)), 0, 218).astype(np.float32)
        q = np.clip(np.rint(mix(np.random.default_rng(seed + 1000), nq, 200, 0.12, 100.0)), 0, 218).astype(np.float32)
        return d, q
    if kind == "posture":
        d = mix(rng, n, 64, 0.05, 1e5)
        return d, d[:nq].copy()
    raise ValueError(kind)


class ClockSampler:
    """nvidia-smi clocks + throttle reasons DURING the timed region (B200_PROFILING.md): one background
    `nvidia-smi -lms 100` process, parsed when the region ends."""
    FIELDS = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu):
        self.gpu, self.proc, self.path = gpu, None, None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=fd, stderr=subprocess.DEVNULL)
            os.close(fd)
            time.sleep(0.5)  # let the first samples land before the timed region starts
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()

    def summary(self):
        rows = []
        try:
            rows = [[x.strip() for x in ln.split(",")] for ln in open(self.path) if ln.count(",") >= 6]
            os.unlink(self.path)
        except Exception:
            pass
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        # "under load" = samples drawing more than half of the maximum power seen
        pw = [float(r[2]) if r[2].replace(".", "", 1).isdigit() else 0.0 for r in rows]
        load = [r for r, p in zip(rows, pw) if p >= 0.5 * max(pw)] or rows
        sm = sorted(int(r[0]) for r in load if r[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in load)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": int(rows[0][1]) if rows[0][1].isdigit() else None,
                "power_w_max": max(pw), "reasons": reasons, "samples": len(rows), "samples_under_load": len(load)}


def cpu_reference_sample(data, queries, K, pct, n_sample, cores):
    """Times the reference CPU path on `n_sample` queries using `cores` host processes (the reference is
    single-threaded: one process per core on disjoint query chunks).  Returns (q/s, kind, S_ref per query)."""
    import oracle as orc
    n_sample = min(n_sample, len(queries))
    qs = np.ascontiguousarray(queries[:n_sample])
    if orc.have_reference():
        from io_formats import read_records, write_fmat
        with tempfile.TemporaryDirectory() as td:
            dp, qp = os.path.join(td, "d.fmat"), os.path.join(td, "q.fmat")
            write_fmat(dp, data)
            write_fmat(qp, qs)
            cores = max(1, min(cores, n_sample))
            edges = np.linspace(0, n_sample, cores + 1).astype(int)
            procs = []
            for c in range(cores):
                if edges[c + 1] == edges[c]:
                    continue
                op = os.path.join(td, "o%d.knn" % c)
                cmd = [orc.REF_BIN, "knn", dp, repr(float(pct)), "1", qp, str(K), "2", str(edges[c]), str(edges[c + 1]), op]
                procs.append((subprocess.Popen(cmd, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL), op, edges[c + 1] - edges[c]))
            secs, dists, hyper = [], 0, 0
            for p, op, m in procs:
                if p.wait() != 0:
                    raise RuntimeError("ref_knn failed")
                r = read_records(op)
                secs.append(float(r["seconds"][1]))
                dists += int(r["counters"][0])
        return n_sample / max(secs), "reference", dists / n_sample, len(procs)
    # port: the C oracle with OpenMP over the same sample
    pkg = importlib.import_module("semi-convex_hull_tree_b200")
    ht = pkg.HostTree(data, pct, 1)
    t0 = time.perf_counter()
    _, _, _, _, cnt = orc.knn(ht, qs, K, alg=2, threads=cores)
    dt = time.perf_counter() - t0
    return n_sample / dt, "port", int(cnt[0]) / n_sample, cores


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="uniform_1Mx16_q100k_K10", choices=sorted(WORKLOADS))
    ap.add_argument("--cpu-sample", type=int, default=0, help="queries of the CPU baseline sample (0 = auto: ~20 s of CPU work)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    n, dim, nq, K, pct, kind = WORKLOADS[args.workload]
    cores = min(os.cpu_count() or 1, 64)
    config = {"workload": args.workload, "n_points": n, "dim": dim, "queries_per_step_per_gpu": nq, "K": K, "leaf_pts_percent": pct,
              "parallelism": "query-sharded x%d, tree replicated" % world, "l2": "point matrix %d MB; L2 not flushed between steps "
              "(each step re-reads every page %d times from L2/HBM by design)" % (n * dim * 4 >> 20, (nq + 55) // 56)}

    if args.impl == "reference":
        if rank != 0:
            return 0
        data, queries = make_data(kind, n, dim, nq, 1234)
        per_step = args.cpu_sample or max(cores, min(nq, cores * 24))
        vals = []
        for s in range(args.warmup + args.steps):
            off = (s * per_step) % max(1, nq - per_step + 1)
            v, knd, s_ref, used = cpu_reference_sample(data, queries[off:off + per_step], K, pct, per_step, cores)
            if s >= args.warmup:
                vals.append(v)
        v = float(np.mean(vals))
        line = {"impl": "reference", "metric": "exact k-NN queries/sec (K=%d)" % K, "value": v, "unit": "queries/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * per_step / v, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": v, "unit": "queries/s", "cores": used, "kind": knd,
                                 "sample": "%d queries per step, one single-threaded reference process per core" % per_step},
                "e2e": {"value": v, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
        print(json.dumps(line))
        return 0

    import torch
    import torch.distributed as dist
    pkg = importlib.import_module("semi-convex_hull_tree_b200")
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)

    data, queries = make_data(kind, n, dim, nq, 1234)
    if world > 1:  # weak scaling: every rank gets its own 100k queries, the tree is replicated
        _, queries = make_data(kind, 1, dim, nq, 1234 + 17 * rank) if kind == "uniform" else (None, queries[np.random.default_rng(rank).permutation(nq)])
    # the tree: built on this rank's GPU by scht_build_tree_device (bit-identical to the host builder, tests/test_gpu_build.py)
    t0 = time.perf_counter()
    ht = pkg.HostTree(data, pct, 1, device=local_rank)
    t_build = time.perf_counter() - t0
    ix = pkg.Index(ht, device=local_rank)
    info = ix.info()
    arr = ht.arrays()
    config.update({"n_leaves": int(arr["n_leaves"]), "tree_depth": int(arr["depth"]), "device_build_s": round(t_build, 2),
                   "device_tree_MB": info["device_bytes"] >> 20})

    # device-resident leg
    dq = torch.from_numpy(queries).to(dev)
    didx = torch.empty((nq, K), dtype=torch.int32, device=dev)
    dd2 = torch.empty((nq, K), dtype=torch.float32, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    st = pkg.Stats()

    def step_device(stats=None):
        ix.query_device(dq.data_ptr(), nq, dim, K, didx.data_ptr(), dd2.data_ptr(), None, stream, stats)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(args.warmup):
        step_device()
    sampler = ClockSampler(local_rank)
    sampler.start()
    sync_all()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step_device()
    e1.record()
    sync_all()
    ms = e0.elapsed_time(e1)
    # one extra instrumented step (not part of the timed loop) for the kernel-level numbers
    step_device(st)
    torch.cuda.synchronize()
    sampler.stop()
    stats = st.as_dict()

    # end-to-end leg: pinned host buffers through scht_query
    hq = torch.from_numpy(queries).pin_memory()
    hidx = torch.empty((nq, K), dtype=torch.int32).pin_memory()
    hd2 = torch.empty((nq, K), dtype=torch.float32).pin_memory()
    for _ in range(max(1, args.warmup)):
        ix.query_raw(hq.data_ptr(), nq, dim, K, hidx.data_ptr(), hd2.data_ptr())
    sync_all()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ix.query_raw(hq.data_ptr(), nq, dim, K, hidx.data_ptr(), hd2.data_ptr())
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    assert torch.equal(hidx, didx.cpu()), "host-buffer and device-buffer results differ"

    if world > 1:
        t = torch.tensor([ms, e2e_s * 1e3], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_s = float(t[0]), float(t[1]) / 1e3
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    value = world * nq * args.steps / (ms / 1e3)
    e2e = world * nq * args.steps / e2e_s
    line = {"metric": "exact k-NN queries/sec (K=%d)" % K, "value": value, "unit": "queries/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": config,
            "e2e": {"value": e2e, "unit": "queries/s", "h2d_bytes_per_step": nq * dim * 4, "d2h_bytes_per_step": nq * K * 8},
            "gpu_launches": int(stats["kernel_launches"]) * args.steps, "clocks": sampler.summary()}

    # CPU baseline + the reference algorithm's scan volume (algorithmic bytes of the roofline)
    s_ref = None
    if n > 10_000_000:
        args.no_cpu_baseline = True  # the reference's single-threaded host build alone would take hours at this size
    if not args.no_cpu_baseline:
        try:
            guess_qps_core = 30.0 * (1_000_000 * 16) / (n * dim)
            n_sample = args.cpu_sample or int(max(cores, min(nq, 20.0 * guess_qps_core * cores)))
            v, knd, dists_per_q, used = cpu_reference_sample(data, queries, K, pct, n_sample, cores)
            s_ref = dists_per_q
            line["cpu_baseline"] = {"value": v, "unit": "queries/s", "cores": used, "kind": knd,
                                    "sample": "first %d queries of the step, one single-threaded reference process per core" % n_sample}
        except Exception as ex:  # never lose the GPU line because the CPU leg failed
            line["cpu_baseline"] = {"value": None, "unit": "queries/s", "cores": cores, "kind": "port", "sample": "failed: %r" % (ex,)}
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    tpeak = float(peaks.get("bf16_tflops", 1590.0))
    s_gpu = stats["dist_computations"] / max(1, stats["n_queries"])
    if s_ref is None:
        s_ref = s_gpu
    used_tc = stats["prefilter_batches"] > 0
    # dominant kernel = the main leaf scan (k_scan_tc with the tensor-core prefilter, else k_scan); one launch per step
    main_s = stats["ms_main_scan"] / 1e3
    alg_bytes = s_ref * dim * 4 * nq  # points the REFERENCE algorithm scans x D x 4 B, for the queries of one step
    achieved = alg_bytes / main_s / 1e9
    clk = (line["clocks"].get("sm_mhz") or 1965) * 1e6
    pairs = stats["dist_computations"]
    traffic = None  # dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel, from the committed ncu --set full capture
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        ent = tj.get(args.workload, {}).get("k_scan_tc" if used_tc else "k_scan")
        if ent:
            traffic = ent["dram_bytes_per_launch"]
    except Exception:
        pass
    line["roofline"] = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                        "peak_source": "MEASURED_PEAKS.json hbm_gbs (burst)" if peaks else "fallback 6650 GB/s",
                        "kernel": "k_scan_tc (main scan, tensor-core prefilter + exact rescoring)" if used_tc else "k_scan (main scan)",
                        "kernel_ms": stats["ms_main_scan"], "launches_per_step": 1,
                        "algorithmic_bytes_per_launch": alg_bytes, "ref_scanned_points_per_query": s_ref, "gpu_scanned_points_per_query": s_gpu,
                        "note": "algorithmic bytes per SURVEY 8d (reference-scanned points x D x 4 B). Every staged page is reused by all "
                                "queries of a tile and re-read from L2 by the other tiles, so DRAM traffic is far below the algorithmic "
                                "bytes and frac >> 1: the kernel is not HBM-bound (see `tensor` / `fp32` below and DESIGN.md 5)",
                        "other_kernels_ms": {"seed_scan": stats["ms_seed_scan"], "traverse": stats["ms_traverse"],
                                             "rest": stats["ms_device"] - stats["ms_seed_scan"] - stats["ms_traverse"] - stats["ms_main_scan"]}}
    if used_tc:
        kp = (dim + 3 + 15) // 16 * 16
        line["roofline"]["tensor"] = {"achieved_tflops_algorithmic": pairs * dim * 2 / main_s / 1e12,
                                      "achieved_tflops_issued": pairs * kp * 2 / main_s / 1e12, "peak_tflops": tpeak,
                                      "frac_issued": pairs * kp * 2 / main_s / 1e12 / tpeak,
                                      "peak_source": "MEASURED_PEAKS.json bf16_tflops (burst; FP16 and BF16 MMA run at the same rate)",
                                      "note": "tensor pipe ~30 % busy (ncu): the epilogue (one TMEM read + one min per pair, ALU pipe ~60 % "
                                              "busy) and the MMA->epilogue hand-off latency bound the kernel, not the tensor pipe"}
    else:
        line["roofline"]["fp32"] = {"lane_ops_per_s": pairs * dim * 2 / main_s,
                                    "frac_of_128_lanes_per_clk_per_sm": pairs * dim * 2 / main_s / (148 * 128 * clk)}
    line["stats_one_step"] = stats
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
