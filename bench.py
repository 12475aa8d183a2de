#!/usr/bin/env python
"""Benchmark of the batched exact k-NN query path (BASELINE.json metric: exact k-NN queries/s, K=10).

Default workload = BASELINE.json configs[1]: synthetic uniform 1M points x 16-D, 100k queries, K=10, leaf <= 500
(pct 0.0005).  A "step" is one pass of the hot path over one batch of 100k queries.

    python bench.py [--gpus N] [--steps K] [--warmup W]            this repo's CUDA path (one process per GPU under torchrun)
    python bench.py --impl reference [...]                         the reference's CPU_Convex_Tree on the host cores

value  : queries/s with the queries already resident in HBM (scht_query_device), CUDA events, max over ranks
e2e    : queries/s through scht_query with pinned HOST buffers (H2D of the queries and D2H of idx+dist2 inside the timed region)
roofline: the dominant kernel (main leaf scan) against the roof that binds it: the tensor pipe when the tcgen05 prefilter
         scan ran (issued MMA flops / measured bf16 peak), the FP32 pipe for the exact scan; the SURVEY 8d HBM form
         (points the REFERENCE algorithm scans x D x 4 B / kernel time) is kept next to it as `hbm_algorithmic`
cpu_baseline: the reference (oracle/_ref/ref_knn, kind "reference") or the C oracle (kind "port") on a bounded query sample;
         the GPU's answers for the same queries are compared with it bit for bit (parity_checked / parity_mismatches)
With WORLD_SIZE > 1 the line also carries `sharded_dataset` (one leaf shard per rank, NCCL all-reduce of the seed bounds +
all-to-all of the K-lists, merged per query slice) and `single_process` (rank 0 alone drives all GPUs through
scht_create_multi: strong scaling of one call, query-sharded and dataset-sharded).
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
    "gauss_10Mx32_q1M_K1": (10_000_000, 32, 1_000_000, 1, 0.00005, "gauss1000"),
    "gauss_10Mx32_q1M_K100": (10_000_000, 32, 1_000_000, 100, 0.00005, "gauss1000"),
    "gauss_2Mx32_q200k_K10": (2_000_000, 32, 200_000, 10, 0.00025, "gauss1000"),
    "gauss_2Mx32_q200k_K100": (2_000_000, 32, 200_000, 100, 0.00025, "gauss1000"),
    "gauss_2Mx32_q200k_K1": (2_000_000, 32, 200_000, 1, 0.00025, "gauss1000"),
    "siftlike_1Mx128_q100k_K10": (1_000_000, 128, 100_000, 10, 0.0005, "sift"),
    "posture_like_78095x15_self_K10": (78_095, 15, 78_095, 10, 0.005, "posture"),
    "smoke_50kx16_q4k_K10": (50_000, 16, 4_096, 10, 0.005, "uniform"),
    "uniform_100Mx16_q100k_K10": (100_000_000, 16, 100_000, 10, 0.000005, "uniform"),  # cfg5's data set on one GPU (13 GB device image)
    "uniform_100Mx16_q10M_K10": (100_000_000, 16, 10_000_000, 10, 0.000005, "uniform"),  # cfg5 (use --single-process with --gpus N)
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


def cpu_reference_sample(data, queries, K, pct, n_sample, cores, host_tree=None):
    """Times the reference CPU path on the first `n_sample` queries using `cores` host processes (the reference is
    single-threaded: one process per core on disjoint query chunks).  With `host_tree` (big data sets, where
    every reference process would first spend minutes building the same tree) the C oracle port runs instead, OpenMP over
    the sample on the given tree.  Returns dict(qps, kind, scanned_per_query, cores, idx, dist2)."""
    import oracle as orc
    n_sample = min(n_sample, len(queries))
    qs = np.ascontiguousarray(queries[:n_sample])
    if host_tree is None and orc.have_reference():
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
            secs, dists, idx, d2 = [], 0, [], []
            for p, op, m in procs:
                if p.wait() != 0:
                    raise RuntimeError("ref_knn failed")
                r = read_records(op)
                secs.append(float(r["seconds"][1]))
                dists += int(r["counters"][0])
                idx.append(r["idx"].reshape(m, K))
                d2.append(r["dist2"].reshape(m, K))
        return dict(qps=n_sample / max(secs), kind="reference", scanned_per_query=dists / n_sample, cores=len(procs),
                    idx=np.concatenate(idx), dist2=np.concatenate(d2))
    # port: the C oracle with OpenMP over the same sample
    pkg = importlib.import_module("semi-convex_hull_tree_b200")
    ht = host_tree if host_tree is not None else pkg.HostTree(data, pct, 1)
    t0 = time.perf_counter()
    idx, d2, _, _, cnt = orc.knn(ht, qs, K, alg=2, threads=cores)
    dt = time.perf_counter() - t0
    return dict(qps=n_sample / dt, kind="port", scanned_per_query=int(cnt[0]) / n_sample, cores=cores, idx=idx, dist2=d2)


def dropin_e2e(data, queries, K, pct, reps=4):
    """Convex_Tree<float>::kNN() of the drop-in subclass, timed inside the C++ harness (last of `reps` calls): pageable host
    memory in and out, one GPU.  Also checks the harness's answers against nothing here -- tests/test_host_cpp.py does that."""
    import re
    import oracle as orc
    from io_formats import write_fmat
    binary = os.path.join(os.path.dirname(orc.REF_BIN), "ref_b200_knn")
    if not os.access(binary, os.X_OK):
        raise RuntimeError("oracle/_ref/ref_b200_knn not built")
    with tempfile.TemporaryDirectory() as td:
        dp, qp, op = os.path.join(td, "d.fmat"), os.path.join(td, "q.fmat"), os.path.join(td, "o.knn")
        write_fmat(dp, data)
        write_fmat(qp, queries)
        r = subprocess.run([binary, dp, repr(float(pct)), "1", qp, str(K), "0", op, "0", "0", str(reps)], capture_output=True, text=True, timeout=600)
        if r.returncode != 0:
            raise RuntimeError("ref_b200_knn failed: " + r.stderr[-300:])
        m = re.search(r"call_ms=([0-9.]+)", r.stderr)
        ms = float(m.group(1))
    return {"value": len(queries) / (ms / 1e3), "unit": "queries/s", "ms_per_call": ms,
            "path": "Convex_Tree<float>::kNN() on B200_Convex_Tree (reference host tree, pageable Matrix<T> in, std::vector out), call %d of %d" % (reps, reps)}


def parity(gpu_idx, gpu_d2, cpu):
    """bit-for-bit comparison of the GPU's answers with the CPU leg's on the same queries"""
    m = len(cpu["idx"])
    bad = int((gpu_idx[:m] != cpu["idx"]).any(axis=1).sum() + (gpu_d2[:m].view(np.uint32) != cpu["dist2"].view(np.uint32)).any(axis=1).sum())
    return m, bad


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="uniform_1Mx16_q100k_K10", choices=sorted(WORKLOADS))
    ap.add_argument("--cpu-sample", type=int, default=0, help="queries of the CPU baseline sample (0 = auto: ~20 s of CPU work)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--single-process", action="store_true",
                    help="one process drives --gpus N devices through scht_create_multi (no torchrun): strong scaling of one call")
    ap.add_argument("--mode", default="both", choices=["queries", "dataset", "both"], help="--single-process: sharding mode(s) to time")
    ap.add_argument("--no-multi-extras", action="store_true", help="WORLD_SIZE > 1: skip the sharded_dataset / single_process records")
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
            r = cpu_reference_sample(data, queries[off:off + per_step], K, pct, per_step, cores)
            if s >= args.warmup:
                vals.append(r["qps"])
        v = float(np.mean(vals))
        line = {"impl": "reference", "metric": "exact k-NN queries/sec (K=%d)" % K, "value": v, "unit": "queries/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * per_step / v, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": v, "unit": "queries/s", "cores": r["cores"], "kind": r["kind"], "queries_per_s_per_core": v / r["cores"],
                                 "sample": "%d queries per step, one single-threaded reference process per core" % per_step},
                "e2e": {"value": v, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
        print(json.dumps(line))
        return 0

    if args.single_process:
        return single_process_main(args, config)

    import torch
    import torch.distributed as dist
    pkg = importlib.import_module("semi-convex_hull_tree_b200")
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)

    data, queries0 = make_data(kind, n, dim, nq, 1234)
    queries = queries0
    if world > 1:  # weak scaling: every rank gets its own 100k queries, the tree is replicated
        _, queries = make_data(kind, 1, dim, nq, 1234 + 17 * rank) if kind == "uniform" else (None, queries0[np.random.default_rng(rank).permutation(nq)])
    # the tree: built on this rank's GPU by scht_build_tree_device (bit-identical to the host builder, tests/test_gpu_build.py)
    t0 = time.perf_counter()
    ht = pkg.HostTree(data, pct, 1, device=local_rank)
    t_build = time.perf_counter() - t0
    t0 = time.perf_counter()
    ix = pkg.Index(ht, device=local_rank)
    t_create = time.perf_counter() - t0
    info = ix.info()
    lay = ix.layout()
    arr = ht.arrays()
    tree_info = {"n_leaves": int(arr["n_leaves"]), "tree_depth": int(arr["depth"]), "device_build_s": round(t_build, 2),
                 "device_create_s": round(t_create, 2), "device_tree_MB": info["device_bytes"] >> 20, "scan_groups": lay["scan_groups"],
                 "super_groups": lay["super_groups"], "scan_pages": lay["scan_pages"]}

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

    extras = {}
    if world > 1 and not args.no_multi_extras:
        try:
            extras = multi_gpu_extras(pkg, dist, torch, ht, ix, queries0, dim, K, rank, local_rank, world, dev, min(args.steps, 3))
        except Exception as ex:  # never lose the headline line because an extra failed
            extras = {"multi_gpu_extras_error": repr(ex)}
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    value = world * nq * args.steps / (ms / 1e3)
    e2e = world * nq * args.steps / e2e_s
    line = {"metric": "exact k-NN queries/sec (K=%d)" % K, "value": value, "unit": "queries/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": config, "tree": tree_info,
            "e2e": {"value": e2e, "unit": "queries/s", "h2d_bytes_per_step": nq * dim * 4, "d2h_bytes_per_step": nq * K * 8},
            "gpu_launches": int(stats["kernel_launches"]) * args.steps, "clocks": sampler.summary()}
    line.update(extras)

    # CPU baseline (+ parity of the GPU's answers on the same queries) and the reference algorithm's scan volume
    s_ref = None
    exit_code = 0
    if not args.no_cpu_baseline:
        try:
            big = n * dim > 64_000_000  # every reference process would rebuild the tree for minutes: run the C oracle port on the device-built tree
            guess_qps_core = 30.0 * (1_000_000 * 16) / (n * dim)
            n_sample = args.cpu_sample or int(max(cores, min(nq, (6.0 if big else 20.0) * guess_qps_core * cores)))
            r = cpu_reference_sample(data, queries, K, pct, n_sample, cores, host_tree=ht if big else None)
            s_ref = r["scanned_per_query"]
            line["cpu_baseline"] = {"value": r["qps"], "unit": "queries/s", "cores": r["cores"], "kind": r["kind"],
                                    "queries_per_s_per_core": r["qps"] / r["cores"],
                                    "sample": "first %d queries of the step, %s" % (n_sample, "one single-threaded reference process per core"
                                                                                    if r["kind"] == "reference" else "C oracle port, OpenMP over the sample")}
            checked, bad = parity(hidx.numpy(), hd2.numpy(), r)
            line["parity_checked"], line["parity_mismatches"] = checked, bad
            if bad:
                exit_code = 1
        except Exception as ex:  # never lose the GPU line because the CPU leg failed
            line["cpu_baseline"] = {"value": None, "unit": "queries/s", "cores": cores, "kind": "port", "sample": "failed: %r" % (ex,)}
    # end to end through the C++ drop-in (integration/B200_Convex_Tree.h inside the reference's own headers, oracle/_ref/ref_b200_knn):
    # the reference's constructor builds the tree on the host, kNN() takes the reference's pageable Matrix<T> and fills std::vectors
    if world == 1 and not args.no_cpu_baseline and n * dim <= 32_000_000:
        try:
            line["e2e_dropin"] = dropin_e2e(data, queries, K, pct)
        except Exception as ex:
            line["e2e_dropin"] = {"skipped": repr(ex)}
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    tpeak = float(peaks.get("bf16_tflops", 1590.0))
    s_gpu = stats["dist_computations"] / max(1, stats["n_queries"])
    used_tc = stats["prefilter_batches"] > 0
    # dominant kernel = the main leaf scan (k_scan_tc with the tensor-core prefilter, else k_scan); one launch per step
    main_s = stats["ms_main_scan"] / 1e3
    clk = (line["clocks"].get("sm_mhz") or 1965) * 1e6
    pairs = stats["dist_computations"]
    prof = {}  # numbers of the committed ncu --set full capture of this workload's dominant kernel (profiles/traffic.json)
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        prof = tj.get(args.workload, {}).get("k_scan_tc" if used_tc else "k_scan") or {}
    except Exception:
        pass
    kp = (dim + 3 + 15) // 16 * 16
    if used_tc:
        ach = pairs * kp * 2 / main_s / 1e12
        roof = {"bound": "tensor", "achieved": ach, "peak": tpeak, "unit": "TFLOP/s", "frac": ach / tpeak,
                "peak_source": "MEASURED_PEAKS.json bf16_tflops (burst; FP16 and BF16 MMA run at the same rate)" if peaks else "fallback 1590 TFLOP/s",
                "kernel": "k_scan_tc (main scan: tcgen05 FP16 prefilter + exact rescoring)",
                "flops": "issued: (query, point) pairs the kernel put through the MMA x K'=%d x 2; algorithmic 2*D flops = %.1f TFLOP/s" % (
                    kp, pairs * dim * 2 / main_s / 1e12)}
    else:
        ach = pairs * dim * 2 / main_s / 1e12
        fpeak = 148 * 128 * 2 * clk / 1e12
        roof = {"bound": "fp32", "achieved": ach, "peak": fpeak, "unit": "TFLOP/s", "frac": ach / fpeak,
                "peak_source": "148 SMs x 128 FP32 lanes x 2 flop x measured SM clock (FADD2+FFMA2 per dimension: 2 lane-ops per pair and dimension)",
                "kernel": "k_scan (main scan, exact FP32)"}
    roof.update({"traffic": prof.get("dram_bytes_per_launch"), "kernel_ms": stats["ms_main_scan"], "launches_per_step": 1,
                 "gpu_scanned_points_per_query": s_gpu, "ref_scanned_points_per_query": s_ref,
                 "ncu": {k: prof[k] for k in ("tensor_pipe_pct", "alu_pipe_pct", "l2_hit_pct", "capture") if k in prof} or None,
                 "other_kernels_ms": {"seed_scan": stats["ms_seed_scan"], "select_or_traverse": stats["ms_traverse"],
                                      "rest": stats["ms_device"] - stats["ms_seed_scan"] - stats["ms_traverse"] - stats["ms_main_scan"]}})
    if s_ref is not None:
        alg_bytes = s_ref * dim * 4 * nq  # points the REFERENCE algorithm scans x D x 4 B, for the queries of one step
        roof["hbm_algorithmic"] = {"achieved": alg_bytes / main_s / 1e9, "peak": peak, "unit": "GB/s", "frac": alg_bytes / main_s / 1e9 / peak,
                                   "algorithmic_bytes_per_launch": alg_bytes, "measured_dram_bytes_per_launch": prof.get("dram_bytes_per_launch"),
                                   "note": "SURVEY 8d form. Every staged page is shared by the 128 queries of a tile and the images stay in L2, "
                                           "so DRAM moves orders of magnitude fewer bytes and this fraction exceeds 1: HBM does not bind the kernel"}
    line["roofline"] = roof
    line["stats_one_step"] = stats
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return exit_code


def multi_gpu_extras(pkg, dist, torch, ht, ix, queries0, dim, K, rank, local_rank, world, dev, steps):
    """WORLD_SIZE > 1 only.  (1) sharded-dataset mode, one process per GPU: rank r holds the leaves of shard r
    (scht_create_shard), every rank scans its shard for ALL queries; the seed bounds are all-reduced (MIN) and the K-lists
    exchanged with ONE all-to-all, both NCCL over NVLink; rank r merges its query slice.  (2) rank 0 alone drives all GPUs
    through scht_create_multi (the other ranks wait on the host): strong scaling of one call."""
    sh = importlib.import_module("semi-convex_hull_tree_b200.sharding")
    out = {}
    nq = len(queries0)
    stream = torch.cuda.current_stream().cuda_stream
    parts = sh.leaf_shards(ht.arrays()["leaf_count"], world)
    shard = pkg.Index(ht, device=local_rank, leaf_range=parts[rank])
    ids = [pkg.Comm.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    comm = pkg.Comm(ids[0], rank, world, local_rank)  # the library's own NCCL communicator (scht_comm_create)
    dq = torch.from_numpy(queries0).to(dev)
    qb, qe = comm.query_slice(nq)
    idx = torch.empty((qe - qb, K), dtype=torch.int32, device=dev)
    d2 = torch.empty((qe - qb, K), dtype=torch.float32, device=dev)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]

    def step():  # seed scan -> ncclAllReduce(min) of the bounds -> shard scan -> grouped ncclSend/ncclRecv of the K-lists -> merge
        comm.query_sharded_device(shard, dq.data_ptr(), nq, dim, K, idx.data_ptr(), d2.data_ptr(), None, stream)

    step()
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    ev[0].record()
    for _ in range(steps):
        step()
    ev[1].record()
    torch.cuda.synchronize()
    # the exchange alone (same payload through torch's all_to_all): what the NVLink fabric delivers for this message size
    keys = torch.empty((nq, K), dtype=torch.int64, device=dev)
    a2a = sh.torch_all_to_all()
    sh.exchange_partial_keys(keys, world, rank, a2a)
    torch.cuda.synchronize()
    ev[2].record()
    sh.exchange_partial_keys(keys, world, rank, a2a)
    ev[3].record()
    torch.cuda.synchronize()
    t = torch.tensor([ev[0].elapsed_time(ev[1]) / steps, ev[2].elapsed_time(ev[3])], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    # the same query slice through the replicated tree of this rank
    ridx, rd2 = torch.empty_like(idx), torch.empty_like(d2)
    ix.query_device(dq[qb:qe].contiguous().data_ptr(), qe - qb, dim, K, ridx.data_ptr(), rd2.data_ptr(), None, stream)
    torch.cuda.synchronize()
    ok = torch.tensor([int(torch.equal(ridx, idx) and torch.equal(rd2.view(torch.int32), d2.view(torch.int32)))], device=dev)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    a2a_bytes = nq * K * 8 * (world - 1) // world
    out["sharded_dataset"] = {"mode": "one process per GPU, leaf shards; inside the library: ncclAllReduce(min) of the seed bounds, grouped "
                                      "ncclSend/ncclRecv of the K-lists, merge per query slice (scht_query_sharded_device)", "n_gpus": world,
                              "queries": nq, "ms_per_step": float(t[0]), "queries_per_s": nq / (float(t[0]) / 1e3),
                              "all_to_all_bytes_per_rank": a2a_bytes, "all_to_all_ms": float(t[1]),
                              "all_to_all_GBps_per_rank": a2a_bytes / max(float(t[1]), 1e-6) / 1e6,
                              "shard_points": int(ht.arrays()["leaf_count"][parts[rank][0]:parts[rank][1]].sum()),
                              "shard_device_MB": shard.info()["device_bytes"] >> 20,
                              "bit_identical_to_replicated_query": bool(ok.item())}
    comm.close()
    shard.close()
    del dq, keys
    torch.cuda.synchronize()

    # (2) single process, all GPUs: the other ranks wait on the host (a NCCL barrier would spin on their GPUs)
    store = dist.distributed_c10d._get_default_store()
    if rank == 0:
        try:
            if pkg.lib().scht_device_count() >= world:
                big_q = np.concatenate([queries0] + [queries0[np.random.default_rng(g).permutation(nq)] for g in range(1, world)])
                out["single_process"], _, _ = single_process_runs(pkg, torch, ht, big_q, dim, K, list(range(world)), steps, ("queries", "dataset"), single=ix)
            else:
                out["single_process"] = {"skipped": "only %d device(s) visible to rank 0" % pkg.lib().scht_device_count()}
        finally:
            store.set("scht_single_process_done", "1")
    else:
        store.wait(["scht_single_process_done"])
    return out


def single_process_runs(pkg, torch, ht, queries, dim, K, devices, steps, modes, single=None, warm=1):
    """one host process, scht_create_multi over `devices`: end-to-end q/s of one scht_query call with pinned host buffers.
    Returns (record, idx, dist2 of the last call)."""
    nq = len(queries)
    hq = torch.from_numpy(queries).pin_memory()
    res = {"n_gpus": len(devices), "queries_per_call": nq, "timing": "wall clock around scht_query (pinned host buffers in and out)"}
    ref_idx = ref_d2 = None
    one_ms = None
    if single is not None:  # one GPU, same call: the strong-scaling denominator
        hi, hd = torch.empty((nq, K), dtype=torch.int32).pin_memory(), torch.empty((nq, K), dtype=torch.float32).pin_memory()
        for _ in range(warm):
            single.query_raw(hq.data_ptr(), nq, dim, K, hi.data_ptr(), hd.data_ptr())
        t0 = time.perf_counter()
        for _ in range(steps):
            single.query_raw(hq.data_ptr(), nq, dim, K, hi.data_ptr(), hd.data_ptr())
        dt = (time.perf_counter() - t0) / steps
        one_ms = dt * 1e3
        res["one_gpu"] = {"ms_per_call": one_ms, "queries_per_s": nq / dt}
        ref_idx, ref_d2 = hi.clone(), hd.clone()
    hi = hd = None
    for mode in modes:
        t0 = time.perf_counter()
        mx = pkg.Index(ht, devices=devices, mode=pkg.SHARD_QUERIES if mode == "queries" else pkg.SHARD_DATASET)
        t_create = time.perf_counter() - t0
        hi, hd = torch.empty((nq, K), dtype=torch.int32).pin_memory(), torch.empty((nq, K), dtype=torch.float32).pin_memory()
        for _ in range(warm):
            mx.query_raw(hq.data_ptr(), nq, dim, K, hi.data_ptr(), hd.data_ptr())
        t0 = time.perf_counter()
        for _ in range(steps):
            mx.query_raw(hq.data_ptr(), nq, dim, K, hi.data_ptr(), hd.data_ptr())
        dt = (time.perf_counter() - t0) / steps
        ent = {"ms_per_call": dt * 1e3, "queries_per_s": nq / dt, "create_s": round(t_create, 2), "device_MB_total": mx.info()["device_bytes"] >> 20}
        if ref_idx is not None:
            ent["bit_identical_to_" + ("one_gpu" if one_ms is not None else "the_other_mode")] = bool(
                torch.equal(hi, ref_idx) and torch.equal(hd.view(torch.int32), ref_d2.view(torch.int32)))
        else:
            ref_idx, ref_d2 = hi.clone(), hd.clone()
        if one_ms is not None:
            ent["speedup_vs_one_gpu"] = one_ms / (dt * 1e3)
        res["query_sharded" if mode == "queries" else "dataset_sharded"] = ent
        mx.close()
    return res, hi, hd


def single_process_main(args, config):
    """python bench.py --single-process --gpus N [--workload ...]: one process, N GPUs behind one handle (cfg5: 100M x 16, 10M queries)"""
    import torch
    pkg = importlib.import_module("semi-convex_hull_tree_b200")
    import oracle as orc
    n, dim, nq, K, pct, kind = WORKLOADS[args.workload]
    data, queries = make_data(kind, n, dim, nq, 1234)
    t0 = time.perf_counter()
    ht = pkg.HostTree(data, pct, 1, device=0)
    t_build = time.perf_counter() - t0
    modes = ("queries", "dataset") if args.mode == "both" else (args.mode,)
    sampler = ClockSampler(0)
    sampler.start()
    res, gi, gd = single_process_runs(pkg, torch, ht, queries, dim, K, list(range(args.gpus)), args.steps, modes, warm=min(1, args.warmup))
    sampler.stop()
    best = max(v["queries_per_s"] for k, v in res.items() if isinstance(v, dict) and "queries_per_s" in v)
    config = dict(config, parallelism="one process, %d GPUs behind one handle (scht_create_multi)" % args.gpus)
    line = {"metric": "exact k-NN queries/sec (K=%d)" % K, "value": best, "unit": "queries/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": min(1, args.warmup), "ms_per_step": 1e3 * nq / best, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": config, "tree": {"n_leaves": int(ht.arrays()["n_leaves"]), "device_build_s": round(t_build, 2)},
            "e2e": {"value": best, "unit": "queries/s", "h2d_bytes_per_step": nq * dim * 4, "d2h_bytes_per_step": nq * K * 8},
            "single_process": res, "clocks": sampler.summary()}
    # parity of a query sample (results of the last timed call) against the C oracle on the same tree
    if not args.no_cpu_baseline:
        m = args.cpu_sample or 64
        t0 = time.perf_counter()
        oi, od2, _, _, cnt = orc.knn(ht, queries[:m], K, alg=2)
        dt = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": m / dt, "unit": "queries/s", "cores": min(os.cpu_count() or 1, 64), "kind": "port",
                                "sample": "first %d queries, C oracle port (OpenMP) on the device-built tree" % m,
                                "ref_scanned_points_per_query": int(cnt[0]) / m}
        line["parity_checked"], line["parity_mismatches"] = parity(gi.numpy(), gd.numpy(), dict(idx=oi, dist2=od2))
    print(json.dumps(line))
    return 1 if line.get("parity_mismatches") else 0


if __name__ == "__main__":
    sys.exit(main())
