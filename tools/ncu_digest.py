"""Condenses an ncu report (.ncu-rep, read here without a GPU) into the small files kept under profiles/:
  <out>_ncu_full.csv   selected raw-page metrics (duration, DRAM/L2 traffic, pipe utilisation, stall reasons, launch shape)
  <out>_hotspots.txt   instruction classes by executed count and by stall samples (source page)
usage: python tools/ncu_digest.py gpurun_out/prof.ncu-rep profiles/r01c_k_scan_tc"""
import csv
import io
import subprocess
import sys

KEEP = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_issued.avg.per_cycle_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "launch__registers_per_thread", "launch__block_size",
        "launch__grid_size", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor")


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep, out = sys.argv[1], sys.argv[2]
    rows = page(rep, "raw")
    hdr, units, vals = rows[0], rows[1], rows[-1]
    with open(out + "_ncu_full.csv", "w") as f:
        w = csv.writer(f)
        w.writerow(["metric", "value", "unit"])
        w.writerow(["kernel", vals[hdr.index("Kernel Name")], ""])
        for i, h in enumerate(hdr):
            if h in KEEP or (h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")):
                w.writerow([h, vals[i], units[i]])
    src = page(rep, "source")
    h2 = src[1]
    isrc, iex, ismp = h2.index("Source"), h2.index("Instructions Executed"), h2.index("Warp Stall Sampling (All Samples)")
    ex, smp = {}, {}
    for r in src[2:]:
        try:
            toks = r[isrc].split()
            op = toks[1] if toks[0].startswith("@") else toks[0]
            op = op.rstrip(";")
            ex[op] = ex.get(op, 0) + int(r[iex])
            smp[op] = smp.get(op, 0) + int(r[ismp])
        except Exception:
            pass
    te, ts = max(1, sum(ex.values())), max(1, sum(smp.values()))
    with open(out + "_hotspots.txt", "w") as f:
        f.write("kernel: %s\nwarp instructions executed: %d, stall samples: %d\n\n" % (vals[hdr.index("Kernel Name")], te, ts))
        f.write("%-36s %8s %8s\n" % ("instruction", "% exec", "% stall samples"))
        for op, v in sorted(ex.items(), key=lambda x: -x[1])[:28]:
            f.write("%-36s %7.2f%% %7.2f%%\n" % (op, 100.0 * v / te, 100.0 * smp.get(op, 0) / ts))
    print("wrote", out + "_ncu_full.csv", out + "_hotspots.txt")


if __name__ == "__main__":
    main()
