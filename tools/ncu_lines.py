"""Source lines of a kernel by warp-stall samples (ncu report captured with --import-source on; read here without a GPU).
usage: python tools/ncu_lines.py report.ncu-rep [top N]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
fname, hdr, acc = "", None, []
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        fname = r[1].split("/")[-1]
    elif r and r[0] == "Line No":
        hdr = r
        i_s, i_x = hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
    elif hdr and len(r) > 8 and r[0].isdigit():
        num = lambda v: int(v) if v.strip().isdigit() else 0
        acc.append((num(r[i_s]), num(r[i_x]), fname, int(r[0]), r[1].strip()[:110]))
tot = sum(a[0] for a in acc) or 1
totx = sum(a[1] for a in acc) or 1
print("stall samples %d, warp instructions %d" % (tot, totx))
for s, x, f, ln, src in sorted(acc, reverse=True)[:top]:
    print("%5.1f%% stall %5.1f%% exec  %s:%d  %s" % (100.0 * s / tot, 100.0 * x / totx, f, ln, src))
