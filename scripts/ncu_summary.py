#!/usr/bin/env python
"""Summarise an .ncu-rep (read here on the CPU box): key metrics per launch + top stall lines."""
import csv, io, subprocess, sys, collections

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
keys = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "lts__t_bytes.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__cycles_active.avg"]
for k in keys:
    if k in hdr:
        i = hdr.index(k)
        print("%-70s %-10s %s" % (k, units[i], " | ".join(r[i][:40] for r in data)))
# stall breakdown
print("\n-- warp stall reasons (pct of warp-active), launch 0")
for i, h in enumerate(hdr):
    if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio"):
        print("  %-80s %s" % (h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), data[0][i]))
if len(sys.argv) > 2:
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda", "--launch-skip", "0", "--launch-count", "1"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    if rows:
        h = rows[0]
        print(h[:12])
        try:
            si = h.index("# Samples") if "# Samples" in h else [i for i, x in enumerate(h) if "Sampl" in x][0]
        except Exception:
            si = None
        if si is not None:
            tot = sum(float(r[si] or 0) for r in rows[1:] if len(r) > si and r[si].replace('.', '').isdigit())
            top = sorted((r for r in rows[1:] if len(r) > si and r[si].replace('.', '').isdigit()), key=lambda r: -float(r[si]))[:int(sys.argv[2])]
            for r in top:
                print("%6.2f%%  L%s  %s" % (100 * float(r[si]) / max(tot, 1), r[0], r[1][:110] if len(r) > 1 else ""))
