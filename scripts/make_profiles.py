#!/usr/bin/env python
"""Turn the raw artefacts a gpurun call left in gpurun_out/ into the tracked summaries under profiles/.

  python scripts/make_profiles.py launches gpurun_out/r01l_launches.csv profiles/r01l_launches_summary.md "title"
  python scripts/make_profiles.py ncu      gpurun_out/r01l_seq_encode.ncu-rep profiles/r01l_seq_encode_tc_ncu.md "title"
"""
import collections
import csv
import io
import subprocess
import sys


def launches(src, dst, title):
    with open(src) as fh:
        lines = [l for l in fh if not l.startswith("==")]
    r = csv.reader(lines)
    hdr = next(r)
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg, tot = collections.OrderedDict(), 0.0
    for row in r:
        if len(row) <= vi:
            continue
        v = float(row[vi].replace(",", ""))
        if row[ui] == "ns":
            v /= 1e3
        name = row[ki].split("(")[0].replace("void ", "").replace("dmt::<unnamed>::", "").replace("dmt::", "")
        if "at_cuda_detail" in name or "native::" in name or "at::" in name:
            name = "[torch plumbing] " + name.split("::")[-1][:40]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
        tot += v
    with open(dst, "w") as out:
        out.write("# %s\n\n`ncu --metrics gpu__time_duration.sum --clock-control none` launch list (cold-cache, serialised: "
                  "compare SHARES, not absolute times).  Source: `%s`.\n\n" % (title, src))
        out.write("| kernel | launches | total us | avg us | share |\n|---|---|---|---|---|\n")
        for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
            out.write("| %s | %d | %.1f | %.1f | %.3f |\n" % (k, n, t, t / n, t / tot))
        out.write("\ntotal %.1f us over %d launches\n" % (tot, sum(n for n, _ in agg.values())))


KEYS = ["launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
        "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def ncu(src, dst, title, max_launches=16):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:2 + max_launches]
    with open(dst, "w") as out:
        out.write("# %s\n\n`ncu --set full --clock-control none --import-source on`, read back with `ncu -i ... --page raw "
                  "--csv`.  Source: `%s` (kept in gpurun_out/, not tracked).\n\n" % (title, src))
        ki = hdr.index("Kernel Name")
        out.write("| metric | unit | " + " | ".join("L%d" % i for i in range(len(data))) + " |\n")
        out.write("|---|---|" + "---|" * len(data) + "\n")
        out.write("| kernel | | " + " | ".join(r[ki].split("(")[0].replace("void ", "")[-28:] for r in data) + " |\n")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                out.write("| %s | %s | %s |\n" % (k, units[i], " | ".join(r[i][:10] for r in data)))
        stall = [(h, i) for i, h in enumerate(hdr)
                 if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio")]
        out.write("\nTop warp-stall reasons (stalled warps per issue-active cycle):\n\n")
        for li in range(len(data)):
            top = sorted(stall, key=lambda x: -float(data[li][x[1]] or 0))[:5]
            out.write("* L%d: %s\n" % (li, ", ".join("%s %.2f" % (h.replace("smsp__average_warps_issue_stalled_", "")
                                                               .replace("_per_issue_active.ratio", ""),
                                                               float(data[li][i] or 0)) for h, i in top)))


if __name__ == "__main__":
    kind, src, dst, title = sys.argv[1:5]
    {"launches": launches, "ncu": ncu}[kind](src, dst, title)
    print("wrote", dst)
