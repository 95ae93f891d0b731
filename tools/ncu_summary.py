#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page) into a small text file for profiles/.
    python tools/ncu_summary.py <report.ncu-rep> [out.txt]
"""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_dynamic", "launch__grid_size",
        "launch__block_size", "launch__waves_per_multiprocessor", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        out.append("kernel: %s" % r[hdr.index("Kernel Name")])
        for k in KEYS:
            if k in hdr:
                out.append("  %-62s %s %s" % (k, r[hdr.index(k)], units[hdr.index(k)]))
        st = []
        for i, k in enumerate(hdr):
            if "average_warps_issue_stalled" in k and k.endswith("per_issue_active.ratio"):
                try:
                    st.append((float(r[i]), k.split("issue_stalled_")[1].replace("_per_issue_active.ratio", "")))
                except ValueError:
                    pass
        out.append("  warp stalls per issued instruction: " + ", ".join("%s %.2f" % (n, v) for v, n in sorted(st, reverse=True)[:8]))
    text = "\n".join(out) + "\n"
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(text)
    else:
        sys.stdout.write(text)


if __name__ == "__main__":
    main()
