"""Turn one tools/round_measure.sh run (gpurun_out/<R>_*) into the tracked summaries under profiles/:
    python tools/profiles_from_run.py r02g r02 [compressible|plain]       (memory of the observation tensor in that run)
<out>_step_<workload>_{ncu_full.txt, sass_top.txt, metrics.json}, <out>_bench_def_small.json, <out>_bench_reference_arm.json,
<out>_launches_def_small.csv, <out>_e2e_sweep.txt and profiles/traffic.json (DRAM bytes per launch, read by bench.py)."""
import csv
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"]
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main():
    run, out = sys.argv[1], sys.argv[2]
    memory = sys.argv[3] if len(sys.argv) > 3 else "compressible"
    G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
    traffic_path = os.path.join(P, "traffic.json")
    traffic = json.load(open(traffic_path)) if os.path.exists(traffic_path) else {}
    for wl in ("def-small", "atk-small", "def-middle-multi", "2p-large"):
        raw = os.path.join(G, "%s_step_%s_raw.csv" % (run, wl))
        if not os.path.exists(raw):
            continue
        rows = list(csv.reader(open(raw)))
        hdr, units, vals = rows[0], rows[1], rows[2]
        m = {"kernel": vals[hdr.index("Kernel Name")]}
        for k in KEYS:
            if k in hdr:
                m[k] = {"value": vals[hdr.index(k)], "unit": units[hdr.index(k)]}
        stalls = {}
        for i, k in enumerate(hdr):
            if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio") or \
               (k.startswith("smsp__average_warp") and "issue_stalled" in k and k.endswith(".ratio")):
                try:
                    stalls[k.split("issue_stalled_")[1].split("_per_issue")[0]] = float(vals[i])
                except ValueError:
                    pass
        m["stall_cycles_per_issued_instruction"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1])[:8])
        json.dump(m, open(os.path.join(P, "%s_step_%s_metrics.json" % (out, wl)), "w"), indent=1)
        shutil.copy(os.path.join(G, "%s_step_%s_details.txt" % (run, wl)), os.path.join(P, "%s_step_%s_ncu_full.txt" % (out, wl)))
        shutil.copy(os.path.join(G, "%s_step_%s_sass_top.txt" % (run, wl)), os.path.join(P, "%s_step_%s_sass_top.txt" % (out, wl)))
        rd = float(m["dram__bytes_read.sum"]["value"]) * SCALE[m["dram__bytes_read.sum"]["unit"]]
        wr = float(m["dram__bytes_write.sum"]["value"]) * SCALE[m["dram__bytes_write.sum"]["unit"]]
        traffic.setdefault(wl, {})[memory] = {
            "dram_bytes_per_launch": rd + wr, "dram_bytes_read": rd, "dram_bytes_write": wr,
            "source": "ncu --set full --clock-control none, one launch after 1,305 steps (profiles/%s_step_%s_metrics.json)" % (out, wl)}
    json.dump(traffic, open(traffic_path, "w"), indent=1)
    for src, dst in (("%s_bench.json" % run, "%s_bench_def_small.json" % out), ("%s_bench_reference.json" % run, "%s_bench_reference_arm.json" % out),
                     ("%s_launches.csv" % run, "%s_launches_def_small.csv" % out), ("%s_e2e_sweep.txt" % run, "%s_e2e_sweep.txt" % out)):
        if os.path.exists(os.path.join(G, src)):
            if src.endswith("_bench.json"):
                d = json.loads(open(os.path.join(G, src)).read().strip().splitlines()[-1])
                json.dump(d, open(os.path.join(P, dst), "w"), indent=1)
            else:
                shutil.copy(os.path.join(G, src), os.path.join(P, dst))


if __name__ == "__main__":
    main()
