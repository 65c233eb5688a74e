#!/usr/bin/env python
"""Condense `ncu -i X.ncu-rep --page raw --csv` into the per-kernel table kept under profiles/ (one column per launch).

    python tools/summarize_ncu.py gpurun_out/ncu_denoise_raw.csv > profiles/rN_ncu_denoise_summary.txt
    python tools/summarize_ncu.py raw.csv --traffic-json profiles/atrous_traffic.json   # also refresh bench.py's `traffic`
"""
import argparse
import csv
import json

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sector_hit_rate.pct",
    "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "sm__cycles_elapsed.avg", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
]


def num(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("raw_csv")
    ap.add_argument("--traffic-json")
    a = ap.parse_args()
    rows = list(csv.reader(open(a.raw_csv)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    names = [r[col["Kernel Name"]] for r in data]
    short = [n.split("(")[0].replace("void ", "").replace("<unnamed>::", "")[:44] for n in names]
    print("source: %s (ncu --set full --clock-control none; per-launch values, cold-cache and serialised)" % a.raw_csv)
    print("%-72s %s" % ("kernel", short))
    for m in METRICS:
        if m in col:
            vals = []
            for r in data:
                v = num(r[col[m]])
                vals.append(r[col[m]] if v is None else ("%.4g" % v))
            print("%-72s %-8s %s" % (m, units[col[m]], vals))
    print("warp stall reasons (warps stalled per issue-active cycle, per launch):")
    for h in hdr:
        if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
            vals = [num(r[col[h]]) or 0.0 for r in data]
            if max(vals) >= 0.25:
                print("  %-40s %s" % (h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")], ["%.2f" % v for v in vals]))
    if a.traffic_json:
        rd, wr = col["dram__bytes_read.sum"], col["dram__bytes_write.sum"]

        def to_bytes(r, c):
            scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[units[c]]
            return num(r[c]) * scale
        tiled = [to_bytes(r, rd) + to_bytes(r, wr) for r, n in zip(data, names) if "atrous_tiled" in n]
        kl = [to_bytes(r, rd) + to_bytes(r, wr) for r, n in zip(data, names) if "atrous_kl" in n]
        n = min(len(tiled), len(kl))
        per_level = [tiled[i] + kl[i] for i in range(n)]
        out = {"source": a.raw_csv + " (ncu --set full --clock-control none, bench.py C2, whole frames)",
               "kernel": "atrous_tiled_kernel (+ atrous_kl_kernel pre-pass)", "per_launch_tiled_bytes": tiled, "per_launch_kl_bytes": kl,
               "dram_bytes_per_launch": sum(per_level) / max(n, 1),
               "algorithmic_bytes_per_launch_1080p": [116121600, 116121600, 116121600, 116121600, 141004800]}
        # bench.py reports `roofline.traffic` from this file only while the a-trous kernel sources are the ones captured
        import os, sys
        sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        import bench
        out["kernel_source_hash"] = bench.kernel_source_hash()
        json.dump(out, open(a.traffic_json, "w"), indent=1)


if __name__ == "__main__":
    main()
