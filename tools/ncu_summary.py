#!/usr/bin/env python
"""Summarises an `ncu --set full` report of one step-kernel launch (raw page as csv) into the few numbers
bench.py's roofline / roofline_issue blocks quote.
usage: ncu -i X.ncu-rep --page raw --csv > raw.csv; tools/ncu_summary.py raw.csv ENVS T [key] >> merged into profiles/r02_ncu_summary.json"""
import csv
import json
import os
import sys


def main():
    raw, envs, T = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    key = sys.argv[4] if len(sys.argv) > 4 else "mo_4096"
    out = sys.argv[5] if len(sys.argv) > 5 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "r02_ncu_summary.json")
    rows = list(csv.reader(open(raw)))
    hdr, vals = rows[0], rows[2]
    d = dict(zip(hdr, vals))

    def f(name):
        return float(d[name].replace(",", ""))
    stalls = {k.split("issue_stalled_")[1].split("_per_")[0]: float(v) for k, v in d.items()
              if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio") and "not_issued" not in k}
    tot = sum(stalls.values())
    inst = f("smsp__inst_executed.sum")
    s = {
        "kernel": d["Kernel Name"], "grid": d["Grid Size"], "block": d["Block Size"],
        "duration_ms_under_ncu": f("gpu__time_duration.sum") / (1e6 if d.get("gpu__time_duration.sum") and float(d["gpu__time_duration.sum"].replace(",", "")) > 1e4 else 1.0),
        "dram_bytes_per_launch": f("dram__bytes_read.sum") * (1e6 if rows[1][hdr.index("dram__bytes_read.sum")] == "Mbyte" else 1e3 if rows[1][hdr.index("dram__bytes_read.sum")] == "Kbyte" else 1e9 if rows[1][hdr.index("dram__bytes_read.sum")] == "Gbyte" else 1)
        + f("dram__bytes_write.sum") * (1e6 if rows[1][hdr.index("dram__bytes_write.sum")] == "Mbyte" else 1e3 if rows[1][hdr.index("dram__bytes_write.sum")] == "Kbyte" else 1e9 if rows[1][hdr.index("dram__bytes_write.sum")] == "Gbyte" else 1),
        "warp_inst_per_launch": inst,
        "warp_inst_per_env_step": inst / (envs * T),
        "inst_executed_pct_of_peak": f("sm__inst_executed.sum.pct_of_peak_sustained_elapsed") if "sm__inst_executed.sum.pct_of_peak_sustained_elapsed" in d else None,
        "issue_active_pct_of_peak_elapsed": f("sm__issue_active.avg.pct_of_peak_sustained_elapsed") if "sm__issue_active.avg.pct_of_peak_sustained_elapsed" in d else None,
        "active_lanes_per_inst": f("smsp__thread_inst_executed_per_inst_executed.ratio"),
        "sm_cycles_active_avg": f("sm__cycles_active.avg"), "sm_cycles_elapsed_max": f("sm__cycles_elapsed.max"),
        "registers_per_thread": f("launch__registers_per_thread"),
        "l1_hit_pct": f("l1tex__t_sector_hit_rate.pct") if "l1tex__t_sector_hit_rate.pct" in d else None,
        "l2_hit_pct": f("lts__t_sector_hit_rate.pct") if "lts__t_sector_hit_rate.pct" in d else None,
        "stall_share_pct": {k: round(100.0 * v / tot, 1) for k, v in sorted(stalls.items(), key=lambda x: -x[1])[:8]},
        "cache_control": "ncu default (--cache-control all): caches flushed before every replay pass, i.e. the cold-L2 condition of the bench's timed steps",
        "source": "profiles/" + os.path.basename(raw),
    }
    cur = json.load(open(out)) if os.path.exists(out) else {}
    cur[key] = s
    json.dump(cur, open(out, "w"), indent=1)
    print(json.dumps(s, indent=1))


if __name__ == "__main__":
    main()
