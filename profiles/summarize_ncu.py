#!/usr/bin/env python
"""Summarise an `ncu --set full` report into the small JSON committed under
profiles/ (the .ncu-rep itself stays in gpurun_out/, which is scratch).

  python profiles/summarize_ncu.py gpurun_out/prof.ncu-rep profiles/out.json [kernel-substring] [min-us]

Launches shorter than min-us (default 20) are dropped: the solver's kernel sequence is static, so
some launches find nothing to do (block still open / LP already terminal) and return at once.
"""
import csv
import io
import json
import subprocess
import sys

KEEP = ["dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum",
        "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__waves_per_multiprocessor",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_warps", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__inst_executed.sum", "smsp__inst_executed.sum", "sm__cycles_active.avg",
        "lts__t_bytes.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"]
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12,
         "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}


def main():
    rep, out = sys.argv[1], sys.argv[2]
    sub = sys.argv[3] if len(sys.argv) > 3 else ""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True,
                         text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    launches = []
    for r in rows[2:]:
        if sub and sub not in r[idx["Kernel Name"]]:
            continue
        d = {"kernel": r[idx["Kernel Name"]], "grid": r[idx["Grid Size"]],
             "block": r[idx["Block Size"]]}
        for k in KEEP:
            if k in idx and r[idx[k]] != "":
                try:
                    v = float(r[idx[k]].replace(",", ""))
                except ValueError:
                    continue
                u = units[idx[k]]
                if u in SCALE and ("bytes" in k or "duration" in k):
                    v *= SCALE[u]
                    u = "byte" if "bytes" in k else "us"
                d[k] = v
                d[k + "__unit"] = u
        launches.append(d)
    min_us = float(sys.argv[4]) if len(sys.argv) > 4 else 20.0
    launches = [l for l in launches if l.get("gpu__time_duration.sum", 0.0) >= min_us]
    n = max(len(launches), 1)
    rd = sum(l.get("dram__bytes_read.sum", 0.0) for l in launches) / n
    wr = sum(l.get("dram__bytes_write.sum", 0.0) for l in launches) / n
    dur = sum(l.get("gpu__time_duration.sum", 0.0) for l in launches) / n
    summary = {"report": rep, "launches_captured": len(launches),
               "dram_bytes_read_per_launch": rd, "dram_bytes_write_per_launch": wr,
               "dram_bytes_per_launch": rd + wr, "avg_duration_us_under_ncu": dur,
               "launches": launches}
    json.dump(summary, open(out, "w"), indent=1)
    print(json.dumps({k: v for k, v in summary.items() if k != "launches"}))


if __name__ == "__main__":
    main()
