"""Summarise `ncu --set full` captures for profiles/: key raw metrics per kernel + dram bytes per launch.

    python tools/ncu_summary.py OUT.txt NAME=report.ncu-rep [NAME=report.ncu-rep ...]   (also updates profiles/traffic.json)
"""
import csv
import io
import json
import os
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.max"]
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out_path = sys.argv[1]
traffic_path = os.path.join(ROOT, "profiles", "traffic.json")
traffic = json.load(open(traffic_path)) if os.path.exists(traffic_path) else {}
lines = []
for spec in sys.argv[2:]:
    name, rep = spec.split("=", 1)
    if not os.path.exists(rep):   # (the capture's kernel filter matched nothing: say so, keep the other summaries)
        lines.append("## %s: no report (%s)" % (name, rep))
        continue
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = next((i for i, r in enumerate(rows) if r and r[0] == "ID"), None)
    if hdr is None:
        lines.append("## %s: unreadable report (%s)" % (name, rep))
        continue
    cols, units, vals = rows[hdr], rows[hdr + 1], rows[hdr + 2]
    ix = {c: i for i, c in enumerate(cols)}
    lines.append("## %s: %s" % (name, vals[ix["Kernel Name"]]))
    dram = 0.0
    for k in KEYS:
        if k not in ix:
            continue
        v, u = vals[ix[k]], units[ix[k]]
        lines.append("  %s = %s %s" % (k, v, u))
        if k.startswith("dram__bytes"):
            dram += float(v.replace(",", "")) * UNIT.get(u, 1.0)
    lines.append("  dram traffic per launch = %.2f MB" % (dram / 1e6))
    traffic[name] = dram
open(out_path, "w").write("\n".join(lines) + "\n")
json.dump(traffic, open(traffic_path, "w"))
print("\n".join(lines))
