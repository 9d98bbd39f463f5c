"""Headline counters + warp-stall breakdown (stalled warps per issued instruction) of every launch in an .ncu-rep.  usage: ncu_stalls.py <in.ncu-rep>"""
import csv
import subprocess
import sys

raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h = rows[0]
for r in rows[2:]:
    d = dict(zip(h, r))
    print(d["Kernel Name"][:60])
    for k in ("gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
              "launch__registers_per_thread", "launch__grid_size", "smsp__thread_inst_executed_per_inst_executed.ratio", "dram__bytes_read.sum", "dram__bytes_write.sum"):
        print("  ", k, d.get(k))
    st = {k: float(v) for k, v in d.items() if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("_per_issue_active.ratio")}
    for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:8]:
        print("     %-28s %.3f" % (k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), v))
