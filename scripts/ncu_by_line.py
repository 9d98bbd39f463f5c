"""Attribute ncu SASS-level counters (ncu -i X.ncu-rep --page source --csv) to CUDA source lines using
nvdisasm -g line markers.  usage: ncu_by_line.py <src.csv> <cubin> <kernel-substr> <warps*iters> [top]"""
import collections
import csv
import re
import subprocess
import sys

src_csv, cubin, kname, denom = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.split("\n")
cur, insts, ink = None, [], False
for l in dis:
    if l.startswith(".text."):
        ink = kname in l
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if ink:
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m:
            insts.append((m.group(2), cur))
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
data = [r for r in rows[2:] if len(r) >= len(hdr) - 2 and r[0].startswith("0x")][:len(insts)]
ia, isamp = hdr.index("Instructions Executed"), hdr.index("# Samples")
ibar = hdr.index("stall_barrier")
agg = collections.defaultdict(lambda: [0, 0, 0])
for r, (t, c) in zip(data, insts):
    agg[c][0] += int(r[ia]); agg[c][1] += int(r[isamp]); agg[c][2] += int(r[ibar] or 0)
ts = sum(v[1] for v in agg.values())
srcs = {}
tot = 0
for (f, ln), (n, s, b) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    if f not in srcs:
        try:
            srcs[f] = open("/root/repo/cn_chess_ai_b200/csrc/" + f).read().split("\n")
        except OSError:
            srcs[f] = []
    text = srcs[f][ln - 1].strip()[:100] if ln - 1 < len(srcs[f]) else ""
    print(f"{n / denom:7.1f} inst/iter {s / ts * 100:5.1f}% samples (barrier {b / ts * 100:4.1f}%)  {f}:{ln}: {text}")
print("total inst/iter", sum(v[0] for v in agg.values()) / denom)
