"""Source line of every branch (BRA / BSSY) inside the ply loop of a rollout kernel.  usage: sass_branches.py <lib.so> <cubin-name-substr> <first-barrier-index> <last-barrier-index>"""
import re
import subprocess
import sys
import tempfile

lib, cub, b0, b1 = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
with tempfile.TemporaryDirectory() as d:
    subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=d, capture_output=True)
    dis = subprocess.run(f"nvdisasm -g -c xq_{cub}.sm_100a.cubin", shell=True, cwd=d, capture_output=True, text=True).stdout
cur, out = None, []
for l in dis.split("\n"):
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        out.append((int(m.group(1), 16), m.group(2).strip(), cur))
bars = [i for i, (a, t, c) in enumerate(out) if "BAR.SYNC" in t]
print("barriers", bars)
for i in range(bars[b0], bars[b1]):
    a, t, c = out[i]
    if re.search(r"\b(BRA|BSSY|BAR)\b", t):
        print(i, hex(a), t[:60], c)
