"""Static SASS statistics of a kernel between its barriers (phases of the rollout kernels).  usage: sass_phases.py <lib.so> <kernel-substr>"""
import collections
import re
import subprocess
import sys
import tempfile

lib, kname = sys.argv[1], sys.argv[2]
with tempfile.TemporaryDirectory() as d:
    subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=d, capture_output=True)
    out = subprocess.run("for f in xq_*.cubin; do cuobjdump -sass $f; done", shell=True, cwd=d, capture_output=True, text=True).stdout
cur, fn = None, collections.defaultdict(list)
for l in out.split("\n"):
    m = re.search(r"Function : (\S+)", l)
    if m:
        cur = m.group(1)
    elif cur and re.match(r"\s+/\*[0-9a-f]+\*/\s+\S", l):
        fn[cur].append(l.split("*/")[1].split("/*")[0].strip())
for name, lines in fn.items():
    if kname not in name:
        continue
    bars = [i for i, l in enumerate(lines) if "BAR.SYNC" in l]
    print(name[:70], len(lines), "instructions, barriers at", bars)
    for a, b in zip([0] + bars, bars + [len(lines)]):
        c = collections.Counter(re.sub(r"@!?U?P\d+\s+", "", l).split()[0].split(".")[0] for l in lines[a:b])
        print(f"  [{a:5d},{b:5d}) {b - a:5d}", {k: c[k] for k in ("BRA", "BSSY", "SEL", "ISETP", "IMAD", "LOP3", "SHF", "LDS", "STS")})
    break
