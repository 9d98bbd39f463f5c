"""per-kernel histogram of the SASS opcodes that prove (or disprove) a Blackwell-native kernel, from `cuobjdump -sass` of the built library:
UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG / UBLKCP = TMA, UTCBAR / SYNCS = mbarrier traffic, HMMA = legacy mma.sync
(none expected), plus the size of each kernel.  Writes profiles/r2_sass_opcodes.txt."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "cn_chess_ai_b200", "libxq_b200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
kern, hist = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        hist[kern] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and kern:
        hist[kern][m.group(1)] += 1
keys = ["UTCHMMA", "UTCQMMA", "UTCIMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "SYNCS", "HMMA", "IDP", "REDUX", "ATOMG", "RED"]
lines = ["# cuobjdump -sass cn_chess_ai_b200/libxq_b200.so (sm_100a), opcode counts per kernel; `total` = SASS instructions of the kernel",
         "# UTC*MMA = tcgen05.mma | LDTM/STTM = tcgen05.ld/st | UTMALDG/UTMASTG/UBLKCP = TMA | SYNCS = mbarrier | HMMA = mma.sync (legacy; none expected) | IDP = dp4a",
         f"{'kernel':58s} {'total':>6s} " + " ".join(f"{k:>7s}" for k in keys)]
for k, c in hist.items():
    lines.append(f"{k[:58]:58s} {sum(c.values()):6d} " + " ".join(f"{c.get(x, 0):7d}" for x in keys))
tot = collections.Counter()
for c in hist.values():
    tot.update(c)
lines.append(f"{'ALL KERNELS':58s} {sum(tot.values()):6d} " + " ".join(f"{tot.get(x, 0):7d}" for x in keys))
open(os.path.join(ROOT, "profiles", "r2_sass_opcodes.txt"), "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
