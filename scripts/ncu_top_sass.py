"""Top stall-sampled SASS instructions of one kernel: ncu -i X.ncu-rep --page source --csv --kernel-name regex:K > src.csv; ncu_top_sass.py src.csv [top]"""
import csv, sys
rows=list(csv.reader(open(sys.argv[1])))
top=int(sys.argv[2]) if len(sys.argv)>2 else 40
hdr=rows[1]
isrc=hdr.index("Source"); isamp=hdr.index("# Samples"); iex=hdr.index("Instructions Executed")
stalls=[i for i,h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
data=[r for r in rows[2:] if len(r)>max(isamp,iex) and r[0].startswith("0x")]
tot=sum(int(r[isamp]) for r in data)
print("total samples",tot,"total warp-instr",sum(int(r[iex]) for r in data), "n sass", len(data))
for k,r in enumerate(data): r.append(k)
for r in sorted(data,key=lambda r:-int(r[isamp]))[:top]:
    s=sorted(((int(r[i] or 0),hdr[i][6:]) for i in stalls),reverse=True)[:2]
    print(f"#{r[-1]:4d} {100*int(r[isamp])/tot:5.1f}% ex={r[iex]:>7s} {r[isrc].strip()[:70]:70s} {s}")
