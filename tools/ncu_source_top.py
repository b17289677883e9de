"""Developer tool: summarise `ncu --page source --csv` output (top stall sites, opcode mix)."""
import csv
import sys
from collections import Counter

path = sys.argv[1]
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = list(csv.reader(open(path)))
# split into kernels at "Kernel Name" rows
kernels, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "data": []}
        kernels.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = r
    elif cur is not None and len(r) == len(cur["hdr"]):
        cur["data"].append(r)
k = kernels[which]
hdr, data = k["hdr"], k["data"]
iS, iSrc, iEx = hdr.index("# Samples"), hdr.index("Source"), hdr.index("Instructions Executed")
tot = sum(int(r[iS]) for r in data)
print(k["name"], "| samples", tot, "| warp-instr executed", sum(int(r[iEx]) for r in data), "| SASS lines", len(data))
for r in sorted(data, key=lambda r: -int(r[iS]))[:40]:
    print(f"{int(r[iS]):6d} {100 * int(r[iS]) / tot:5.1f}%  ex={r[iEx]:>8s}  {r[iSrc].strip()[:120]}")
c = Counter()
for r in data:
    t = r[iSrc].strip().split()
    op = t[1] if t[0].startswith("@") else t[0]
    c[op.split(".")[0]] += int(r[iEx])
print(c.most_common(30))
