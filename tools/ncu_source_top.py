"""Top stall-sample instructions from `ncu --page source --csv` output (needs -lineinfo + --import-source on).
    ncu -i rep --page source --csv --kernel-id ::name:N > src.csv ; python tools/ncu_source_top.py src.csv [topN]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
h = next(i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r)
hdr = rows[h]
iS, iSrc = hdr.index("# Samples"), hdr.index("Source")
data = []
for r in rows[h + 1:]:
    if len(r) <= iS:
        continue
    try:
        data.append((int(r[iS] or 0), r[iSrc].strip()))
    except ValueError:
        pass
tot = sum(d[0] for d in data) or 1
print("total samples", tot, "instructions", len(data))
top = sorted(enumerate(data), key=lambda x: -x[1][0])[:topn]
for i, (s, src) in sorted(top):
    print(f"{i:6d} {s:7d} {100 * s / tot:5.1f}%  {src[:110]}")
