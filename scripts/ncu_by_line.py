"""Aggregate an `ncu --page source --csv --print-source sass,cuda` export by CUDA source line.
usage: python scripts/ncu_by_line.py file.csv [topN]"""
import csv
import sys
from collections import defaultdict

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path)))
# the export has a file header per source file followed by a table; find the SASS table (has "Address")
hdr_i = next(i for i, r in enumerate(rows) if "Address" in r and "Instructions Executed" in r)
hdr = rows[hdr_i]
col = {n: i for i, n in enumerate(hdr)}
src_i = [i for i, n in enumerate(hdr) if n == "Source"]
agg = defaultdict(lambda: [0, 0, 0])
total = [0, 0, 0]
for r in rows[hdr_i + 1:]:
    if len(r) < len(hdr):
        continue
    try:
        ie = int(r[col["Instructions Executed"]]); te = int(r[col["Thread Instructions Executed"]])
        smp = int(r[col["# Samples"]])
    except ValueError:
        continue
    key = r[col["Line No"]] if "Line No" in col else "?"
    # first "Source" column is the CUDA line text when print-source sass,cuda
    agg[(key, r[src_i[0]].strip()[:90])][0] += ie
    agg[(key, r[src_i[0]].strip()[:90])][1] += te
    agg[(key, r[src_i[0]].strip()[:90])][2] += smp
    total[0] += ie; total[1] += te; total[2] += smp
print(f"total warp-inst {total[0]:,}  thread-inst {total[1]:,}  avg active {total[1] / max(total[0], 1):.2f}  samples {total[2]:,}")
for (k, s), (ie, te, smp) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100 * ie / total[0]:5.1f}% inst  {100 * smp / max(total[2], 1):5.1f}% smp  act {te / max(ie, 1):5.1f}  L{k:>5}  {s}")
