"""Collapse an `ncu --page source --csv --print-source sass` export into straight-line segments:
consecutive SASS instructions with the same execution count.  Shows where warp-instructions go
and how many lanes were active there.  usage: python scripts/ncu_segments.py file.csv [min_pct]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
min_pct = float(sys.argv[2]) if len(sys.argv) > 2 else 0.3
hdr_i = next(i for i, r in enumerate(rows) if "Address" in r and "Instructions Executed" in r)
hdr = rows[hdr_i]
c = {n: i for i, n in enumerate(hdr)}
inst = []
for r in rows[hdr_i + 1:]:
    if len(r) < len(hdr) or not r[c["Address"]]:
        continue
    try:
        inst.append((r[c["Address"]], r[c["Source"]].strip(), int(r[c["Instructions Executed"]]),
                     int(r[c["Thread Instructions Executed"]]), int(r[c["# Samples"]])))
    except ValueError:
        pass
tot = sum(i[2] for i in inst)
tsmp = sum(i[4] for i in inst)
print(f"{len(inst)} SASS instructions, {tot:,} warp-inst executed, avg active {sum(i[3] for i in inst) / tot:.2f}")
segs, cur = [], None
for a, s, ie, te, smp in inst:
    if cur and cur["ie"] == ie and ie > 0:
        cur["n"] += 1; cur["te"] += te; cur["smp"] += smp; cur["ops"].append(s.split()[0] if s else "?")
    else:
        cur = {"a": a, "ie": ie, "n": 1, "te": te, "smp": smp, "ops": [s.split()[0] if s else "?"]}
        segs.append(cur)
for sg in segs:
    w = sg["ie"] * sg["n"]
    if 100 * w / tot < min_pct:
        continue
    ops = {}
    for o in sg["ops"]:
        o = o.lstrip("@!P0123456789 ")
        ops[o] = ops.get(o, 0) + 1
    top = " ".join(f"{k}:{v}" for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:7])
    print(f"{sg['a'][-5:]} n={sg['n']:4d} exec={sg['ie'] / 1e6:8.2f}M  {100 * w / tot:5.1f}% inst {100 * sg['smp'] / max(tsmp, 1):5.1f}% smp "
          f"act {sg['te'] / max(w, 1):5.1f}  {top}")
