"""Turn the raw captures in gpurun_out/ into the committed evidence under profiles/ (run here, after a gpurun call):
ncu key metrics + SASS segments per config, roofline_traffic.json, launch-list shares, bench lines."""
import collections
import csv
import json
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
rnd = sys.argv[1] if len(sys.argv) > 1 else "r1"

traffic = {}
for c in (2, 3, 5):
    rep = os.path.join(G, f"{rnd}_prof_c{c}.ncu-rep")
    if not os.path.exists(rep):
        continue
    key = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_key.py"), rep], capture_output=True, text=True).stdout
    open(os.path.join(P, f"{rnd}_ncu_c{c}_key_metrics.txt"), "w").write(key)
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    tmp = f"/tmp/src_c{c}.csv"
    open(tmp, "w").write(src)
    seg = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_segments.py"), tmp, "0.8"], capture_output=True, text=True).stdout
    open(os.path.join(P, f"{rnd}_ncu_c{c}_segments.txt"), "w").write(seg)

    def val(k):
        m = re.search(re.escape(k) + r"\s+([0-9.]+)\s+(\S+)", key)
        return float(m.group(1)), m.group(2)
    unit = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}
    rd, u1 = val("dram__bytes_read.sum")
    wr, u2 = val("dram__bytes_write.sum")
    traffic[f"config{c}"] = int(rd * unit[u1] + wr * unit[u2])
    traffic[f"config{c}_inst_executed"] = int(val("smsp__inst_executed.sum")[0])
    traffic[f"config{c}_issue_active_pct"] = val("smsp__issue_active.avg.pct_of_peak_sustained_active")[0]
    traffic[f"config{c}_capture"] = (f"profiles/{rnd}_ncu_c{c}_key_metrics.txt (ncu --set full --clock-control none, one render_kernel "
                                     f"launch of scripts/profile_target.py {c}, num_samples 4)")
if traffic:
    json.dump(traffic, open(os.path.join(P, "roofline_traffic.json"), "w"), indent=1)

for f in (f"{rnd}_bench.json", f"{rnd}_bench_reference.json", f"{rnd}_bench_launches.csv", f"{rnd}_configs_n1.jsonl",
          f"{rnd}_bench_n2.json", f"{rnd}_configs_n2.jsonl", f"{rnd}_bench_n4.json", f"{rnd}_configs_n4.jsonl",
          f"{rnd}_bench_n8.json", f"{rnd}_configs_n8.jsonl"):
    if os.path.exists(os.path.join(G, f)):
        shutil.copy(os.path.join(G, f), os.path.join(P, f))

ll = os.path.join(P, f"{rnd}_bench_launches.csv")
if os.path.exists(ll):
    rows = list(csv.reader(open(ll)))
    hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    h = rows[hi]
    kn, mv, mu = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[hi + 1:]:
        if len(r) <= mv:
            continue
        v = float(r[mv].replace(",", ""))
        ms = v / 1e6 if r[mu].startswith("ns") else v / 1e3 if r[mu].startswith("us") else v
        agg[r[kn][:90]][0] += 1
        agg[r[kn][:90]][1] += ms
    tot = sum(v[1] for v in agg.values())
    with open(os.path.join(P, f"{rnd}_bench_launch_shares.txt"), "w") as f:
        f.write("launches   total ms   share  kernel   (ncu --metrics gpu__time_duration.sum --clock-control none, bench.py --steps 2 --warmup 3)\n")
        for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{n:8d} {ms:10.3f} {100 * ms / tot:6.1f}%  {k}\n")
    print(open(os.path.join(P, f"{rnd}_bench_launch_shares.txt")).read())
