"""Turn the raw captures in gpurun_out/ into the committed evidence under profiles/ (run here, after a gpurun call of
scripts/r2_final.sh): ncu key metrics + SASS segments per config, roofline_traffic.json (what bench.py copies into
`roofline.traffic` / `fma_pipe` / `issue`), launch-list shares, bench lines, SASS instruction classes.
    python scripts/make_profiles.py r2"""
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
rnd = sys.argv[1] if len(sys.argv) > 1 else "r2"
UNIT = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1, "Tbyte": 1e12}
# launches covered by the captured kernel: config 2 is captured as bench.py times it (one 16-launch batch + accumulate)
LAUNCHES = {2: 16, 3: 1, 5: 1}


def kernels_of(key_text):
    """Split ncu_key.py's output into one dict per captured kernel."""
    out = []
    for block in key_text.split("kernel:")[1:]:
        d = {"name": block.splitlines()[0].strip()}
        for line in block.splitlines()[1:]:
            m = re.match(r"\s+(\S+)\s+([0-9.]+)\s*(\S*)", line)
            if m:
                d[m.group(1)] = (float(m.group(2)), m.group(3))
        out.append(d)
    return out


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
    ks = kernels_of(key)
    render = next(k for k in ks if "render_kernel" in k["name"])
    dram = sum(k[m][0] * UNIT[k[m][1]] for k in ks for m in ("dram__bytes_read.sum", "dram__bytes_write.sum") if m in k)
    n = LAUNCHES[c]
    traffic[f"config{c}"] = {
        "dram_bytes_per_launch": int(dram / n),
        "dram_bytes_kernels": [k["name"].split("(")[0] for k in ks],
        "warp_inst_per_launch": int(render["smsp__inst_executed.sum"][0] / n),
        "issue_active_pct": render["smsp__issue_active.avg.pct_of_peak_sustained_active"][0],
        "fma_pipe": {"inst_executed_pipe_fma_pct": render["sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"][0],
                     "pipe_fma_cycles_active_pct": render["sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"][0]},
        "warp_execution_efficiency": render["smsp__thread_inst_executed_per_inst_executed.ratio"][0],
        "registers_per_thread": int(render["launch__registers_per_thread"][0]),
        "capture": (f"profiles/{rnd}_ncu_c{c}_key_metrics.txt (ncu --set full --clock-control none; "
                    + ("one 16-launch srt_render_batch kernel + its accumulate_kernel, as bench.py's timed step runs them"
                       if c == 2 else f"one render_kernel launch of scripts/profile_target.py {c}, num_samples 4") + ")")}
if traffic:  # configs without a fresh capture keep their entries
    tj = os.path.join(P, "roofline_traffic.json")
    merged = json.load(open(tj)) if os.path.exists(tj) else {}
    merged.update(traffic)
    json.dump(merged, open(tj, "w"), indent=1)

for f in (f"{rnd}_bench.json", f"{rnd}_bench_reference.json", f"{rnd}_bench_launches.csv", f"{rnd}_environment.txt",
          f"{rnd}_pytest_gpu.log", f"{rnd}_fma_operands.txt", f"{rnd}_variants.txt"):
    if os.path.exists(os.path.join(G, f)):
        shutil.copy(os.path.join(G, f), os.path.join(P, f))
for n in (2, 4, 8):
    for f in (f"{rnd}_bench_n{n}.json", f"{rnd}_bench_reference_n{n}.json", f"{rnd}_mgpu_check_n{n}.json"):
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
        f.write("launches   total ms   share  kernel   (ncu --metrics gpu__time_duration.sum --clock-control none, "
                "bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras)\n")
        for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{n:8d} {ms:10.3f} {100 * ms / tot:6.1f}%  {k}\n")
    print(open(os.path.join(P, f"{rnd}_bench_launch_shares.txt")).read())

sass = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "sass_classes.py")], capture_output=True, text=True).stdout
open(os.path.join(P, f"{rnd}_sass_classes.txt"), "w").write(sass)
print(json.dumps(traffic, indent=1))
