"""Instruction-class census of every kernel in libsrt_b200.so (static SASS, `cuobjdump -sass`): the mnemonics that show
what the sm_100a build uses -- bulk async copies (UBLKCP), packed FP32x2 (FFMA2 / FMUL2 / FADD2), mbarrier traffic
(SYNCS), warp votes / shuffles / matches, MUFU, uniform constant loads (LDCU), and that no tensor-core instruction is
present (the path has no dense contraction).   python scripts/sass_classes.py [lib] > profiles/r2_sass_classes.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "simple_raytracer_b200", "libsrt_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
CLASSES = ["UBLKCP", "SYNCS", "FFMA2", "FMUL2", "FADD2", "FFMA", "FMUL", "FADD", "MUFU", "DFMA", "VOTE", "SHFL", "MATCH",
           "REDUX", "LDS", "STS", "LDG", "STG", "LDCU", "LDC", "ATOM", "RED", "BAR", "HMMA", "UTCHMMA", "UTCMMA", "LDTM",
           "UTMALDG", "WARPSYNC", "CALL"]
cur, counts, totals = None, {}, {}
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", cur)
        counts[cur], totals[cur] = collections.Counter(), 0
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        op = m.group(1).split(".")[0]
        totals[cur] += 1
        counts[cur][op] += 1
print(f"# {os.path.relpath(lib, ROOT)}: static SASS instruction classes per kernel (cuobjdump -sass)")
print(f"{'kernel':58s} {'total':>6s} " + " ".join(f"{c:>7s}" for c in CLASSES if any(counts[k][c] for k in counts)))
used = [c for c in CLASSES if any(counts[k][c] for k in counts)]
for k in counts:
    print(f"{k[:58]:58s} {totals[k]:6d} " + " ".join(f"{counts[k][c]:7d}" for c in used))
absent = [c for c in ("HMMA", "UTCHMMA", "UTCMMA", "LDTM", "UTMALDG") if not any(counts[k][c] for k in counts)]
print("absent by design (no dense contraction on this path):", ", ".join(absent))
