#!/bin/bash
# Round-2 last refresh on one B200 after SRT_RING_V4 became the default: canvas digests, GPU suite, ncu --set full capture of
# config 2's timed step (roofline counters of THIS build), both bench arms.
mkdir -p gpurun_out
T=${1:-r2}
timeout 200 python scripts/variant_time.py 2 4 1 3 2>&1 | tee gpurun_out/${T}_final_digests.txt
( time timeout 600 python -m pytest tests -m gpu -q ) > gpurun_out/${T}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed" gpurun_out/${T}_pytest_gpu.log | tail -2
timeout 200 scripts/ncu_capture.sh 2 ${T}_prof_c2 16 - batch
timeout 200 python scripts/make_profiles.py ${T} > gpurun_out/${T}_make_profiles.log 2>&1; cp profiles/roofline_traffic.json gpurun_out/${T}_roofline_traffic.json
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_ref_err.log; echo "ref rc=$?"
( time timeout 400 python bench.py ) > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench_err.log; echo "bench rc=$?"; tail -3 gpurun_out/${T}_bench_err.log
python - $T <<'PY'
import json, sys
T = sys.argv[1]
r = json.loads([l for l in open(f'gpurun_out/{T}_bench_reference.json') if l.startswith('{')][-1])
j = json.loads([l for l in open(f'gpurun_out/{T}_bench.json') if l.startswith('{')][-1])
print('value', j['value'], 'e2e', j['e2e']['value'], 'frac', j['roofline']['frac'], 'ref', r['value'], 'e2e/ref', j['e2e']['value'] / r['value'])
for c in j['configs'] or []:
    print(c['config']['workload'][:28], c['accel'], 'value %.1f' % c['value'], 'frac %.3f' % c['roofline']['frac'], 'launch_ms %.3f' % c['roofline']['launch_ms'])
PY
