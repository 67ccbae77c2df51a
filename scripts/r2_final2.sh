#!/bin/bash
# Round-2 closing evidence pass on one B200 (the analytic kernel changed at the end of the round: pair scan, division-free
# item mapping, range-tested normalize / packed sqrt, carried path state): GPU suite, smoke, a fresh ncu --set full capture
# of config 2's timed step (so that bench.py's roofline carries this build's counters), both bench arms, ncu launch list.
mkdir -p gpurun_out
T=${1:-r2}
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm,memory.total --format=csv > gpurun_out/${T}_environment.txt
echo "host cores: $(nproc)" >> gpurun_out/${T}_environment.txt; lscpu | grep -E "Model name" >> gpurun_out/${T}_environment.txt
( time timeout 900 python -m pytest tests -m gpu -q ) > gpurun_out/${T}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed" gpurun_out/${T}_pytest_gpu.log | tail -2
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 scripts/ncu_capture.sh 2 ${T}_prof_c2 16 - batch
timeout 300 python scripts/make_profiles.py ${T} > gpurun_out/${T}_make_profiles.log 2>&1; cp profiles/roofline_traffic.json gpurun_out/${T}_roofline_traffic.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_ref_err.log; echo "ref rc=$?"
( time timeout 600 python bench.py ) > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench_err.log; echo "bench rc=$?"; tail -3 gpurun_out/${T}_bench_err.log
python - $T <<'PY'
import json, sys
T = sys.argv[1]
try:
    r = json.loads([l for l in open(f'gpurun_out/{T}_bench_reference.json') if l.startswith('{')][-1])
    j = json.loads([l for l in open(f'gpurun_out/{T}_bench.json') if l.startswith('{')][-1])
    print('value', j['value'], 'e2e', j['e2e']['value'], 'frac', j['roofline']['frac'], 'ref', r['value'], 'e2e/ref', j['e2e']['value'] / r['value'])
    for c in j['configs'] or []:
        print(c['config']['workload'][:28], c['accel'], 'value %.1f' % c['value'], 'frac %.3f' % c['roofline']['frac'], 'launch_ms %.3f' % c['roofline']['launch_ms'], 'cpu', (c.get('cpu_baseline') or {}).get('value'))
except Exception as e:
    print('bench parse failed', e)
PY
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${T}_bench_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/${T}_ncu_bench.log 2>&1; echo "ncu launch list rc=$?"
for a in "2 4" "2 1"; do timeout 120 python scripts/frame_timing.py $a 2>&1 | head -4; done > gpurun_out/${T}_frame_pipelines_c2.txt; cat gpurun_out/${T}_frame_pipelines_c2.txt
