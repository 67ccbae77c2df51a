#!/bin/bash
# Round-2 GPU pass 2 (one B200): full GPU test-suite (no -x), bench with BVH entries.
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q ) > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2b_pytest.log
grep -E "passed|failed|FAILED|Error" gpurun_out/r2b_pytest.log | tail -15
( time python bench.py --steps 5 --no-cpu-baseline ) > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench_err.log; echo "bench rc=$?"; tail -3 gpurun_out/r2b_bench_err.log
python - <<'PY'
import json
try:
    j = json.loads([l for l in open('gpurun_out/r2b_bench.json') if l.startswith('{')][-1])
    print('value', j['value'], 'e2e', j['e2e']['value'], 'frac', j['roofline']['frac'])
    for c in j['configs'] or []:
        print(c['config']['workload'][:28], c['accel'], 'value %.1f' % c['value'], 'frac %.3f' % c['roofline']['frac'], 'launch_ms %.3f' % c['roofline']['launch_ms'], 'Gtests/s %.1f' % c['roofline']['Gtests_per_s'])
except Exception as e:
    print('bench parse failed', e)
PY
