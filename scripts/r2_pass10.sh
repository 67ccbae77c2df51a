#!/bin/bash
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/p10_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/p10_pytest.log
python scripts/frame_timing.py 2 2>&1 | head -4
timeout 600 python bench.py --no-extras --no-cpu-baseline > gpurun_out/p10_bench.json 2> gpurun_out/p10_bench_err.log; echo "bench rc=$?"; tail -3 gpurun_out/p10_bench_err.log
python - <<'PY'
import json
j = json.loads([l for l in open('gpurun_out/p10_bench.json') if l.startswith('{')][-1])
print('value', j['value'], 'e2e', j['e2e']['value'], 'ms/step', j['ms_per_step'])
PY
