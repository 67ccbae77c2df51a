#!/bin/bash
# fused-frame pass: new parity tests, then the headline bench without the extra configs
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_fused_frame.py -x -q ) > gpurun_out/p9_pytest_fused.log 2>&1; echo "pytest fused rc=$?"; tail -5 gpurun_out/p9_pytest_fused.log
timeout 600 python bench.py --no-extras --no-cpu-baseline > gpurun_out/p9_bench.json 2> gpurun_out/p9_bench_err.log; echo "bench rc=$?"; tail -3 gpurun_out/p9_bench_err.log
python - <<'PY'
import json
j = json.loads([l for l in open('gpurun_out/p9_bench.json') if l.startswith('{')][-1])
print('value', j['value'], 'e2e', j['e2e']['value'], 'ms/step', j['ms_per_step'])
PY
