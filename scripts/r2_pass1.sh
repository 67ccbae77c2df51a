#!/bin/bash
# Round-2 GPU pass 1 (one B200): tests, bench both arms, ncu launch list, batched-kernel traffic capture.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm --format=csv > gpurun_out/r2_env.txt; nproc >> gpurun_out/r2_env.txt
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2a_pytest.log
tail -5 gpurun_out/r2a_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --impl reference --steps 2 --warmup 0 > gpurun_out/r2a_bench_reference.json 2> gpurun_out/r2a_bench_ref_err.log; echo "ref rc=$?"
( time python bench.py ) > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench_err.log; echo "bench rc=$?"; tail -3 gpurun_out/r2a_bench_err.log
python - <<'PY'
import json
try:
    j = json.loads([l for l in open('gpurun_out/r2a_bench.json') if l.startswith('{')][-1])
    print('value', j['value'], 'e2e', j['e2e']['value'], 'frac', j['roofline']['frac'], 'cpu', j['cpu_baseline'])
    for c in j['configs'] or []:
        print(c['config']['workload'][:28], 'value %.1f' % c['value'], 'frac %.3f' % c['roofline']['frac'], 'launch_ms %.3f' % c['roofline']['launch_ms'], 'cpu', c.get('cpu_baseline', {}).get('value'))
except Exception as e:
    print('bench parse failed', e)
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2a_bench_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2a_ncu_bench.log 2>&1; echo "ncu launch list rc=$?"
scripts/ncu_capture.sh 2 r2a_prof_c2_batch 16 - batch
scripts/ncu_capture.sh 3 r2a_prof_c3 2
