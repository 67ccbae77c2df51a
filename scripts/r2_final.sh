#!/bin/bash
# Round-2 evidence pass on one B200: GPU tests, both bench arms, ncu launch list, ncu --set full captures of the three
# render-kernel builds (config 2 batched + its accumulate epilogue, config 3, config 5).  Outputs -> gpurun_out/r2_*.
mkdir -p gpurun_out
T=${1:-r2}
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm,memory.total --format=csv > gpurun_out/${T}_environment.txt
echo "host cores: $(nproc)" >> gpurun_out/${T}_environment.txt; lscpu | grep -E "Model name" >> gpurun_out/${T}_environment.txt
( time python -m pytest tests -m gpu -q ) > gpurun_out/${T}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed" gpurun_out/${T}_pytest_gpu.log | tail -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_ref_err.log; echo "ref rc=$?"
( time python bench.py ) > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench_err.log; echo "bench rc=$?"; tail -3 gpurun_out/${T}_bench_err.log
python - $T <<'PY'
import json, sys
T = sys.argv[1]
try:
    r = json.loads([l for l in open(f'gpurun_out/{T}_bench_reference.json') if l.startswith('{')][-1])
    j = json.loads([l for l in open(f'gpurun_out/{T}_bench.json') if l.startswith('{')][-1])
    print('value', j['value'], 'e2e', j['e2e']['value'], 'frac', j['roofline']['frac'], 'ref', r['value'], 'e2e/ref', j['e2e']['value'] / r['value'])
    for c in j['configs'] or []:
        print(c['config']['workload'][:28], c['accel'], 'value %.1f' % c['value'], 'frac %.3f' % c['roofline']['frac'], 'launch_ms %.3f' % c['roofline']['launch_ms'], 'Gtests/s %.1f' % c['roofline']['Gtests_per_s'], 'cpu', (c.get('cpu_baseline') or {}).get('value'))
except Exception as e:
    print('bench parse failed', e)
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${T}_bench_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/${T}_ncu_bench.log 2>&1; echo "ncu launch list rc=$?"
scripts/ncu_capture.sh 2 ${T}_prof_c2 16 - batch
scripts/ncu_capture.sh 3 ${T}_prof_c3 2
scripts/ncu_capture.sh 5 ${T}_prof_c5 2
# frame pipelines (srt_render_frame): separate steps / epilogue into the pinned vector (auto)
for a in "2 1" "2 2" "2 4" "2 8" "1 1" "3 4"; do python scripts/frame_timing.py $a 2>&1 | head -4; done > gpurun_out/${T}_frame_pipelines.txt; cat gpurun_out/${T}_frame_pipelines.txt
# the whole GPU suite once more on the build with device-side invariant checks
if [ -f build/variants/libsrt_checks.so ]; then ( SRT_LIB=$PWD/build/variants/libsrt_checks.so timeout 1500 python -m pytest tests -m gpu -q -x ) > gpurun_out/${T}_pytest_gpu_checks.log 2>&1; echo "checks rc=$?"; tail -2 gpurun_out/${T}_pytest_gpu_checks.log; fi
