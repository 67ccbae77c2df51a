#!/bin/bash
# Round-2 multi-GPU pass: N = $1 GPUs of one box.  mgpu parity script, bench both arms under torchrun.
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
( time python -m pytest tests/test_gpu_distributed.py tests/test_gpu_round2.py -q -k "two_gpu or work_item" ) > gpurun_out/r2_pytest_n$N.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest_n$N.log
$TR scripts/mgpu_check.py > gpurun_out/r2_mgpu_check_n$N.json 2> gpurun_out/r2_mgpu_check_n$N.err; echo "mgpu_check rc=$?"; cat gpurun_out/r2_mgpu_check_n$N.json; tail -5 gpurun_out/r2_mgpu_check_n$N.err
$TR bench.py --impl reference --gpus $N --steps 2 --warmup 0 > gpurun_out/r2_bench_reference_n$N.json 2> gpurun_out/r2_bench_ref_n$N.err; echo "ref rc=$?"
( time $TR bench.py --gpus $N ) > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err; echo "bench rc=$?"; tail -8 gpurun_out/r2_bench_n$N.err
python - $N <<'PY'
import json, sys
n = sys.argv[1]
try:
    r = json.loads([l for l in open(f'gpurun_out/r2_bench_reference_n{n}.json') if l.startswith('{')][-1])
    print('reference arm', r['value'], r['cpu_baseline'])
    j = json.loads([l for l in open(f'gpurun_out/r2_bench_n{n}.json') if l.startswith('{')][-1])
    print('value', j['value'], 'e2e', j['e2e']['value'], 'ms/step', j['ms_per_step'], 'cpu', j['cpu_baseline'], 'parity', j['mgpu_parity'])
    for s in j['strong'] or []:
        print({k: v for k, v in s.items() if k not in ('config', 'timing', 'exchange')})
except Exception as e:
    print('bench parse failed', e)
PY
