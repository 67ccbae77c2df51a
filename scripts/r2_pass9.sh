#!/bin/bash
# Round-2 pass 9 (one B200): the pair-scan variants of the analytic wavefront kernel against the default build --
# canvas digests (bit-exact builds agree), kernel time, then the whole GPU suite and a short bench on the fastest one.
mkdir -p gpurun_out
out=gpurun_out/r2_pass9.txt; : > $out
for v in default pair pairsq; do
  lib=""; [ $v != default ] && lib=$PWD/build/variants/libsrt_$v.so
  echo "== $v" | tee -a $out
  SRT_LIB=$lib timeout 300 python scripts/variant_time.py 2 4 2>&1 | tee -a $out
done
best=$(python - <<'PY'
import re
cur=None; res={}
for l in open('gpurun_out/r2_pass9.txt'):
    if l.startswith('== '): cur=l.split()[1]; res[cur]={}
    m=re.match(r'cfg (\d+)\s+kernel_ms\s+([\d.]+).*canvas (\w+)',l)
    if m: res[cur][int(m.group(1))]=(float(m.group(2)),m.group(3))
ref=res.get('default',{})
ok=[v for v in res if v!='default' and res[v] and all(c in ref and res[v][c][1]==ref[c][1] for c in res[v]) and len(res[v])==len(ref)]
ok.sort(key=lambda v:res[v][2][0])
print(ok[0] if ok and res[ok[0]][2][0] < ref[2][0]*0.995 else 'none')
PY
)
echo "best variant: $best" | tee -a $out
if [ "$best" != none ]; then
  lib=$PWD/build/variants/libsrt_$best.so
  ( time SRT_LIB=$lib timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2_pass9_pytest_$best.log 2>&1; echo "pytest($best) rc=$?" | tee -a $out
  tail -3 gpurun_out/r2_pass9_pytest_$best.log | tee -a $out
  SRT_LIB=$lib timeout 600 python bench.py --no-cpu-baseline --no-extras > gpurun_out/r2_pass9_bench_$best.json 2> gpurun_out/r2_pass9_bench_err.log; echo "bench rc=$?" | tee -a $out
  python -c "
import json
j=json.loads([l for l in open('gpurun_out/r2_pass9_bench_$best.json') if l.startswith('{')][-1]); print('value',j['value'],'e2e',j['e2e']['value'])" | tee -a $out
fi
