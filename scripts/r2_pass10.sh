#!/bin/bash
# Round-2 pass 10 (one B200): follow-ups to the pair scan (now the default build) -- combined miss test, division-free
# item mapping, op loop unrolled by two.  Canvas digests + kernel time per variant; the new analytic-scan tests and the
# randomised launch-parameter tests on every variant that is bit-identical.
mkdir -p gpurun_out
out=gpurun_out/r2_pass10.txt; : > $out
for v in default any fdiv anyfdiv unr2; do
  lib=""; [ $v != default ] && lib=$PWD/build/variants/libsrt_$v.so
  echo "== $v" | tee -a $out
  SRT_LIB=$lib timeout 300 python scripts/variant_time.py 2 4 1 2>&1 | tee -a $out
done
( SRT_LIB=$PWD/build/variants/libsrt_anyfdiv.so timeout 600 python -m pytest tests/test_gpu_analytic_scan.py tests/test_gpu_parity.py -x -q ) > gpurun_out/r2_pass10_pytest_anyfdiv.log 2>&1; echo "pytest(anyfdiv) rc=$?" | tee -a $out
tail -3 gpurun_out/r2_pass10_pytest_anyfdiv.log | tee -a $out
( timeout 600 python -m pytest tests/test_gpu_analytic_scan.py -x -q ) > gpurun_out/r2_pass10_pytest_default.log 2>&1; echo "pytest(default, analytic scan) rc=$?" | tee -a $out
tail -3 gpurun_out/r2_pass10_pytest_default.log | tee -a $out
