#!/bin/bash
# Round-2 pass 12 (one B200): path-state registers carried across trips of the wavefront loop instead of re-created.
mkdir -p gpurun_out
out=gpurun_out/r2_pass12.txt; : > $out
for v in default carry; do
  lib=""; [ $v != default ] && lib=$PWD/build/variants/libsrt_$v.so
  echo "== $v" | tee -a $out
  SRT_LIB=$lib timeout 300 python scripts/variant_time.py 2 4 1 2>&1 | tee -a $out
done
( SRT_LIB=$PWD/build/variants/libsrt_carry.so timeout 600 python -m pytest tests/test_gpu_analytic_scan.py tests/test_gpu_parity.py tests/test_gpu_properties.py -x -q ) > gpurun_out/r2_pass12_pytest_carry.log 2>&1; echo "pytest(carry) rc=$?" | tee -a $out
tail -3 gpurun_out/r2_pass12_pytest_carry.log | tee -a $out
