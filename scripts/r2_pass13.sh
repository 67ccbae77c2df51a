#!/bin/bash
# Round-2 pass 13 (one B200): hit / sky ring records through 16-byte shared-memory accesses (SRT_RING_V4).
mkdir -p gpurun_out
out=gpurun_out/r2_pass13.txt; : > $out
for v in default v4; do
  lib=""; [ $v != default ] && lib=$PWD/build/variants/libsrt_$v.so
  echo "== $v" | tee -a $out
  SRT_LIB=$lib timeout 300 python scripts/variant_time.py 2 4 1 2>&1 | tee -a $out
done
( SRT_LIB=$PWD/build/variants/libsrt_v4.so timeout 600 python -m pytest tests/test_gpu_analytic_scan.py tests/test_gpu_parity.py -x -q ) > gpurun_out/r2_pass13_pytest_v4.log 2>&1; echo "pytest(v4) rc=$?" | tee -a $out
tail -2 gpurun_out/r2_pass13_pytest_v4.log | tee -a $out
