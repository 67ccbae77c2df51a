#!/bin/bash
# Round-2 pass 11 (one B200): range-tested fast forms -- rcp_sqrt_ inside normalize (one range test instead of two
# intrinsics' worth), sqrt_x2 inside the paired normal draws -- and the op loop unrolled by two, against the default
# build (pair scan + combined miss test + division-free item mapping).  Digests + kernel time per variant; the
# every-pattern device-math tests and the parity suites on the variant with both fast forms.
mkdir -p gpurun_out
out=gpurun_out/r2_pass11.txt; : > $out
for v in default fnorm sq2 fnsq2 unr2; do
  lib=""; [ $v != default ] && lib=$PWD/build/variants/libsrt_$v.so
  echo "== $v" | tee -a $out
  SRT_LIB=$lib timeout 300 python scripts/variant_time.py 2 4 1 3 2>&1 | tee -a $out
done
( SRT_LIB=$PWD/build/variants/libsrt_fnsq2.so timeout 600 python -m pytest tests/test_gpu_device_math.py tests/test_gpu_analytic_scan.py tests/test_gpu_parity.py -x -q ) > gpurun_out/r2_pass11_pytest_fnsq2.log 2>&1; echo "pytest(fnsq2) rc=$?" | tee -a $out
tail -3 gpurun_out/r2_pass11_pytest_fnsq2.log | tee -a $out
