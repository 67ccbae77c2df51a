#!/bin/bash
# sweep-loop variants: parity tests on the tree build, then timings of every variant
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_properties.py tests/test_gpu_round2.py -q -x -k "not work_item" 2>&1 | tail -4
for lib in build/variants/libsrt_*.so; do
  echo "== $lib"
  SRT_LIB=$PWD/$lib python scripts/variant_time.py 3 5 2>&1 | tail -2
done | tee gpurun_out/r2d_variants.txt
