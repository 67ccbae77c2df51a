#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -3
for lib in build/variants/libsrt_*.so; do
  echo "== $lib"
  SRT_LIB=$PWD/$lib python scripts/variant_time.py 1 2 2>&1 | tail -2
done | tee gpurun_out/r2h_variants.txt
