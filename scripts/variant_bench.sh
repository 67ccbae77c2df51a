#!/bin/bash
# usage: scripts/variant_bench.sh "<configs>" lib1.so lib2.so ...   (developer experiment helper)
cfgs="$1"; shift
for lib in "$@"; do
  echo "== $lib"
  SRT_LIB=$PWD/$lib python scripts/dev_gpu_check.py $cfgs 2>&1 | grep '"cfg"' | python -c "
import sys, json
for l in sys.stdin:
    j = json.loads(l); print('   cfg', j['cfg'], 'kernel_ms %.2f' % j['kernel_ms'], 'Msamples/s %.1f' % j['Msamples/s'], 'Gtests/s %.1f' % j['Gtests/s'])"
done
