#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -3
python scripts/variant_time.py 1 2 3 5 2>&1 | tail -4
python bench.py --steps 10 --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import sys, json
j = json.loads(sys.stdin.read()); print('value %.1f  e2e %.1f' % (j['value'], j['e2e']['value']))"
