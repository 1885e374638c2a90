#!/bin/bash
# prints ms/step and the per-phase device times of bench.py for the current environment; extra args go to bench.py
timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline "$@" 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('%.2f ms/step' % d['ms_per_step'], {k: v['ms'] for k, v in d['phases_ms'].items()})"
