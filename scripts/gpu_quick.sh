#!/bin/bash
# quick GPU check: parity tests + stage timings (used during optimisation)
set -x
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline 2>&1 | python -c "
import sys, json
for line in sys.stdin:
    line=line.strip()
    if line.startswith('{'):
        d=json.loads(line); print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'launches/step', d['gpu_launches']/d['steps'])
        for k,v in d['roofline']['kernels'].items(): print('  ', k, v)
        print('clocks', d['clocks'])
    else: print(line)
"
