#!/bin/bash
# A/B of two builds of the library in the same GPU call: scripts/ab_bench.sh <lib_b.so> [steps]
B=$1; STEPS=${2:-200}
for v in "" "$B" "" "$B"; do
  SE_B200_LIB=$v timeout 300 python bench.py --steps $STEPS --warmup 10 --no-cpu-baseline 2>&1 | python -c "
import sys, json
for line in sys.stdin:
    line = line.strip()
    if line.startswith('{'):
        d = json.loads(line)
        print('lib=[$v] value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], {k: v['ms'] for k, v in d['roofline']['kernels'].items()})
    elif line: print(line[:300])
"
done
