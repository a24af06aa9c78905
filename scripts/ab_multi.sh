#!/bin/bash
# baseline vs several variant builds, interleaved twice: scripts/ab_multi.sh lib1.so lib2.so ...
for rep in 1 2; do for v in "" "$@"; do
  SE_B200_LIB=$v timeout 300 python bench.py --steps 150 --warmup 10 --no-cpu-baseline 2>&1 | python -c "
import sys, json
for line in sys.stdin:
    line = line.strip()
    if line.startswith('{'):
        d = json.loads(line)
        print('lib=[$v] value', d['value'], 'ms', d['ms_per_step'], {k: v['ms'] for k, v in d['roofline']['kernels'].items()})
    elif line: print(line[:200])
"
done; done
