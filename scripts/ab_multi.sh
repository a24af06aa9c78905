#!/bin/bash
# current build vs several variant builds, interleaved REPS times: REPS=2 scripts/ab_multi.sh lib1.so lib2.so ...
for rep in $(seq 1 ${REPS:-2}); do for v in "" "$@"; do
  SE_B200_LIB=$v timeout 300 python bench.py --steps 300 --warmup 10 --no-cpu-baseline 2>&1 | python -c "
import sys, json
for line in sys.stdin:
    line = line.strip()
    if line.startswith('{'):
        d = json.loads(line)
        print('lib=[$v] value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], {k: v['ms'] for k, v in d['roofline']['kernels'].items()})
    elif line: print(line[:200])
"
done; done
