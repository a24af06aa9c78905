#!/bin/bash
# A/B of programmatic dependent launch and the integrate work assignment: scripts/ab_pdl.sh [variant.so ...]
run() {
  env "$@" timeout 300 python bench.py --steps 300 --warmup 10 --no-cpu-baseline 2>&1 | python -c "
import sys, json
for line in sys.stdin:
    line = line.strip()
    if line.startswith('{'):
        d = json.loads(line)
        print('$*', 'value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], {k: v['ms'] for k, v in d['roofline']['kernels'].items()})
    elif line: print(line[:200])
"
}
for rep in 1 2; do
  run SE_B200_NO_PDL=0
  run SE_B200_NO_PDL=1
  for v in "$@"; do run SE_B200_LIB=$v; done
done
